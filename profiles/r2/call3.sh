set -x
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2/gpu_tc_3.log 2>&1; tail -3 gpurun_out/r2/gpu_tc_3.log
for p in 0 1; do TFSC_PDL=$p timeout 300 python profiles/time_dense_total.py 3 300 16,32,48,64; done > gpurun_out/r2/dense_tc_pdl.jsonl 2> gpurun_out/r2/dense_tc_pdl.err
cat gpurun_out/r2/dense_tc_pdl.jsonl
for s in none smi nvml; do
timeout 600 python bench.py --steps 20 --warmup 3 --samplers $s --skip-e2e --no-cpu-baseline > gpurun_out/r2/bench_ab_$s.json 2> gpurun_out/r2/bench_ab_$s.err
tail -5 gpurun_out/r2/bench_ab_$s.err
done
python - <<'P'
import json
for f in ("none","smi","nvml"):
    try:
        d=json.load(open(f"gpurun_out/r2/bench_ab_{f}.json")); print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"])
    except Exception as e: print(f, "ERR", e)
P
