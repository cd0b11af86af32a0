set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/gpu_suite_2.log 2>&1; tail -3 gpurun_out/r2/gpu_suite_2.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2/bench_n1_cluster.json 2> gpurun_out/r2/bench_n1_cluster.err
timeout 600 python bench.py --steps 20 --warmup 3 --samplers none --skip-e2e --no-cpu-baseline > gpurun_out/r2/bench_n1_cluster_nosampler.json 2> gpurun_out/r2/bench_n1_cluster_nosampler.err
python - <<'P'
import json
for f in ("bench_n1_cluster","bench_n1_cluster_nosampler"):
    try:
        d=json.load(open(f"gpurun_out/r2/{f}.json")); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"])
    except Exception as e: print(f, "ERR", e)
P
