set -x
mkdir -p gpurun_out/r2
nvidia-smi topo -m | head -6
TFSC_REQUIRE_MULTI=1 timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_forward.py -m gpu -x -q > gpurun_out/r2/gpu_multi_6.log 2>&1; tail -5 gpurun_out/r2/gpu_multi_6.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2/bench_n2_fwd.json 2> gpurun_out/r2/bench_n2_fwd.err
tail -5 gpurun_out/r2/bench_n2_fwd.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --forward-frac 0 --no-extras > gpurun_out/r2/bench_n2_nofwd.json 2> gpurun_out/r2/bench_n2_nofwd.err
tail -5 gpurun_out/r2/bench_n2_nofwd.err
python - <<'P'
import json
for f in ("bench_n2_fwd","bench_n2_nofwd"):
    try:
        d=json.load(open(f"gpurun_out/r2/{f}.json"))
        print(f, "value", d["value"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["p50_ms"], "fail", d["e2e"]["failed"])
        print("  forward", d.get("forward")); print("  q5", d["e2e"].get("qps_at_p50_5ms")); print("  per_rank", d.get("per_rank"))
    except Exception as ex: print(f, "ERR", ex)
P
