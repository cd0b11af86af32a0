set -x
mkdir -p gpurun_out/r2
TFSC_BENCH_DUMP_S=170 timeout 225 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2/bench_n8b.json 2> gpurun_out/r2/bench_n8b.err
echo rc=$?
grep "bench +" gpurun_out/r2/bench_n8b.err | head -12
python - <<'P'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2/bench_n8b.json") if l.startswith("{")][-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "us/launch", d["roofline"]["avg_launch_us"], "e2e", d["e2e"]["value"], d["e2e"]["p50_ms"], d["e2e"]["p99_ms"], "fail", d["e2e"]["failed"], "cold", d["config"]["cold_load_s"])
    print("forward", d.get("forward")); print("q5", d["e2e"]["qps_at_p50_5ms"])
    for r in d.get("per_rank",[]): print("  ", r)
    print("pressure", d.get("cache_pressure"))
except Exception as e: print("ERR", e)
P
