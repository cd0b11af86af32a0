set -x
mkdir -p gpurun_out/r2
cat /sys/fs/cgroup/cpu.max; free -g | head -2
timeout 360 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2/bench_n8.json 2> gpurun_out/r2/bench_n8.err
echo rc=$?
tail -5 gpurun_out/r2/bench_n8.err | cut -c1-300
python - <<'P'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2/bench_n8.json") if l.startswith("{")][-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "us/launch", d["roofline"]["avg_launch_us"], "e2e", d["e2e"]["value"], d["e2e"]["p50_ms"], d["e2e"]["p99_ms"], "fail", d["e2e"]["failed"], "cold", d["config"]["cold_load_s"])
    print("forward", d.get("forward")); print("q5", d["e2e"]["qps_at_p50_5ms"]); print("sweep", d["e2e"]["sweep"])
    for r in d.get("per_rank",[]): print("  ", r)
    print("pressure", d.get("cache_pressure"))
except Exception as e: print("ERR", e)
P
