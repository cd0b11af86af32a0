set -x
mkdir -p gpurun_out/r2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2/bench_n1_v2.json 2> gpurun_out/r2/bench_n1_v2.err
tail -5 gpurun_out/r2/bench_n1_v2.err
python - <<'P'
import json
try:
    d=json.load(open("gpurun_out/r2/bench_n1_v2.json"))
    print("value", d["value"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "us/launch", d["roofline"]["avg_launch_us"])
    e=d["e2e"]; print("e2e", e["value"], e["p50_ms"], e["p99_ms"], e["mean_batch_rows"], "h2d", e["h2d_bytes_per_step"])
    print("sweep", e["sweep"]); print("qps@p50<5", e["qps_at_p50_5ms"])
    print("pressure", d.get("cache_pressure")); print("extra", d.get("extra")); print("cpu", d.get("cpu_baseline",{}).get("value"))
except Exception as ex: print("ERR", ex)
P
