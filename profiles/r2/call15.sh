set -x
mkdir -p gpurun_out/r2
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2/gpu_suite_15.log 2>&1; tail -6 gpurun_out/r2/gpu_suite_15.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2/bench_n1_15.json 2> gpurun_out/r2/bench_n1_15.err; tail -3 gpurun_out/r2/bench_n1_15.err
python - <<'P'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2/bench_n1_15.json") if l.startswith("{")][-1])
    print("value", d["value"], "frac", d["roofline"]["frac"], "us/launch", d["roofline"]["avg_launch_us"], "e2e", d["e2e"]["value"], d["e2e"]["p50_ms"])
    print("q5", d["e2e"]["qps_at_p50_5ms"]); print("extra", d.get("extra")); print("clocks", d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"), d["config"].get("cpu_affinity"))
except Exception as e: print("ERR", e)
P
