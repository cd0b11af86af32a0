set -x
mkdir -p gpurun_out/r2
(go version; nvidia-smi topo -m; cat /sys/fs/cgroup/cpu.max; nproc; numactl -H 2>/dev/null | head -5; free -g | head -2) > gpurun_out/r2/probe_env.log 2>&1
timeout 200 python profiles/r2/probe_ipc.py > gpurun_out/r2/probe_ipc.log 2>&1
TFSC_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_experimental.py -m gpu -q -x > gpurun_out/r2/experimental.log 2>&1
for p in 0 1; do for v in 1 2 4 5; do TFSC_PDL=$p timeout 300 python profiles/time_dense_total.py $v; done; done > gpurun_out/r2/dense_ab.jsonl 2> gpurun_out/r2/dense_ab.err
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2/gpu_suite_1.log 2>&1; tail -3 gpurun_out/r2/gpu_suite_1.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2/bench_n1_base.json 2> gpurun_out/r2/bench_n1_base.err
tail -3 gpurun_out/r2/experimental.log; cat gpurun_out/r2/dense_ab.jsonl; cat gpurun_out/r2/probe_ipc.log | tail -3
