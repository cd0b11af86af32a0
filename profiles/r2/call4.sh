set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -x -q > gpurun_out/r2/gpu_fwd_4.log 2>&1; tail -30 gpurun_out/r2/gpu_fwd_4.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/gpu_suite_4.log 2>&1; tail -15 gpurun_out/r2/gpu_suite_4.log
