set -x
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_forward.py -m gpu -x -q > gpurun_out/r2/gpu_fwd_17.log 2>&1; tail -15 gpurun_out/r2/gpu_fwd_17.log
