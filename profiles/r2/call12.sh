set -x
mkdir -p gpurun_out/r2
TFSC_GT_TRACE=1 timeout 300 python profiles/r2/trace_gemm.py > gpurun_out/r2/trace_gemm.jsonl 2>&1; cat gpurun_out/r2/trace_gemm.jsonl
