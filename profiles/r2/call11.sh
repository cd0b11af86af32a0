set -x
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests/test_gpu_graph.py tests/test_model_pins.py -m gpu -x -q > gpurun_out/r2/gpu_graph_11.log 2>&1; tail -12 gpurun_out/r2/gpu_graph_11.log
timeout 300 python profiles/time_graph.py resnet50 1 8 16 > gpurun_out/r2/time_resnet_11.jsonl 2>&1; tail -3 gpurun_out/r2/time_resnet_11.jsonl
timeout 300 python profiles/time_graph.py bert 1 8 > gpurun_out/r2/time_bert_11.jsonl 2>&1; tail -2 gpurun_out/r2/time_bert_11.jsonl
TFSC_GEMM_PERSIST=0 timeout 300 python profiles/time_graph.py resnet50 8 > gpurun_out/r2/time_resnet_11_nopersist.jsonl 2>&1; tail -1 gpurun_out/r2/time_resnet_11_nopersist.jsonl
TFSC_GEMM_PERSIST=0 timeout 300 python profiles/time_graph.py bert 8 > gpurun_out/r2/time_bert_11_nopersist.jsonl 2>&1; tail -1 gpurun_out/r2/time_bert_11_nopersist.jsonl
