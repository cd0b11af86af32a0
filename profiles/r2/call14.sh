set -x
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests/test_gpu_graph.py tests/test_model_pins.py -m gpu -x -q > gpurun_out/r2/gpu_graph_14.log 2>&1; tail -4 gpurun_out/r2/gpu_graph_14.log
timeout 300 python profiles/time_graph.py resnet50 1 8 16 > gpurun_out/r2/time_resnet_14.jsonl 2>&1; tail -3 gpurun_out/r2/time_resnet_14.jsonl
timeout 300 python profiles/time_graph.py bert 1 8 > gpurun_out/r2/time_bert_14.jsonl 2>&1; tail -2 gpurun_out/r2/time_bert_14.jsonl
