set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_graph.py -m gpu -x -q > gpurun_out/r2/gpu_graph_8.log 2>&1; tail -25 gpurun_out/r2/gpu_graph_8.log
timeout 900 python -m pytest tests/test_model_pins.py tests/test_gpu_arena.py -m gpu -x -q > gpurun_out/r2/gpu_pins_8.log 2>&1; tail -15 gpurun_out/r2/gpu_pins_8.log
timeout 300 python profiles/time_graph.py resnet50 1 8 16 > gpurun_out/r2/time_resnet_8.jsonl 2>&1; cat gpurun_out/r2/time_resnet_8.jsonl | tail -4
timeout 300 python profiles/time_graph.py bert 1 8 > gpurun_out/r2/time_bert_8.jsonl 2>&1; cat gpurun_out/r2/time_bert_8.jsonl | tail -3
TFSC_CONV_TC=0 timeout 300 python profiles/time_graph.py resnet50 8 > gpurun_out/r2/time_resnet_8_explicit.jsonl 2>&1; tail -2 gpurun_out/r2/time_resnet_8_explicit.jsonl
