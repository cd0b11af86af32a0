set -x
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_graph.py -m gpu -x -q -k "im2col or resnet" > gpurun_out/r2/gpu_graph_10.log 2>&1; tail -4 gpurun_out/r2/gpu_graph_10.log
timeout 300 python profiles/time_graph.py resnet50 8 16 > gpurun_out/r2/time_resnet_10.jsonl 2>&1; tail -2 gpurun_out/r2/time_resnet_10.jsonl
# ncu --set full on representative graph GEMMs: BERT QKV (24,8,1) etc. and ResNet 3x3 implicit convs
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 150 -c 8 -o gpurun_out/r2/gemm_tc_bert python profiles/time_graph.py bert 8 > gpurun_out/r2/ncu_gemm_bert.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 120 -c 12 -o gpurun_out/r2/gemm_tc_resnet python profiles/time_graph.py resnet50 8 > gpurun_out/r2/ncu_gemm_resnet.log 2>&1
ls -la gpurun_out/r2/gemm_tc_*.ncu-rep
