set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_examples.py tests/test_savedmodel.py tests/test_gpu_frontends.py -m gpu -x -q > gpurun_out/r2/gpu_examples_7.log 2>&1; tail -25 gpurun_out/r2/gpu_examples_7.log
# launch lists (cold, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2/launches_resnet50_b8.csv python profiles/time_graph.py resnet50 8 > gpurun_out/r2/launches_resnet50_b8.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2/launches_bert_b8.csv python profiles/time_graph.py bert 8 > gpurun_out/r2/launches_bert_b8.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2/launches_bench.csv python bench.py --steps 1 --warmup 1 --models-per-gpu 32 --arena-gib 40 --no-cpu-baseline --skip-e2e --no-extras --preheat-s 0 > gpurun_out/r2/launches_bench.log 2>&1
# --set full of the two dominant kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_cluster -s 3 -c 3 -o gpurun_out/r2/dense_cluster_R8 python profiles/prof_dense.py 8 8 > gpurun_out/r2/ncu_cluster.log 2>&1
TFSC_DENSE_VARIANT=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_tc -s 3 -c 3 -o gpurun_out/r2/dense_tc_RP64 python profiles/prof_dense.py 64 8 > gpurun_out/r2/ncu_tc.log 2>&1
ls -la gpurun_out/r2/*.ncu-rep
python - <<'P'
import csv, collections, sys
for f in ("launches_resnet50_b8","launches_bert_b8","launches_bench"):
    try:
        rows=[r for r in csv.reader(open(f"gpurun_out/r2/{f}.csv")) if len(r)>10 and r[0].isdigit()]
        agg=collections.defaultdict(lambda:[0,0.0])
        for r in rows:
            name=r[4].split("(")[0][:60]; v=float(r[-1].replace(",",""))
            agg[name][0]+=1; agg[name][1]+=v
        tot=sum(v[1] for v in agg.values())
        print(f, "launches", len(rows), "total", tot)
        for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:12]: print("   %-60s n=%4d  sum=%10.1f  avg=%8.1f  share=%.3f"%(k,v[0],v[1],v[1]/v[0],v[1]/tot))
    except Exception as e: print(f,"ERR",e)
P
