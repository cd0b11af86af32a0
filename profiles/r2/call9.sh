set -x
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests/test_gpu_graph.py tests/test_model_pins.py -m gpu -x -q > gpurun_out/r2/gpu_graph_9.log 2>&1; tail -25 gpurun_out/r2/gpu_graph_9.log
timeout 300 python profiles/time_graph.py resnet50 1 8 16 > gpurun_out/r2/time_resnet_9.jsonl 2>&1; tail -3 gpurun_out/r2/time_resnet_9.jsonl
timeout 300 python profiles/time_graph.py bert 1 8 > gpurun_out/r2/time_bert_9.jsonl 2>&1; tail -2 gpurun_out/r2/time_bert_9.jsonl
TFSC_GEMM_SPLITK=0 timeout 300 python profiles/time_graph.py resnet50 8 > gpurun_out/r2/time_resnet_9_nosplit.jsonl 2>&1; tail -1 gpurun_out/r2/time_resnet_9_nosplit.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 240 --csv --log-file gpurun_out/r2/launches_resnet50_b8_v2.csv python profiles/time_graph.py resnet50 8 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 270 --csv --log-file gpurun_out/r2/launches_bert_b8_v2.csv python profiles/time_graph.py bert 8 > /dev/null 2>&1
python - <<'P'
import csv, collections
for f in ("launches_resnet50_b8_v2","launches_bert_b8_v2"):
    try:
        rows=[r for r in csv.reader(open(f"gpurun_out/r2/{f}.csv")) if len(r)>10 and r[0].isdigit()]
        agg=collections.defaultdict(lambda:[0,0.0])
        for r in rows:
            name=r[4].split("(")[0][:50]+" grid="+r[-4]+r[-5] if False else r[4].split("(")[0][:60]
            v=float(r[-1].replace(",","")); agg[name][0]+=1; agg[name][1]+=v
        tot=sum(v[1] for v in agg.values())
        print(f, "launches", len(rows), "total_us", tot/1e3)
        for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:10]: print("   %-60s n=%4d sum_us=%9.1f avg_us=%7.1f share=%.3f"%(k,v[0],v[1]/1e3,v[1]/v[0]/1e3,v[1]/tot))
        # the 12 slowest individual launches with their grid sizes
        hdr=None
        for r in csv.reader(open(f"gpurun_out/r2/{f}.csv")):
            if "Kernel Name" in r: hdr=r; break
        gi=hdr.index("Grid Size") if hdr and "Grid Size" in hdr else None
        slow=sorted(rows,key=lambda r:-float(r[-1].replace(",","")))[:12]
        for r in slow: print("      slow:", r[4].split("(")[0][:40], "grid", r[gi] if gi is not None else "?", "us", float(r[-1].replace(",",""))/1e3)
    except Exception as e: print(f,"ERR",e)
P
