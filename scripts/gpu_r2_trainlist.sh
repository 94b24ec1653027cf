#!/bin/bash
# ncu launch list of ONE whole training step per denoiser: which kernels run (is anything cuDNN / cuBLAS left?)
mkdir -p gpurun_out
for d in ffdnet SimpleCNN; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/train_launches_$d.csv python scripts/bench_train.py --denoiser $d --steps 1 --warmup 0 > /dev/null 2>&1; echo "$d exit $?"
  python - <<PY
import csv,collections
rows=[l for l in open("gpurun_out/train_launches_$d.csv") if l.startswith(chr(34))]
agg=collections.OrderedDict()
for r in csv.DictReader(rows):
    n=r["Kernel Name"].split("(")[0][:90]; v=float(r["Metric Value"].replace(",","")); u=r["Metric Unit"]
    v = v/1e3 if u=="ns" else (v*1e3 if u=="ms" else v)
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(v[1] for v in agg.values())
with open("gpurun_out/train_step_kernels_$d.md","w") as f:
    f.write("# ncu launch list of ONE DE-GAP-$d training step (bench_train.py --steps 1 --warmup 0: includes first-call setup), %d launches, %.1f ms of kernel time\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n" % (sum(v[0] for v in agg.values()), tot/1e3))
    for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]):
        f.write("| \`%s\` | %d | %.1f | %.2f %% |\n" % (k, v[0], v[1], 100*v[1]/tot))
lib=[k for k in agg if any(t in k.lower() for t in ("cudnn","cutlass","gemm","implicit","xmma","sm90","sm80","wgrad_alg","dgrad"))]
print("$d: distinct kernels", len(agg), "library-looking:", lib)
PY
done
