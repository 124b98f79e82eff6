#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_loss_gpu.py tests/test_surface_gpu.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
timeout 200 python tools/bench_loss.py 2>&1 | tee gpurun_out/loss_bw.txt
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"laplace_train|aggregate" --csv --log-file gpurun_out/loss_ncu.csv python tools/bench_loss.py --reps 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/loss_ncu.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; st=i+1; break
k=h.index('Kernel Name'); m=h.index('Metric Name'); v=h.index('Metric Value'); idc=h.index('ID')
d={}
for r in rows[st:]:
    if len(r)<=v: continue
    d.setdefault((r[idc],r[k][:50]),{})[r[m]]=float(r[v].replace(',',''))
seen=set()
for (i,n),mm in d.items():
    t=mm.get('gpu__time_duration.sum',0)/1e3; b=(mm.get('dram__bytes_read.sum',0)+mm.get('dram__bytes_write.sum',0))/1e6
    key=(n,round(b))
    if key in seen: continue
    seen.add(key)
    print(f"{n:50s} {t:8.1f} us  dram {b:8.1f} MB  {b/t/1e3 if t else 0:6.2f} TB/s")
PY
