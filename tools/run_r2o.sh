set -x
D=gpurun_out/r2o; mkdir -p $D
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:kcf_ -c 200 --csv --log-file $D/launches_part.csv python tools/part_profile.py 2 > $D/ncu_launch.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r2o/launches_part.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[h]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.defaultdict(list)
for r in rows[h+1:]:
    if len(r)>vi:
        v=float(r[vi].replace(',','')); u=r[ui]
        ms=v/1e6 if u in('ns','nsecond') else (v/1e3 if u in ('us','usecond') else v)
        agg[r[ki][:70]].append(ms)
for k,v in agg.items(): print(f"{k:70s} n={len(v):5d} avg={sum(v)/len(v):8.4f} ms sum={sum(v):9.3f} ms")
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kcf_xg_answer -s 2 -c 1 -o $D/prof_answer -f python tools/part_profile.py 2 > $D/ncu_answer.log 2>&1
ls -la $D
