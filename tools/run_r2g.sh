set -x
D=gpurun_out/r2k; mkdir -p $D
timeout 900 python -m pytest tests -m gpu -x -q > $D/tests.log 2>&1; echo "tests rc=$?" >> $D/tests.log; tail -6 $D/tests.log
timeout 300 python bench.py --only resident,e2e,cold --steps 10 > $D/bench_quick.json 2> $D/bench_quick.err
python -c "
import json;j=json.load(open('$D/bench_quick.json'));print('value',j['value'],'e2e',j['e2e']['ms_per_step'],j['e2e']['h2d_copy_alone_ms'],'cold',json.dumps(j['e2e_cold']['runs']))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:kcf_in -c 1600 --csv --log-file $D/launches_ingest.csv python bench.py --only resident --steps 2 --warmup 3 > $D/ncu_launch.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r2k/launches_ingest.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[h]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.defaultdict(list)
for r in rows[h+1:]:
    if len(r)>vi:
        v=float(r[vi].replace(',','')); u=r[ui]
        ms=v/1e6 if u in('ns','nsecond') else (v/1e3 if u in ('us','usecond') else v)
        agg[r[ki][:60]].append(ms)
for k,v in agg.items(): print(f"{k:60s} n={len(v):5d} avg={sum(v)/len(v):8.4f} ms sum={sum(v):9.3f} ms")
PY
