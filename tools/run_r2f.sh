set -x
D=gpurun_out/r2f; mkdir -p $D
timeout 300 python bench.py --only resident --steps 20 > $D/bench_base.json 2> $D/bench_base.err; python -c "import json;j=json.load(open('$D/bench_base.json'));print('base', j['value'], j['ms_per_step'])"
KCF_LIB_PATH=$PWD/kcftools_b200/libkcfgpu_q32.so timeout 300 python bench.py --only resident --steps 20 > $D/bench_q32.json 2> $D/bench_q32.err; python -c "import json;j=json.load(open('$D/bench_q32.json'));print('q32', j['value'], j['ms_per_step'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kcf_ingest_kernel -s 50 -c 1 -o $D/prof_ingest -f python bench.py --only resident --steps 2 --warmup 3 > $D/ncu_ingest.log 2>&1
ls -la $D
