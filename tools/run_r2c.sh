set -x
D=gpurun_out/r2c; mkdir -p $D
timeout 900 python -m pytest tests -m gpu -x -q > $D/tests.log 2>&1; echo "tests rc=$?" >> $D/tests.log; tail -5 $D/tests.log
for v in pf1 pf2 pf3 w48 pf1w48; do
  KCF_LIB_PATH=$PWD/kcftools_b200/libkcfgpu_$v.so timeout 300 python bench.py --only resident --steps 20 > $D/variant_$v.json 2> $D/variant_$v.err; python -c "import json;j=json.load(open('$D/variant_$v.json'));print('$v', j['value'], j['ms_per_step'], j['db_load_s'])"
done
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $D/bench_c2.json 2> $D/bench_c2.err; echo "bench rc=$?"; tail -25 $D/bench_c2.err
cut -c1-300 $D/bench_c2.json
