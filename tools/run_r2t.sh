set -x
D=gpurun_out/r2t; mkdir -p $D
timeout 900 python -m pytest tests -m gpu -x -q > $D/tests.log 2>&1; echo "tests rc=$?" >> $D/tests.log; tail -4 $D/tests.log
for v in base p48 p96; do
  if [ $v = base ]; then unset KCF_LIB_PATH; else export KCF_LIB_PATH=$PWD/kcftools_b200/libkcfgpu_$v.so; fi
  timeout 300 python bench.py --only resident,e2e --steps 5 --e2e-steps 8 > $D/e2e_$v.json 2> $D/e2e_$v.err; python -c "import json;j=json.load(open('$D/e2e_$v.json'));print('$v', j['e2e']['ms_per_step'], j['e2e']['ms_each_step'])"
done
