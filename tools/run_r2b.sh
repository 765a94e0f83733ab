set -x
D=gpurun_out/r2b; mkdir -p $D
timeout 900 python -m pytest tests -m gpu -x -q > $D/tests.log 2>&1; echo "tests rc=$?" >> $D/tests.log; tail -3 $D/tests.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $D/bench_c2.json 2> $D/bench_c2.err; echo "bench rc=$?"; tail -25 $D/bench_c2.err
for lf in 0.15 0.3 0.5 0.7 0.9; do
  timeout 300 python bench.py --only resident --lf $lf --steps 10 > $D/density_$lf.json 2> $D/density_$lf.err; tail -1 $D/density_$lf.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:kcf_ -c 400 --csv --log-file $D/launches_c2.csv python bench.py --only resident,e2e,cold --steps 3 --warmup 3 --e2e-steps 1 > $D/ncu_launch.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:kcf_screen_kernel -s 3 -c 1 --csv --log-file $D/traffic_c2.csv python bench.py --only resident --steps 2 --warmup 3 > $D/ncu_traffic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kcf_screen_kernel -s 3 -c 1 -o $D/prof_screen_c2 -f python bench.py --only resident --steps 2 --warmup 3 > $D/ncu_full.log 2>&1
ls -la $D; cut -c1-400 $D/bench_c2.json
