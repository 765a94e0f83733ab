set -x
D=gpurun_out/r1h; mkdir -p $D
timeout 900 python -m pytest tests -m gpu -x -q > $D/tests.log 2>&1; echo "tests rc=$?" >> $D/tests.log
timeout 600 python bench.py > $D/bench_c2.json 2> $D/bench_c2.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $D/bench_reference_arm.json 2> $D/bench_reference_arm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:kcf_ -c 400 --csv --log-file $D/launches_c2.csv python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $D/ncu_launch.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:kcf_screen_kernel -s 3 -c 1 --csv --log-file $D/traffic_c2.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > $D/ncu_traffic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kcf_screen_kernel -s 3 -c 1 -o $D/prof_screen_c2 -f python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > $D/ncu_full.log 2>&1
tail -n 2 $D/tests.log; cut -c1-250 $D/bench_c2.json
