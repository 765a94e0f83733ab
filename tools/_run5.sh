set -x
D=gpurun_out/r1f; mkdir -p $D
timeout 900 python -m pytest tests -m gpu -x -q > $D/tests.log 2>&1; echo "tests rc=$?" >> $D/tests.log
timeout 600 python bench.py --cli > $D/bench_c2.json 2> $D/bench_c2.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $D/bench_reference_arm.json 2> $D/bench_reference_arm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:kcf_ -c 400 --csv --log-file $D/launches_c2.csv python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $D/ncu_launch.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:kcf_screen_kernel -s 3 -c 1 --csv --log-file $D/traffic_c2.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > $D/ncu_traffic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kcf_screen_kernel -s 3 -c 1 -o $D/prof_screen_c2 -f python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > $D/ncu_full.log 2>&1
timeout 300 python bench.py --workload c3s --no-cpu-baseline --steps 20 > $D/bench_c3s.json 2> $D/bench_c3s.err
timeout 300 python bench.py --workload c3st --no-cpu-baseline --steps 20 > $D/bench_c3st.json 2> $D/bench_c3st.err
timeout 600 python tools/pipeline_c5s.py --samples 4 > $D/pipeline_c5s.json 2> $D/pipeline_c5s.err
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kcf_tools.py tests/test_gpu_parity.py -m gpu -x -q -k "device_results or scan_path or fixed_windows" > $D/sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" >> $D/sanitizer_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_fixture_small.py -m gpu -x -q > $D/sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" >> $D/sanitizer_racecheck.txt
tail -3 $D/tests.log; cut -c1-300 $D/bench_c2.json; cat $D/pipeline_c5s.json; tail -4 $D/sanitizer_memcheck.txt $D/sanitizer_racecheck.txt
