set -x
D=gpurun_out/r2r; mkdir -p $D
timeout 1200 python -m pytest tests -m gpu -x -q > $D/tests.log 2>&1; echo "tests rc=$?" >> $D/tests.log; tail -4 $D/tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $D/smoke.log 2>&1; tail -2 $D/smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $D/bench_c2.json 2> $D/bench_c2.err; echo "bench rc=$?"; tail -12 $D/bench_c2.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $D/bench_reference_arm.json 2> $D/bench_reference_arm.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:kcf_ -c 1500 --csv --log-file $D/launches_c2.csv python bench.py --only resident,e2e,cold --steps 3 --warmup 3 --e2e-steps 1 > $D/ncu_launch.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sharded_job_over or kmers_longer or exchange_path or fixed_windows or random_window" > $D/sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" >> $D/sanitizer_memcheck.txt; tail -3 $D/sanitizer_memcheck.txt
cut -c1-300 $D/bench_c2.json
