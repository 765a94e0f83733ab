#!/usr/bin/env bash
# The commands behind profiles/r2* (one B200 unless stated; run through gpurun, outputs under gpurun_out/).  Numbers printed
# by a run under ncu are never bench values.
set -x
D=gpurun_out/round2; mkdir -p $D
# parity: every GPU test through the C ABI, the smoke check, compute-sanitizer over the newest paths
timeout 1200 python -m pytest tests -m gpu -x -q > $D/tests.log 2>&1; echo "tests rc=$?" >> $D/tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $D/smoke.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
    -k "sharded_job_over or kmers_longer or exchange_path or fixed_windows or random_window" > $D/sanitizer_memcheck.txt 2>&1
# the two bench lines the driver takes
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $D/bench_c2.json 2> $D/bench_c2.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $D/bench_reference_arm.json 2> $D/bench_reference_arm.err
# ncu: launch list (shares), DRAM bytes of the screening kernel, one full capture of it
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:kcf_(screen|finalize|pack|plan|tile)" -c 400 --csv --log-file $D/launches_c2.csv \
    python bench.py --only resident,e2e --steps 3 --warmup 3 --e2e-steps 1 > $D/ncu_launch.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:kcf_screen_kernel -s 3 -c 1 --csv \
    --log-file $D/traffic_c2.csv python bench.py --only resident --steps 2 --warmup 3 > $D/ncu_traffic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kcf_screen_kernel -s 3 -c 1 -o $D/prof_screen_c2 -f \
    python bench.py --only resident --steps 2 --warmup 3 > $D/ncu_full.log 2>&1
# table density sweep
for lf in 0.15 0.3 0.5 0.7 0.9; do timeout 300 python bench.py --only resident --lf $lf --steps 10 > $D/density_$lf.json 2> $D/density_$lf.err; done
# N GPUs (gpurun --gpus N): ONE job, strong scaling, + the placements of the 3e9-record table
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus N --steps 20 --warmup 5
# configs[3] at full size (gpurun --gpus 8):
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/c4_full.py c4
