mkdir -p gpurun_out/r1e
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
timeout 400 $T bench.py --gpus 8 --workload c2 --steps 10 --warmup 3 --placement partitioned-scan --table-parts 2 --no-cpu-baseline > gpurun_out/r1e/scan_c2_n8_t2.json 2> gpurun_out/r1e/scan_c2_n8_t2.err
timeout 400 $T bench.py --gpus 8 --workload c2 --steps 10 --warmup 3 --placement partitioned-scan --table-parts 4 --no-cpu-baseline > gpurun_out/r1e/scan_c2_n8_t4.json 2> gpurun_out/r1e/scan_c2_n8_t4.err
timeout 400 $T bench.py --gpus 8 --workload c2 --steps 10 --warmup 3 --e2e-steps 2 --no-cpu-baseline > gpurun_out/r1e/repl_c2_n8.json 2> gpurun_out/r1e/repl_c2_n8.err
tail -n 2 gpurun_out/r1e/*n8*.err; cat gpurun_out/r1e/*n8*.json | cut -c1-260
