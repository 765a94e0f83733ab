set -x
D=gpurun_out/r2d; mkdir -p $D
nvidia-smi -L | head -3; free -g | head -2; nproc
timeout 300 python bench.py --only resident,e2e,cold --steps 10 > $D/bench_n1_quick.json 2> $D/bench_n1_quick.err; tail -3 $D/bench_n1_quick.err
python -c "
import json;j=json.load(open('$D/bench_n1_quick.json'));print('value',j['value'],'e2e',j['e2e']['ms_per_step'],j['e2e']['h2d_copy_alone_ms'],'cold',json.dumps(j['e2e_cold']['runs']))"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $D/bench_n2.json 2> $D/bench_n2.err; echo "n2 rc=$?"; tail -40 $D/bench_n2.err | cut -c1-300
cut -c1-500 $D/bench_n2.json
