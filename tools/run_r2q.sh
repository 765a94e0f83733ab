set -x
D=gpurun_out/r2q; mkdir -p $D
free -g | head -2; df -h /dev/shm | tail -1; nproc
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > $D/bench_n8.json 2> $D/bench_n8.err; echo "n8 rc=$?"; tail -30 $D/bench_n8.err | cut -c1-250
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2q/bench_n8.json'))
print('value',j['value'],'ms',j['ms_per_step'],'e2e',j['e2e']['value'],j['e2e']['ms_per_step'],j['e2e']['h2d_copy_alone_ms'], j['checks'])
print(json.dumps(j.get('placements'),indent=1)[:6000])
PY
