set -x
D=gpurun_out/r2l; mkdir -p $D
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "partitioned or sharded or kmers_longer" > $D/tests.log 2>&1; echo "tests rc=$?" >> $D/tests.log; tail -15 $D/tests.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $D/bench_n2.json 2> $D/bench_n2.err; echo "n2 rc=$?"; tail -12 $D/bench_n2.err | cut -c1-300
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2l/bench_n2.json'))
print('value',j['value'],'e2e',j['e2e']['value'],j['e2e']['ms_per_step'])
print(json.dumps(j.get('placements'),indent=1)[:5000])
PY
