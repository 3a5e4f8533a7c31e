#!/bin/bash
# multi-GPU session: the bench exactly as the driver launches it (torchrun, one rank per GPU, NCCL), both arms
N=${1:-2}; TAG=${2:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== gpu arm, N=$N"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 2>$OUT/bench_${TAG}_n$N.err | tee $OUT/bench_${TAG}_n$N.json | python -c "
import sys,json
t=sys.stdin.read().strip().split('\n')[-1]
d=json.loads(t)
print('value', round(d['value']), 'ms', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), 'n_gpus', d['n_gpus'])
for k in ('capacity_search','move_scoring'):
    b=d.get(k); print(k, json.dumps(b)[:900])
"
tail -3 $OUT/bench_${TAG}_n$N.err
echo "== reference arm, N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
