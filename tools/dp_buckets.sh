#!/bin/bash
# bucket size of the overlapped gradient all-reduce (run with gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
OUT=gpurun_out/dp_buckets_n$N.txt
: > $OUT
for mb in 8 16 32 64 128; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 --no-decode --bucket-mb $mb 2>/dev/null | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('n$N bucket_mb $mb', round(d['ms_per_step'],3), 'ms/step', round(d['value']), 'samples/s', flush=True)" | tee -a $OUT
done
