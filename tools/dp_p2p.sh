#!/bin/bash
# The peer-memory sharded optimizer step against the NCCL all-reduce path (run with gpurun --gpus N, N = 2 by default):
# correctness first (tools/dp_check.py at 2 ranks), then bench.py in both modes. Writes gpurun_out/dp_p2p_n$N.txt.
N=${1:-2}
mkdir -p gpurun_out
OUT=gpurun_out/dp_p2p_n$N.txt
: > $OUT
timeout 300 python -m pytest tests/test_gpu_dp_shard.py -x -q 2>&1 | tail -5 | tee -a $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tee -a $OUT
run() { # label, -- bench args
  label=$1; shift; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 --no-decode "$@" 2>gpurun_out/dp_p2p_err.log | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$label', round(d['ms_per_step'],3), 'ms/step', round(d['value']), 'samples/s e2e', round(d['e2e']['value']), d['config'].get('dp_step'), flush=True)" | tee -a $OUT
  tail -3 gpurun_out/dp_p2p_err.log | grep -i "error\|Traceback" | tee -a $OUT
}
timeout 300 python bench.py --steps 20 --warmup 3 --no-decode --no-cpu-baseline 2>/dev/null | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('n1', round(d['ms_per_step'],3), 'ms/step', round(d['value']), 'samples/s', flush=True)" | tee -a $OUT
run "n$N p2p" -- --dp-mode p2p
run "n$N nccl" -- --dp-mode nccl
run "n$N p2p (again)" -- --dp-mode p2p
cat $OUT
