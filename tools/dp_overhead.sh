#!/bin/bash
# Where the data-parallel overhead comes from (run with gpurun --gpus 2): overlap on/off, bucket size, NCCL CTA budget
mkdir -p gpurun_out
run() { # label, env..., -- bench args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 3 "$@" 2>/dev/null | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$label', round(d['ms_per_step'],3), 'ms/step', round(d['value']), 'samples/s')"
}
run "overlap(default)" X=1 --
run "overlap margin8" VY_GEMM_SM_MARGIN=8 --
run "overlap margin16" VY_GEMM_SM_MARGIN=16 --
run "overlap margin16 CTAS16" VY_GEMM_SM_MARGIN=16 NCCL_MAX_CTAS=16 --
