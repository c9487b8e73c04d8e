#!/bin/bash
# Where the data-parallel overhead comes from (run with gpurun --gpus N, N = 2 by default): overlap on/off, SM margin of the
# persistent GEMM grids, NCCL CTA budget. Writes gpurun_out/dp_overhead_n$N.txt.
N=${1:-2}
mkdir -p gpurun_out
OUT=gpurun_out/dp_overhead_n$N.txt
: > $OUT
run() { # label, env..., -- bench args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 --no-decode "$@" 2>/dev/null | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$label', round(d['ms_per_step'],3), 'ms/step', round(d['value']), 'samples/s', flush=True)" | tee -a $OUT
}
run "n$N overlap(default)" X=1 --
run "n$N overlap margin8" VY_GEMM_SM_MARGIN=8 --
run "n$N overlap margin16" VY_GEMM_SM_MARGIN=16 --
run "n$N overlap margin8 CTAS8" VY_GEMM_SM_MARGIN=8 NCCL_MAX_CTAS=8 --
run "n$N overlap margin16 CTAS16" VY_GEMM_SM_MARGIN=16 NCCL_MAX_CTAS=16 --
run "n$N overlap margin4 CTAS4" VY_GEMM_SM_MARGIN=4 NCCL_MAX_CTAS=4 --
run "n$N no-overlap" X=1 -- --no-overlap
cat $OUT
