#!/bin/bash
# 2-GPU sanity of the final code: data-parallel check, bench in both DP modes, the notebook-II workload data-parallel
mkdir -p gpurun_out
OUT=gpurun_out/final_n2.txt
: > $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_check.py 2>&1 | grep -E "DP CHECK|peer-memory" | tee -a $OUT
run() { label=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 "$@" 2>gpurun_out/final_n2.err | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$label', round(d['ms_per_step'],3), 'ms/step', round(d['value']), 'samples/s e2e', round(d['e2e']['value']), d['config'].get('dp_step'), d['config']['workload'], flush=True)" | tee -a $OUT
  tail -2 gpurun_out/final_n2.err | grep -i "error" | tee -a $OUT
}
run "n2 package p2p"
run "n2 slots p2p" --workload slots
run "n2 slots nccl" --workload slots --dp-mode nccl
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-400 | tee -a $OUT
timeout 300 python bench.py --workload slots --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('n1 slots', round(d['ms_per_step'],3), 'ms/step', round(d['value']), 'samples/s e2e', round(d['e2e']['value']), 'gemm frac', round(d['roofline']['frac'],3), flush=True)" | tee -a $OUT
