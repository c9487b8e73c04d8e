#!/bin/bash
# ncu --set full of the attention kernels after the round-2 changes: tcgen05 forward (decoder + ViT shapes), mma.sync forward (C5 shapes)
mkdir -p gpurun_out
CMD="python tools/attn_bench.py"
$CMD > gpurun_out/plain_attn2.log 2>&1; tail -8 gpurun_out/plain_attn2.log
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 3 -c 1 -o gpurun_out/r02_prof_attnfwd_dec_v2 $CMD > gpurun_out/ncu_attn_a.log 2>&1; echo "ncu a rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 30 -c 1 -o gpurun_out/r02_prof_attnfwd_vit_v2 $CMD > gpurun_out/ncu_attn_b.log 2>&1; echo "ncu b rc=$?"
CMD2="python tools/c5_attn_probe.py"
$CMD2 > gpurun_out/plain_c5attn.log 2>&1; tail -4 gpurun_out/plain_c5attn.log
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_mma_kernel -s 4 -c 2 -o gpurun_out/r02_prof_attn_mma $CMD2 > gpurun_out/ncu_attn_c.log 2>&1; echo "ncu c rc=$?"
