#!/bin/bash
# ncu --set full over a window of the train step's own kernels (block 11 forward .. block 11 backward), one GPU.
# usage (under gpurun): tools/ncu_step_window.sh <out-prefix>
# The same command runs first WITHOUT ncu (B200_PROFILING.md); launches are host-launched (--no-graph) so that ncu
# sees kernel nodes in program order.
set -e
out=${1:-gpurun_out/r2_step_window}
cmd="python bench.py --steps 1 --warmup 3 --no-graph --profile-only"
$cmd > ${out}_plain.log 2>&1
# kernels per host-launched step: ~242; 169 matching launches per step: skip 3 warm-up steps + block 0..10 of the forward (1 + 11 x 6)
ncu --set full --clock-control none --import-source on \
    -k regex:'gemm_tc2_kernel|attn_tc|adapter_kernel|ln_bwd_pair|ln_fwd' -s 574 -c 16 \
    -o ${out} -f $cmd > ${out}_ncu.log 2>&1
ncu -i ${out}.ncu-rep --page raw --csv > ${out}_raw.csv 2>/dev/null || true
