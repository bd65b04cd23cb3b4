#!/bin/bash
# Regenerates the one-GPU records under profiles/ (round prefix in $1, default r03) on a B200 box, in the order the
# profiling recipe asks for: plain runs first, anything under ncu only after the same command has exited 0 without it.
#   gpurun --timeout 1500 -- 'bash tools/reproduce_profiles.sh r03'      (about 8 GPU-minutes)
# Multi-GPU records: gpurun --gpus N -- 'bash tools/scale_sweep.sh N' (writes gpurun_out/r2_scaleN_<config>.json).
set -u
R=${1:-r03}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/${R}_gputest.log
cp $O/parity_metrics.jsonl $O/${R}_parity_metrics.jsonl 2>/dev/null
python bench.py > $O/${R}_bench_final.json 2> $O/${R}_bench_final.err
for c in latent_vit image_vit latent_vit_v2; do
  python bench.py --config $c > $O/${R}_bench_$c.json 2> $O/${R}_bench_$c.err
done
python bench.py --impl reference --steps 2 --warmup 1 > $O/${R}_bench_reference_arm.json 2> $O/${R}_bench_reference_arm.err
for c in hybrid latent_vit image_vit latent_vit_v2; do
  python tools/graph_timeline.py --config $c --out $O/${R}_graph_timeline_$c.json 2> /dev/null
done
python tools/small_kernel_bench.py > $O/${R}_small_kernel_bench.jsonl 2>&1
python tools/small_kernel_bench.py --batch 64 --seq 197 --heads 8 --drop 0.1 > $O/${R}_attention_s197_bench.jsonl 2>&1
# ncu launch list of one host-launched step (serialised, cold): the kernel SHARES must agree with the in-graph ones
python bench.py --steps 1 --warmup 3 --no-graph --profile-only > $O/${R}_po.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/${R}_launches_step_final.csv \
      python bench.py --steps 1 --warmup 3 --no-graph --profile-only > $O/${R}_ncu_po.log 2>&1
python tools/ncu_summary.py $O/${R}_launches_step_final.csv > $O/${R}_launches_step_final_summary.txt 2>/dev/null
tail -1 $O/${R}_gputest.log
