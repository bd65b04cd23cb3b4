#!/bin/bash
# Run the GPU parity suites group by group, each in its own process (a faulting kernel poisons only its group),
# logs under gpurun_out/. Usage: bash tests/gpu_run_groups.sh [group ...]
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
declare -A G
G[ops_basic]="tests/test_gpu_ops.py -k 'simt or layernorm or attention or cross_entropy or premodules or adamw'"
G[ops_tc]="tests/test_gpu_ops.py -k 'tcgen05'"
G[ops_wgrad]="tests/test_gpu_ops.py -k 'wgrad'"
G[models_fp32]="tests/test_gpu_models.py -k 'fp32 or fallback or roundtrip or (graphed and not dropout) or fused'"
G[models_bf16]="tests/test_gpu_models.py -k 'bf16'"
G[data_path]="tests/test_gpu_data_path.py"
groups="$@"
[ -z "$groups" ] && groups="ops_basic ops_tc ops_wgrad models_fp32 models_bf16 data_path"
for g in $groups; do
  echo "=== $g"
  eval timeout 900 python -m pytest ${G[$g]} -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/$g.log 2>&1
  echo "exit $?"
  grep -E "passed|failed|error" gpurun_out/$g.log | tail -3
  grep -E "^(FAILED|ERROR)" gpurun_out/$g.log | head -40
done
