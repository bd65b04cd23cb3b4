#!/bin/bash
# ncu --set full of the stand-alone CTA-pair GEMM at the shapes round 1 recorded (profiles/r01_ncu_full_gemm_*):
# qkv at batch 256 (4864 x 2304 x 768) and fc1 at batch 2048 (38912 x 3072 x 768), plain epilogue, cold cache.
# usage (under gpurun): tools/ncu_gemm_shapes.sh <out-prefix>
set -e
out=${1:-gpurun_out/r2_gemm_shapes}
cat > /tmp/ncu_gemm_shapes.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
sys.argv = ["gemm_bench.py"]
import torch
from tools import gemm_bench as G
for (M, N, K) in ((4864, 2304, 768), (38912, 3072, 768), (4864, 768, 3072)):
    print(M, N, K, G.time_gemm(M, N, K, 0, 3) * 1e6)
PY
python /tmp/ncu_gemm_shapes.py > ${out}_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -o ${out} -f \
    python /tmp/ncu_gemm_shapes.py > ${out}_ncu.log 2>&1
ncu -i ${out}.ncu-rep --page raw --csv > ${out}_raw.csv 2>/dev/null || true
