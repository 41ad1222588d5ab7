#!/bin/bash
# Rebuild the fused kernel with different ring depths on the GPU box and time it (tools/bench_gemm.py).
cd "$(dirname "$0")/.."
for cfg in "4 4" "3 4" "4 3" "3 6" "3 5"; do
  set -- $cfg
  OVDET_NVCC_DEFS="-DOVDET_F_A_STAGES=$1 -DOVDET_F_B_STAGES=$2" python -c "from ovdet import build; build.build(force=True)" > /dev/null 2>&1 || { echo "A=$1 B=$2 build failed"; continue; }
  for i in 1 2; do
    echo "A_STAGES=$1 B_STAGES=$2: $(python tools/bench_gemm.py --fused --batch 256 --iters 60 2>&1 | tail -1)"
  done
  echo "A_STAGES=$1 B_STAGES=$2 projected: $(python tools/bench_gemm.py --projected --batch 256 --iters 60 2>&1 | tail -2 | head -1)"
done
python -c "from ovdet import build; build.build(force=True)" > /dev/null 2>&1
