#!/bin/bash
# ncu captures of the fused stage kernel (round 2): scripts/profile_fused.sh [c3|c4] [fuse variant]
# One GPU; the .ncu-rep lands in gpurun_out/ and is summarised with scripts/ncu_summary.py.
W=${1:-c3}; V=${2:-2}
mkdir -p gpurun_out
FUSE=$V timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_stage_fused -s 8 -c 2 \
  -o gpurun_out/prof_fused${V}_${W} -f python scripts/fused_check.py time $W > gpurun_out/prof_fused${V}_${W}.log 2>&1
echo "rc=$?" >> gpurun_out/prof_fused${V}_${W}.log
