#!/bin/bash
# per-kernel durations of a short production run (ncu launch list; run on the GPU box): scripts/time_kernels.sh <tag> [bench args]
tag=$1; shift
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-extra --opt graph=0 "$@" > gpurun_out/launches_$tag.log 2>&1
python scripts/ncu_summary.py gpurun_out/launches_$tag.csv | head -16
