#!/bin/bash
# A/B of two library builds on the same box, alternating: scripts/ab.sh <libA.so> <libB.so> [workloads...]
A=$1; B=$2; shift 2
for w in "${@:-c4}"; do
  for rep in 1 2; do
    for lib in $A $B; do
      FVS2D_GPU_LIB=$PWD/$lib timeout 200 python bench.py --workload $w --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', '$lib'.split('/')[-1].ljust(24), round(d['value']/1e9,3), 'G/s', round(d['ms_per_step'],4), 'ms/step  B', round(d['roofline']['avg_launch_ms'],4), ' A', round(d['gradient_kernel']['avg_launch_ms'],4))"
    done
  done
done
