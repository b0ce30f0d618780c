#!/bin/bash
# A/B of several library builds on one box: scripts/ab_list.sh <workload> <lib.so>...   (libs relative to fvs2d_b200/csrc)
# One short bench pass per library in the order given (repeat a name to bracket drift); one line per run into gpurun_out/ab.txt
W=$1; shift
mkdir -p gpurun_out
for lib in "$@"; do
  FVS2D_GPU_LIB=$PWD/fvs2d_b200/csrc/$lib timeout 120 python bench.py --workload $W --no-cpu-baseline --no-e2e --no-extra 2>gpurun_out/ab_err.txt | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$W', '$lib'.ljust(28), round(d['value']/1e9,3), 'G/s', round(d['ms_per_step'],4), 'ms/step  B', round(d['roofline']['avg_launch_ms'],4), ' A', round(d['gradient_kernel']['avg_launch_ms'],4), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e:
    print('$W', '$lib', 'FAILED', e)" | tee -a gpurun_out/ab.txt
done
