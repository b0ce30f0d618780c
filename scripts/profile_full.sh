#!/bin/bash
# ncu captures for profiles/ (run on the GPU box, one GPU): launch list of a short default bench + --set full of the
# dominant kernels.  Numbers printed by bench.py under ncu are never bench values.
# usage: scripts/profile_full.sh <tag>     -> gpurun_out/launches_<tag>.csv, gpurun_out/prof_{pipe,grad,verr,pipe_c3}_<tag>.ncu-rep
tag=${1:-r1}
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra --opt graph=0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/launches_$tag.log 2>&1
full() {  # name kernel-regex skip count extra-args
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/prof_$1_$tag $B $5 > gpurun_out/prof_$1_$tag.log 2>&1
}
full pipe k_flux_pipe 6 2 ""
full grad k_gradient 6 1 ""
full verr k_vortex_err 1 1 ""
full pipe_c3 k_flux_pipe 6 1 "--workload c3"
ls -la gpurun_out/*_$tag.*
