#!/bin/bash
# small meshes: two threads per cell (pair kernels) vs one thread per cell, C2 (NACA) and C1 (vortex example)
timeout 300 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4
run() { timeout 120 python bench.py --workload $1 --steps 2000 --warmup 100 --no-cpu-baseline --no-e2e --sustain-s 0 --no-parity "${@:2}" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(d['value']/1e9,3), 'G/s', round(d['ms_per_step']*1e3,2), 'us/step launches', d['gpu_launches'], 'grad', round(d['gradient_kernel']['avg_launch_ms']*1e3,2), 'flux', round(d['roofline']['avg_launch_ms']*1e3,2))
"; }
run naca --opt pair=1
run naca --opt pair=0
run naca --opt pair=0 --opt tile=0
run vortex
