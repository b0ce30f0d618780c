#!/bin/bash
# programmatic dependent launch: GPU suite, then timing with / without on C2, C1 and the C4 slice
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { timeout 150 python bench.py --workload $1 --steps $2 --warmup 20 --no-cpu-baseline --no-e2e --sustain-s 0 --no-parity --no-extra "${@:3}" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(d['value']/1e9,3), 'G/s', round(d['ms_per_step']*1e3,2), 'us/step launches', d['gpu_launches'])
"; }
run naca 2000 --opt pdl=1
run naca 2000 --opt pdl=0
run vortex 2000 --opt pdl=1
run vortex 2000 --opt pdl=0
run c4 20 --opt pdl=1
run c4 20 --opt pdl=0
run c3 20 --opt pdl=1
run c3 20 --opt pdl=0
