#!/bin/bash
# round 2, final 1-GPU evidence run on the committed code: full GPU suite, smoke, default bench (as the driver runs it),
# reference arm (short), ncu launch list + --set full of the stage kernels (C4 slice and C3)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2s_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2s_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2s_bench1.json 2> gpurun_out/r2s_bench1.err; echo "rc=$?" >> gpurun_out/r2s_bench1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2s_ref.json 2> gpurun_out/r2s_ref.err; echo "rc=$?" >> gpurun_out/r2s_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-e2e --no-parity --sustain-s 0 > gpurun_out/r2s_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_stage_fused -s 8 -c 2 -o gpurun_out/r2s_fused_c4 -f python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-extra --no-e2e --no-parity --sustain-s 0 > gpurun_out/r2s_ncu_c4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_gradient2|k_flux_rk2" -s 8 -c 2 -o gpurun_out/r2s_pair_c2 -f python bench.py --workload naca --steps 3 --warmup 1 --no-cpu-baseline --no-extra --no-e2e --no-parity --sustain-s 0 > gpurun_out/r2s_ncu_c2.log 2>&1
tail -3 gpurun_out/r2s_tests.log; tail -2 gpurun_out/r2s_smoke.log
for f in r2s_bench1 r2s_ref; do python -c "
import json
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', d['value']/1e9, d['ms_per_step'], d.get('gpu_launches'), d['config'].get('setup_s'), d.get('roofline',{}).get('frac'), (d.get('sustained') or {}).get('value'), (d.get('e2e') or {}).get('value'))
except Exception as e: print('$f unreadable', e)
"; done
