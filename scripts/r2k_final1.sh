#!/bin/bash
# round 2, 1-GPU evidence run: full GPU suite, smoke, default bench (as the driver runs it), reference arm (short), the full
# 69.12 M-cell C4 mesh on one GPU, ncu launch list + --set full of the stage kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2k_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench1.json 2> gpurun_out/r2k_bench1.err; echo "rc=$?" >> gpurun_out/r2k_bench1.err
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2k_ref.json 2> gpurun_out/r2k_ref.err; echo "rc=$?" >> gpurun_out/r2k_ref.err
FVS2D_DEBUG=1 timeout 500 python bench.py --mesh-for-gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-parity --no-e2e > gpurun_out/r2k_c4full.json 2> gpurun_out/r2k_c4full.err; echo "rc=$?" >> gpurun_out/r2k_c4full.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-e2e --no-parity --sustain-s 0 > gpurun_out/r2k_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_stage_fused -s 8 -c 4 -o gpurun_out/r2k_fused_c4 -f python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-extra --no-e2e --no-parity --sustain-s 0 > gpurun_out/r2k_ncu_c4.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_stage_fused -s 4 -c 2 -o gpurun_out/r2k_fused_c3 -f python bench.py --workload c3 --steps 3 --warmup 1 --no-cpu-baseline --no-extra --no-e2e --no-parity --sustain-s 0 > gpurun_out/r2k_ncu_c3.log 2>&1
tail -3 gpurun_out/r2k_tests.log; tail -2 gpurun_out/r2k_smoke.log
for f in r2k_bench1 r2k_ref r2k_c4full; do python -c "
import json
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', d['value']/1e9, d['ms_per_step'], d.get('gpu_launches'), d['config'].get('setup_s'), d.get('roofline',{}).get('frac'), (d.get('sustained') or {}).get('value'), (d.get('e2e') or {}).get('value'))
except Exception as e: print('$f unreadable', e)
"; done
