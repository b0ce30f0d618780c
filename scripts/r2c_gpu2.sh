#!/bin/bash
# round 2, call c (2 GPUs): GPU suite, 2-rank parity (fused in-kernel exchange / NCCL two-pass), default bench at N=1 and N=2, ncu of the default path
mkdir -p gpurun_out
export FVS2D_DEBUG=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
BIG=1 timeout 400 $TR --master-port 29551 scripts/mgpu_parity.py > gpurun_out/r2c_mgpu_fused.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_mgpu_fused.log
unset FVS2D_DEBUG
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench1.json 2> gpurun_out/r2c_bench1.err; echo "rc=$?" >> gpurun_out/r2c_bench1.err
timeout 600 $TR --master-port 29553 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c_bench2.json 2> gpurun_out/r2c_bench2.err; echo "rc=$?" >> gpurun_out/r2c_bench2.err
# ncu: launch list of one bench step sequence, then --set full of the two fused launches on the C4 slice
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-e2e --no-parity --sustain-s 0 > gpurun_out/r2c_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage_fused -s 8 -c 4 -o gpurun_out/r2c_fused_c4 -f python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-extra --no-e2e --no-parity --sustain-s 0 > gpurun_out/r2c_ncu_c4.log 2>&1
tail -4 gpurun_out/r2c_tests.log; grep -h "ranks\|PARITY\|rc=" gpurun_out/r2c_mgpu_fused.log
for f in gpurun_out/r2c_bench1.json gpurun_out/r2c_bench2.json; do python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value']/1e9, d['ms_per_step'], d['gpu_launches'], d['config'].get('setup_s'), d['parity'], d['state_check'], d['sustained'])
except Exception as e: print('$f', 'unreadable', e)
"; done
