#!/bin/bash
# round 2, call b (2 GPUs): GPU suite, 2-rank parity of the in-kernel halo exchange (fused) and of the NCCL two-pass path, N=1 / N=2 bench
mkdir -p gpurun_out
export FVS2D_DEBUG=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2b_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
BIG=1 timeout 400 $TR --master-port 29551 scripts/mgpu_parity.py > gpurun_out/r2b_mgpu_fused.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_mgpu_fused.log
FUSE=0 timeout 400 $TR --master-port 29552 scripts/mgpu_parity.py > gpurun_out/r2b_mgpu_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_mgpu_nccl.log
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2b_bench1.json 2> gpurun_out/r2b_bench1.err; echo "rc=$?" >> gpurun_out/r2b_bench1.err
timeout 500 $TR --master-port 29553 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench2.json 2> gpurun_out/r2b_bench2.err; echo "rc=$?" >> gpurun_out/r2b_bench2.err
timeout 500 $TR --master-port 29554 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --opt fuse=0 > gpurun_out/r2b_bench2_nccl.json 2> gpurun_out/r2b_bench2_nccl.err; echo "rc=$?" >> gpurun_out/r2b_bench2_nccl.err
tail -4 gpurun_out/r2b_tests.log; grep -h "ranks\|PARITY\|rc=" gpurun_out/r2b_mgpu_fused.log gpurun_out/r2b_mgpu_nccl.log
for f in gpurun_out/r2b_bench1.json gpurun_out/r2b_bench2.json gpurun_out/r2b_bench2_nccl.json; do python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value']/1e9, d['ms_per_step'], d['gpu_launches'], d['config'].get('setup_s'))
except Exception as e: print('$f', 'unreadable', e)
"; done
