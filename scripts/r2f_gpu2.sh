#!/bin/bash
# round 2, call f (2 GPUs): GPU suite after the partition-local build, 2-rank parity, N=1 + N=2 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2f_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
BIG=1 FVS2D_DEBUG=1 timeout 300 $TR --master-port 29551 scripts/mgpu_parity.py > gpurun_out/r2f_mgpu_fused.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_mgpu_fused.log
FUSE=0 timeout 300 $TR --master-port 29552 scripts/mgpu_parity.py > gpurun_out/r2f_mgpu_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_mgpu_nccl.log
timeout 400 $TR --master-port 29553 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench2.json 2> gpurun_out/r2f_bench2.err; echo "rc=$?" >> gpurun_out/r2f_bench2.err
tail -4 gpurun_out/r2f_tests.log; grep -h "ranks\|PARITY\|rc=" gpurun_out/r2f_mgpu_fused.log gpurun_out/r2f_mgpu_nccl.log; tail -2 gpurun_out/r2f_bench2.err
python -c "
import json
d=json.loads(open('gpurun_out/r2f_bench2.json').read().strip().splitlines()[-1]); print('N=2', d['value']/1e9, d['ms_per_step'], d['gpu_launches'], d['config'], d['parity']['ok'], d['state_check'], d['sustained']['value']/1e9, d['e2e']['value']/1e9)
"
