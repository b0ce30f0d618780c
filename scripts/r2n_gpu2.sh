#!/bin/bash
# 2 GPUs: full GPU suite (incl. test_gpu_multi), 2-rank parity incl. the output path, both exchange paths; N=2 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2n_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
BIG=1 timeout 400 $TR --master-port 29551 scripts/mgpu_parity.py > gpurun_out/r2n_mgpu_fused.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_mgpu_fused.log
FUSE=0 timeout 400 $TR --master-port 29552 scripts/mgpu_parity.py > gpurun_out/r2n_mgpu_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_mgpu_nccl.log
timeout 400 $TR --master-port 29553 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2n_bench2.json 2> gpurun_out/r2n_bench2.err; echo "rc=$?" >> gpurun_out/r2n_bench2.err
tail -3 gpurun_out/r2n_tests.log; grep -h "ranks\|PARITY\|rc=\|Error" gpurun_out/r2n_mgpu_fused.log gpurun_out/r2n_mgpu_nccl.log | cut -c1-330; tail -1 gpurun_out/r2n_bench2.err
python -c "
import json
d=json.loads(open('gpurun_out/r2n_bench2.json').read().strip().splitlines()[-1]); print('N=2', d['value']/1e9, d['ms_per_step'], d['gpu_launches'], d['config']['setup_s'], d['parity']['ok'], d['sustained']['value']/1e9, d['e2e']['value']/1e9)
"
