#!/bin/bash
# 4 ranks: parity of the in-kernel halo exchange (incl. the 540k mesh), then the default bench
mkdir -p gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" -le 4 ]; then
  BIG=1 FVS2D_DEBUG=1 timeout 400 $TR --master-port 29561 scripts/mgpu_parity.py > gpurun_out/r2e_mgpu_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_mgpu_n$N.log
  grep -h "ranks\|PARITY\|rc=" gpurun_out/r2e_mgpu_n$N.log
fi
timeout 900 $TR --master-port 29563 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2e_bench_n$N.json 2> gpurun_out/r2e_bench_n$N.err; echo "rc=$?" >> gpurun_out/r2e_bench_n$N.err
tail -3 gpurun_out/r2e_bench_n$N.err
python -c "
import json
d=json.loads(open('gpurun_out/r2e_bench_n$N.json').read().strip().splitlines()[-1]); print('N=$N', d['value']/1e9, d['ms_per_step'], d['gpu_launches'], d['config'].get('setup_s'), d['parity'], d['state_check'], d['sustained'], d['e2e'])
"
