#!/bin/bash
# Round-2 checklist for the fused stage kernels (run on a B200 box; everything logs into gpurun_out/).
#   scripts/round2_fused.sh parity      all variants against the two-pass path (bitwise), ~10 s
#   scripts/round2_fused.sh time        ms per stage on C3 and the C4 slice for fuse = 2, 4, 5, ~40 s
#   scripts/round2_fused.sh stages3     the same with a 3-stage ring (alternative build, C3 + C4), ~2 min incl. the build
#   scripts/round2_fused.sh mgpu        2-rank parity of the fused path (gpurun --gpus 2), ~1 min
set -u
mkdir -p gpurun_out
case "${1:-parity}" in
  parity)  FUSE=1,2,3,4,5 FVS2D_DEBUG=1 timeout 120 python scripts/fused_check.py parity 2>&1 | tee gpurun_out/r2_fused_parity.log ;;
  time)    FUSE=2,4,5 FVS2D_DEBUG=1 timeout 200 python scripts/fused_check.py time c3 c4 2>&1 | tee gpurun_out/r2_fused_time.log ;;
  stages3) make -C fvs2d_b200/csrc STAGES=3 SUFFIX=_s3 libfvs2d_gpu_s3.so > gpurun_out/r2_build_s3.log 2>&1
           FVS2D_GPU_LIB=$PWD/fvs2d_b200/csrc/libfvs2d_gpu_s3.so FUSE=2,5 FVS2D_DEBUG=1 timeout 200 python scripts/fused_check.py time c3 c4 2>&1 | tee gpurun_out/r2_fused_time_s3.log ;;
  mgpu)    for f in 2 5; do FUSE=$f timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 scripts/mgpu_parity.py 2>&1 | tail -6 | tee -a gpurun_out/r2_fused_mgpu.log; done ;;
esac
