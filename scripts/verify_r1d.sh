#!/bin/bash
# One-call GPU verification of the output-path entry points (round 1, last GPU minutes): new tests first, every
# step under its own timeout, logs written progressively into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/v_gpu.txt 2>&1
timeout 170 python -m pytest tests/test_gpu_output.py -x -q > gpurun_out/v_output.log 2>&1; echo "output rc=$?" >> gpurun_out/v_rc.txt
timeout 170 python -m pytest tests/test_gpu_driver.py -x -q > gpurun_out/v_driver.log 2>&1; echo "driver rc=$?" >> gpurun_out/v_rc.txt
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/v_rc.txt
timeout 170 python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/v_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/v_rc.txt
timeout 400 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_output.py --deselect tests/test_gpu_driver.py > gpurun_out/v_all.log 2>&1; echo "all rc=$?" >> gpurun_out/v_rc.txt
cat gpurun_out/v_rc.txt; tail -3 gpurun_out/v_output.log gpurun_out/v_driver.log gpurun_out/v_smoke.log gpurun_out/v_all.log; tail -c 600 gpurun_out/v_bench.log
