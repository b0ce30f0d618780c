#!/bin/bash
# round 2 re-entry check: the whole GPU suite (with the 4 000-step vortex example), smoke, default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2q_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2q_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench1.json 2> gpurun_out/r2q_bench1.err; echo "rc=$?" >> gpurun_out/r2q_bench1.err
tail -3 gpurun_out/r2q_tests.log; tail -2 gpurun_out/r2q_smoke.log; tail -c 600 gpurun_out/r2q_bench1.json
