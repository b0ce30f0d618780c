#!/bin/bash
# Round-1 closing GPU call: full GPU suite, smoke, the default bench line, and the ncu captures of the final build
# (launch list + --set full of pass B on the C4 slice and on C3).  Logs go to gpurun_out/ progressively.
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/e_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/e_rc.txt
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/e_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/e_rc.txt
timeout 240 python bench.py > gpurun_out/bench_r1e.json 2> gpurun_out/e_bench.err; echo "bench rc=$?" >> gpurun_out/e_rc.txt
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra --opt graph=0"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1e.csv $B > gpurun_out/launches_r1e.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k "regex:k_flux_pipe" -s 6 -c 1 -f -o gpurun_out/prof_pipe_r1e $B > gpurun_out/prof_pipe_r1e.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k "regex:k_flux_pipe" -s 6 -c 1 -f -o gpurun_out/prof_pipe_c3_r1e $B --workload c3 > gpurun_out/prof_pipe_c3_r1e.log 2>&1
cat gpurun_out/e_rc.txt; tail -n 2 gpurun_out/e_tests.log; tail -n 1 gpurun_out/e_smoke.log; ls -la gpurun_out/*r1e*
