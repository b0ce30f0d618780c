#!/bin/bash
# A/B of a stage-kernel change: bitwise parity of the fused path vs the two-pass path, then C3 / C4 timing (3 passes each)
timeout 200 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -2
FUSE=-1 timeout 300 python scripts/fused_check.py time c3 c4 2>&1 | grep "TIME" | grep -v "fuse 0"
