#!/bin/bash
# Last seconds of the round's GPU budget: the second version of the term-coded replay (correctness, then one timing).
mkdir -p gpurun_out
export QBGPU_TERMS_KERNEL=2
timeout -k 2 12 python scripts/terms_check.py > gpurun_out/terms_check_v2.txt 2>&1; echo "terms_check v2 rc=$?"; tail -20 gpurun_out/terms_check_v2.txt
timeout -k 2 12 python scripts/kbench.py hubbard4x4 --matfree --terms --ids 0 > gpurun_out/kbench_terms_v2_c.txt 2>&1; tail -3 gpurun_out/kbench_terms_v2_c.txt
