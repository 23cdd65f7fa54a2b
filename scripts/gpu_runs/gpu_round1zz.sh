#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 120 python scripts/terms_check.py > gpurun_out/terms_check.txt 2>&1; echo "terms_check rc=$?"; tail -22 gpurun_out/terms_check.txt
timeout -k 5 100 python scripts/kbench.py hubbard4x4 --matfree --terms --ids 0 > gpurun_out/kbench_terms_c.txt 2>&1; tail -4 gpurun_out/kbench_terms_c.txt
timeout -k 5 100 python scripts/kbench.py hubbard4x4 --matfree --terms --ids 0 --real > gpurun_out/kbench_terms_r.txt 2>&1; tail -3 gpurun_out/kbench_terms_r.txt
