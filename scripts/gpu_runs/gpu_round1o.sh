#!/bin/bash
# Validate and time the neighbour-walk matrix-free kernel.
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider -k "matrix_free or matfree" > gpurun_out/pytest_matfree.log 2>&1; echo "pytest matfree rc=$?"; tail -5 gpurun_out/pytest_matfree.log
timeout -k 5 200 python scripts/kbench.py hubbard4x4 --matfree --ids 0 > gpurun_out/kbench_matfree_hubbard4x4.txt 2>&1; echo "rc=$?"; tail -6 gpurun_out/kbench_matfree_hubbard4x4.txt
timeout -k 5 200 python scripts/kbench.py heis_chain28 --matfree --ids 0 > gpurun_out/kbench_matfree_heis28.txt 2>&1; echo "rc=$?"; tail -6 gpurun_out/kbench_matfree_heis28.txt
