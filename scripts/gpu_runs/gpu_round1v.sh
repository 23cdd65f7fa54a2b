#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 200 python scripts/kbench.py hubbard4x4 --ids 0 > gpurun_out/kbench_prod_c.txt 2>&1; tail -2 gpurun_out/kbench_prod_c.txt
timeout -k 5 200 python scripts/kbench.py hubbard4x4 --ids 0 --real > gpurun_out/kbench_prod_r.txt 2>&1; tail -2 gpurun_out/kbench_prod_r.txt
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider -k "block or shard" > gpurun_out/pytest_blocks.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_blocks.log
