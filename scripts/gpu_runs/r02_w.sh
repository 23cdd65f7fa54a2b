#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 120 python scripts/block_bulk_sweep.py hubbard4x3 1 10 12 16 17 18 > gpurun_out/r02w_sweep_4x3.txt 2>&1; echo "4x3 rc=$?"; tail -6 gpurun_out/r02w_sweep_4x3.txt
timeout -k 5 600 python scripts/block_bulk_sweep.py hubbard4x4 1 10 11 12 16 17 18 > gpurun_out/r02w_sweep_4x4.txt 2>&1; echo "4x4 rc=$?"; tail -8 gpurun_out/r02w_sweep_4x4.txt
