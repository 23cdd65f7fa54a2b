#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:'sjds_block_bulk' -s 2 -c 1 -o gpurun_out/r02v_prof_block_bulk python scripts/block_bulk_sweep.py hubbard4x4 12 > gpurun_out/r02v_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02v_ncu.log
