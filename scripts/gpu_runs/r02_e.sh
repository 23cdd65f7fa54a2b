#!/bin/bash
# Round 2, fifth GPU call: ncu of the fp64 two-pass product (block-local + bulk-streamed kernels) and a sweep with lean rings.
mkdir -p gpurun_out
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:'spmv_sjds|sjds_block' -s 2 -c 2 -o gpurun_out/r02e_prof_species_stored_fp64 \
   python scripts/species_ncu_target.py hubbard4x4 stored 2 real > gpurun_out/r02e_ncu_stored_fp64.log 2>&1; echo "ncu stored fp64 rc=$?"; tail -2 gpurun_out/r02e_ncu_stored_fp64.log
QBGPU_VERBOSE=1 timeout -k 5 900 python scripts/bulk_sweep.py hubbard4x4 > gpurun_out/r02e_bulk_sweep.txt 2>&1; grep -v "autotune\|bulk kernel NW" gpurun_out/r02e_bulk_sweep.txt
