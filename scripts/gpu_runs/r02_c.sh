#!/bin/bash
# Round 2, third GPU call: why are the bulk-streamed kernels slow?  ncu --set full of the two passes (new kernels) + mode sweep.
mkdir -p gpurun_out
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:'spmv_sjds|sjds_block' -s 2 -c 2 -o gpurun_out/r02c_prof_species_stored_new \
   python scripts/species_ncu_target.py hubbard4x4 stored 2 > gpurun_out/r02c_ncu_stored.log 2>&1; echo "ncu stored rc=$?"; tail -2 gpurun_out/r02c_ncu_stored.log
timeout -k 5 600 python scripts/bulk_sweep.py hubbard4x4 > gpurun_out/r02c_bulk_sweep.txt 2>&1; cat gpurun_out/r02c_bulk_sweep.txt
