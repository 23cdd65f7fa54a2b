#!/bin/bash
# Round 2, fourth GPU call: second version of the bulk-streamed kernels (pipelined consumer, epilogue prefetch, more warps).
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_species.py -q -x -p no:cacheprovider -k "not arpack and not thick_restart" > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02d_pytest.log
grep -E "^E  " gpurun_out/r02d_pytest.log | head -20
QBGPU_VERBOSE=1 timeout -k 5 900 python scripts/bulk_sweep.py hubbard4x4 > gpurun_out/r02d_bulk_sweep.txt 2>&1; grep -v "autotune" gpurun_out/r02d_bulk_sweep.txt
