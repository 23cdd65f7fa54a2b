#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 120 python scripts/block_bulk_sweep.py hubbard4x3 1 10 12 > gpurun_out/r02u_sweep_4x3.txt 2>&1; echo "4x3 rc=$?"; tail -4 gpurun_out/r02u_sweep_4x3.txt
timeout -k 5 600 python scripts/block_bulk_sweep.py hubbard4x4 > gpurun_out/r02u_sweep_4x4.txt 2>&1; echo "4x4 rc=$?"; tail -8 gpurun_out/r02u_sweep_4x4.txt
timeout -k 5 600 python -m pytest tests/test_gpu_species.py tests/test_gpu_dynamics.py "tests/test_gpu_parity.py::test_matrix_free_sector_product_equals_the_stored_sector_matrix" "tests/test_gpu_parity.py::test_matrix_free_sector_reaches_the_published_E0" -q -p no:cacheprovider > gpurun_out/r02u_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02u_pytest.log
grep -E "^E  |^FAILED" gpurun_out/r02u_pytest.log | head
