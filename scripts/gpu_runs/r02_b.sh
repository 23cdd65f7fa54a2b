#!/bin/bash
# Round 2, second GPU call: first hardware run of the bulk-streamed sliced-jagged kernel (cp.async.bulk rings) and of the
# block-local product with x staged in shared memory; the new species order (odd-site-major) and its permutation cost.
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_species.py -q -x -p no:cacheprovider -k "not arpack and not thick_restart" > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02b_pytest.log
grep -E "^E  " gpurun_out/r02b_pytest.log | head -20
QBGPU_SPECIES_TILE=128 timeout 600 python bench.py --species-probe --workload hubbard4x4 --steps 10 > gpurun_out/r02b_species_probe_W128.json 2> gpurun_out/r02b_species_probe_W128.err; tail -c 2500 gpurun_out/r02b_species_probe_W128.json; tail -3 gpurun_out/r02b_species_probe_W128.err
QBGPU_SJDS_BULK=0 QBGPU_BLOCK_SMEM=0 timeout 600 python bench.py --species-probe --workload hubbard4x4 --steps 10 > gpurun_out/r02b_species_probe_old_kernels.json 2> gpurun_out/r02b_species_probe_old_kernels.err; tail -c 2500 gpurun_out/r02b_species_probe_old_kernels.json
