#!/bin/bash
# Round 2, sixth GPU call: production kernel selection (block-local + bulk-streamed for fp64 vectors, real-content route of
# the reference-order product) -- species tests and the probe.
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_species.py tests/test_gpu_dynamics.py tests/test_gpu_resume.py tests/test_gpu_emax.py -q -p no:cacheprovider > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02f_pytest.log
grep -E "^E  |^FAILED" gpurun_out/r02f_pytest.log | head -20
timeout 600 python bench.py --species-probe --workload hubbard4x4 --steps 10 > gpurun_out/r02f_species_probe.json 2> gpurun_out/r02f_species_probe.err; tail -c 3500 gpurun_out/r02f_species_probe.json; tail -3 gpurun_out/r02f_species_probe.err
for W in 64 256; do QBGPU_SPECIES_TILE=$W timeout 600 python scripts/bulk_sweep.py hubbard4x4 --species-only --quick > gpurun_out/r02f_tile_$W.txt 2>&1; echo "tile $W"; grep -v "bulk kernel" gpurun_out/r02f_tile_$W.txt | head -8; done
