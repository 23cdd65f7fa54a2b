#!/bin/bash
# the last GPU call of round 2 (about 100 s of budget left): the step-level entry points first, then the whole GPU suite
# (the eigenvec_CG loop body was moved into shared pieces: every CG test must still pass)
mkdir -p gpurun_out
timeout -k 3 85 python -m pytest tests/test_gpu_zz_step_entries.py tests/test_gpu_ab.py tests/test_gpu_resume.py tests/test_gpu_parity.py tests/test_gpu_dynamics.py tests/test_gpu_emax.py tests/test_gpu_species.py \
    -m gpu -q -p no:cacheprovider -rf > gpurun_out/r02zk_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02zk_pytest.log
