#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "hubbard_sector or hubbard4x2_sector or hubbard4x4_momentum" -x --durations=5 > gpurun_out/r02zc_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r02zc_pytest.log
