#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider -k "orbit_assembler_matches" > gpurun_out/pytest_orbit2.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_orbit2.log
