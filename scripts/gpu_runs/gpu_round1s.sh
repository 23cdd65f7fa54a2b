#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider -k "orbit or config4" > gpurun_out/pytest_orbit.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_orbit.log
