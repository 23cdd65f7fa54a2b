#!/bin/bash
# S^z_q between sectors + dynamic Lanczos parity, then the config-5 flow at L=28.
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider -k "sector or config5" > gpurun_out/pytest_sectors2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_sectors2.log
timeout -k 5 300 python scripts/config5_flow.py --L 20 --nmom 256 --maxit 100 > gpurun_out/config5_L20.json 2> gpurun_out/config5_L20.err; echo "config5 L20 rc=$?"; tail -c 1500 gpurun_out/config5_L20.json; tail -3 gpurun_out/config5_L20.err
timeout -k 5 900 python scripts/config5_flow.py --L 28 --nmom 1024 --maxit 200 > gpurun_out/config5_L28.json 2> gpurun_out/config5_L28.err; echo "config5 L28 rc=$?"; tail -c 6000 gpurun_out/config5_L28.json; tail -3 gpurun_out/config5_L28.err
