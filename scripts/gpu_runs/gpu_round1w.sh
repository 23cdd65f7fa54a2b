#!/bin/bash
# Final regression of round 1 on one GPU: smoke, every GPU test, the default bench line.
mkdir -p gpurun_out
timeout -k 5 100 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout -k 5 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 gpurun_out/pytest_gpu_full.log
timeout -k 5 600 python bench.py > gpurun_out/bench_hubbard4x4.json 2> gpurun_out/bench_hubbard4x4.err; echo "bench rc=$?"; head -c 900 gpurun_out/bench_hubbard4x4.json; echo; tail -2 gpurun_out/bench_hubbard4x4.err
