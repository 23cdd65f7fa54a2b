#!/bin/bash
# Final record pass of round 1 (one GPU): every GPU test, the bench lines, ncu evidence with the layout pinned.
mkdir -p gpurun_out
timeout -k 5 100 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout -k 5 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 gpurun_out/pytest_gpu_full.log
export QBGPU_VERBOSE=1
timeout -k 5 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_hubbard4x4.json 2> gpurun_out/bench_hubbard4x4.err; echo "bench hubbard4x4 rc=$?"; tail -c 3800 gpurun_out/bench_hubbard4x4.json; grep "qbgpu lanczos" gpurun_out/bench_hubbard4x4.err
timeout -k 5 300 python bench.py --workload heis_chain20 --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_heis20.json 2> gpurun_out/bench_heis20.err; echo "bench heis20 rc=$?"
unset QBGPU_VERBOSE
export QBGPU_FORCE_FORMAT=sell
timeout -k 5 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_hubbard4x4.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-lanczos > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout -k 5 500 ncu --set full --clock-control none --import-source on -k regex:spmv_sjds -s 4 -c 1 -o gpurun_out/prof_spmv_sjds_hubbard4x4 python bench.py --steps 3 --warmup 3 --no-cpu --no-lanczos > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | head -30
