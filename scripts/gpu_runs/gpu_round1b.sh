#!/bin/bash
# Second GPU pass (after the sub-warp shuffle-mask fix): every step under its own timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout -k 5 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout -k 5 200 compute-sanitizer --tool synccheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/smoke_synccheck.log
bash scripts/run_gpu_tests_isolated.sh 150 gpurun_out/pytest_gpu_isolated.log | tee gpurun_out/pytest_gpu_summary.txt
export QBGPU_VERBOSE=1
for w in heis_chain24 heis_chain28; do
  timeout -k 5 240 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"; tail -c 1800 gpurun_out/bench_$w.json; grep autotune gpurun_out/bench_$w.err
done
timeout -k 5 480 python bench.py --workload hubbard4x4 --steps 20 --warmup 5 > gpurun_out/bench_hubbard4x4.json 2> gpurun_out/bench_hubbard4x4.err; echo "bench hubbard4x4 rc=$?"; tail -c 2500 gpurun_out/bench_hubbard4x4.json; tail -8 gpurun_out/bench_hubbard4x4.err
timeout -k 5 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_heis28.csv python bench.py --workload heis_chain28 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:spmv_ -s 24 -c 2 -o gpurun_out/prof_spmv_heis28 python bench.py --workload heis_chain28 --steps 3 --warmup 3 --no-cpu --no-lanczos > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
