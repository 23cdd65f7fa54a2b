#!/bin/bash
# First GPU pass: smoke, parity tests, first bench lines, ncu launch list + one full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
export QBGPU_VERBOSE=1
for w in heis_chain24 heis_chain28; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"; tail -c 1500 gpurun_out/bench_$w.json; grep autotune gpurun_out/bench_$w.err
done
timeout 900 python bench.py --workload hubbard4x4 --steps 20 --warmup 5 > gpurun_out/bench_hubbard4x4.json 2> gpurun_out/bench_hubbard4x4.err; echo "bench hubbard4x4 rc=$?"; tail -c 2500 gpurun_out/bench_hubbard4x4.json; tail -8 gpurun_out/bench_hubbard4x4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_heis28.csv python bench.py --workload heis_chain28 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_csr_vector -s 22 -c 2 -o gpurun_out/prof_spmv_heis28 python bench.py --workload heis_chain28 --steps 3 --warmup 3 --no-cpu --no-lanczos > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
