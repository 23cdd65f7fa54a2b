#!/bin/bash
# Record pass: all GPU tests (isolated), bench N=1 on BASELINE config 3 and config 1, ncu launch list + full capture
# of the tuned kernel with the layout pinned.
mkdir -p gpurun_out
timeout -k 5 100 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
bash scripts/run_gpu_tests_isolated.sh 200 gpurun_out/pytest_gpu_isolated.log | tee gpurun_out/pytest_gpu_summary.txt | grep -v "^PASS"
grep -E "^E  |FAILED" gpurun_out/pytest_gpu_isolated.log | head -30
timeout -k 5 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_hubbard4x4.json 2> gpurun_out/bench_hubbard4x4.err; echo "bench hubbard4x4 rc=$?"; tail -c 4500 gpurun_out/bench_hubbard4x4.json; tail -3 gpurun_out/bench_hubbard4x4.err
timeout -k 5 300 python bench.py --workload heis_chain20 --steps 20 --warmup 5 > gpurun_out/bench_heis20.json 2> gpurun_out/bench_heis20.err; echo "bench heis20 rc=$?"; tail -c 2500 gpurun_out/bench_heis20.json
timeout -k 5 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; echo "bench reference rc=$?"; tail -c 1500 gpurun_out/bench_reference_arm.json
timeout -k 5 300 python scripts/kbench.py hubbard4x4 --dict --ids 0 > gpurun_out/kbench_dict_hubbard.txt 2>&1; cat gpurun_out/kbench_dict_hubbard.txt
export QBGPU_FORCE_FORMAT=sell
timeout -k 5 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_hubbard4x4.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-lanczos > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout -k 5 500 ncu --set full --clock-control none --import-source on -k regex:spmv_sjds -s 4 -c 2 -o gpurun_out/prof_spmv_sjds_hubbard4x4 python bench.py --steps 3 --warmup 3 --no-cpu --no-lanczos > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | head -40
