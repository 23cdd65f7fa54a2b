#!/bin/bash
mkdir -p gpurun_out
bash scripts/run_gpu_tests_isolated.sh 240 gpurun_out/pytest_gpu_isolated.log | tee gpurun_out/pytest_gpu_summary.txt | grep -v "^PASS"
grep -E "^E  |FAILED" gpurun_out/pytest_gpu_isolated.log | head -30
export QBGPU_VERBOSE=1
timeout -k 5 300 python scripts/kbench.py hubbard4x4 --matfree --ids 0 --fused > gpurun_out/kbench6_matfree_hubbard.txt 2>&1; grep -E "^#|^variant|^fused" gpurun_out/kbench6_matfree_hubbard.txt
timeout -k 5 300 python scripts/kbench.py hubbard4x4 --matfree --real --ids 0 > gpurun_out/kbench6_matfree_hubbard_real.txt 2>&1; grep -E "^#|^variant" gpurun_out/kbench6_matfree_hubbard_real.txt
timeout -k 5 300 python scripts/kbench.py heis_chain28 --matfree --ids 0 > gpurun_out/kbench6_matfree_heis28.txt 2>&1; grep -E "^#|^variant" gpurun_out/kbench6_matfree_heis28.txt
timeout -k 5 400 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_hubbard4x4_b.json 2> gpurun_out/bench_hubbard4x4_b.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_hubbard4x4_b.json')); print({k:d[k] for k in ('value','ms_per_step','lanczos','lanczos_value_dict') if k in d})"; grep "qbgpu lanczos" gpurun_out/bench_hubbard4x4_b.err
