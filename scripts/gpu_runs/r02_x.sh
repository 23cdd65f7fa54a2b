#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -rs -x > gpurun_out/r02x_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02x_pytest_gpu.log
grep -E "^E  |^FAILED" gpurun_out/r02x_pytest_gpu.log | head -20
for i in 1 2; do timeout -k 5 600 python bench.py --steps 20 --warmup 5 --no-species --no-cpu > gpurun_out/r02x_bench_quick$i.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/r02x_bench_quick$i.json').read().strip().splitlines()[-1]); print('ms', d['ms_per_step'], 'frac', d['roofline']['frac'], d['clocks'], d['parity_sampled']['rel_l2_error'], 'lanczos', d.get('lanczos',{}).get('iters_per_s'), d.get('lanczos',{}).get('E0'))
PY
done
QBGPU_BLOCK_SMEM=1 timeout -k 5 600 python bench.py --steps 20 --warmup 5 --no-species --no-cpu > gpurun_out/r02x_bench_old_block.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/r02x_bench_old_block.json').read().strip().splitlines()[-1]); print('register-fed pass 1: ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'lanczos', d.get('lanczos',{}).get('iters_per_s'))
PY
