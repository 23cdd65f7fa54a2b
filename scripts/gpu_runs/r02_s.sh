#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_species.py tests/test_gpu_dynamics.py "tests/test_gpu_parity.py::test_hubbard4x3_against_the_references_own_run" "tests/test_gpu_parity.py::test_kpm_moments_pinned_to_the_references_product" -q -p no:cacheprovider > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02s_pytest.log
grep -E "^E  |^FAILED" gpurun_out/r02s_pytest.log | head
bash scripts/gpu_runs/r02_g.sh
for i in 1 2; do timeout -k 5 600 python bench.py --steps 20 --warmup 5 --no-species --no-cpu --no-lanczos > gpurun_out/r02s_bench_quick$i.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/r02s_bench_quick$i.json').read().strip().splitlines()[-1]); print('ms', d['ms_per_step'], 'frac', d['roofline']['frac'], d['clocks'], d['parity_sampled']['rel_l2_error'])
PY
done
QBGPU_FUSE_WAY_OUT=0 timeout -k 5 600 python bench.py --steps 20 --warmup 5 --no-species --no-cpu --no-lanczos > gpurun_out/r02s_bench_nofuse.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/r02s_bench_nofuse.json').read().strip().splitlines()[-1]); print('nofuse ms', d['ms_per_step'], 'frac', d['roofline']['frac'])
PY
