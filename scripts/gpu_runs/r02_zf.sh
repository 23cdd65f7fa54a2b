#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_species.py tests/test_gpu_dynamics.py "tests/test_gpu_parity.py::test_hubbard4x3_against_the_references_own_run" "tests/test_gpu_parity.py::test_native_dist_drivers_one_rank" -q -p no:cacheprovider > gpurun_out/r02zf_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02zf_pytest.log
grep -E "^E  |^FAILED" gpurun_out/r02zf_pytest.log | head
for i in 1 2; do timeout -k 5 600 python bench.py --steps 20 --warmup 5 --no-species --no-cpu > gpurun_out/r02zf_bench_quick$i.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/r02zf_bench_quick$i.json').read().strip().splitlines()[-1]); print('ms', d['ms_per_step'], 'frac', d['roofline']['frac'], d['parity_sampled']['rel_l2_error'], 'lanczos', d.get('lanczos',{}).get('iters_per_s'), d.get('lanczos',{}).get('E0'))
PY
done
QBGPU_ORD_DESC=0 timeout -k 5 600 python bench.py --steps 20 --warmup 5 --no-species --no-cpu > gpurun_out/r02zf_bench_noorddesc.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/r02zf_bench_noorddesc.json').read().strip().splitlines()[-1]); print('no traversal descriptors: ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'lanczos', d.get('lanczos',{}).get('iters_per_s'))
PY
