#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_species.py tests/test_gpu_dynamics.py -q -p no:cacheprovider > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02h_pytest.log
grep -E "^E  |^FAILED" gpurun_out/r02h_pytest.log | head -20
bash scripts/gpu_runs/r02_g.sh
QBGPU_SPECIES_TILE=64 timeout 600 python bench.py --species-probe --workload hubbard4x4 --steps 10 > gpurun_out/r02h_species_probe.json 2> gpurun_out/r02h_species_probe.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02h_species_probe.json').read().strip().splitlines()[-1])
for k in ('stored','matrix_free'):
    print(k, {kk:(round(v['ms_per_product'],3) if isinstance(v,dict) and 'ms_per_product' in v else v) for kk,v in d[k].items() if kk!='pass1_variants_complex_ms'})
PY
