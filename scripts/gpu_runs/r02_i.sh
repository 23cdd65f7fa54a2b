#!/bin/bash
# Round 2: full GPU suite on the new defaults; bench main line (species handle); reference arm on config 3 itself.
mkdir -p gpurun_out
nproc > gpurun_out/r02i_host.txt; free -g >> gpurun_out/r02i_host.txt
timeout -k 5 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02i_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02i_pytest_gpu.log
grep -E "^E  |^FAILED" gpurun_out/r02i_pytest_gpu.log | head -20
timeout -k 5 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02i_bench_n1.json 2> gpurun_out/r02i_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02i_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02i_bench_n1.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','roofline','e2e','parity_sampled','lanczos','real_vectors','cpu_baseline','gpu_launches','clocks','config'):
    print(k, json.dumps(d.get(k))[:700])
PY
( time timeout -k 5 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02i_bench_reference.json 2> gpurun_out/r02i_bench_reference.err ) 2>&1 | grep real; tail -c 1500 gpurun_out/r02i_bench_reference.json; tail -3 gpurun_out/r02i_bench_reference.err
