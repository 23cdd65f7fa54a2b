#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python bench.py --steps 20 --warmup 5 --no-species --no-cpu --no-lanczos > gpurun_out/r02zg_bench_quick.json 2>gpurun_out/r02zg_bench_quick.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02zg_bench_quick.json').read().strip().splitlines()[-1]); print('ms', d['ms_per_step'], 'frac', d['roofline']['frac'], d['clocks'])
PY
tail -3 gpurun_out/r02zg_bench_quick.err
