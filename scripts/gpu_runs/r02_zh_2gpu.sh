#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "native_dist" > gpurun_out/r02zh_pytest_gpu_2gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02zh_pytest_gpu_2gpu.log
timeout -k 5 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02zh_bench_n2.json 2> gpurun_out/r02zh_bench_n2.err; echo "bench n2 rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02zh_bench_n2.json') if l.startswith('{')][-1])
print('N=2 value', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'lanczos', d['lanczos']['iters_per_s'], d['lanczos']['steps'], d['lanczos']['E0'], d['clocks'])
PY
