#!/bin/bash
# 8 GPUs: the bench line with the speculative early parts of the Lanczos loop (default) and without (QBGPU_DIST_SPECULATE=0)
mkdir -p gpurun_out
N=8
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02za_bench_n8.json 2> gpurun_out/r02za_bench_n8.err; echo "bench n8 rc=$?"
QBGPU_DIST_SPECULATE=0 timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29529 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02za_bench_n8_nospec.json 2> gpurun_out/r02za_bench_n8_nospec.err; echo "bench n8 nospec rc=$?"
python - <<PY
import json
for f in ('r02za_bench_n8','r02za_bench_n8_nospec'):
    d=json.loads([l for l in open('gpurun_out/'+f+'.json') if l.startswith('{')][-1])
    print(f, 'value', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'lanczos', round(d['lanczos']['iters_per_s'],1), d['lanczos']['steps'], d['lanczos']['E0'], d['products']['fp64']['decomposition'])
PY
timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_native_check.py 4 3 6 6 > gpurun_out/r02za_dist_check_n8.json 2> gpurun_out/r02za_dist_check_n8.err; echo "check n8 rc=$?"; tail -c 150 gpurun_out/r02za_dist_check_n8.json
