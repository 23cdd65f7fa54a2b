#!/bin/bash
# Round 2 record, 2 GPUs: the 2-GPU pytest tests, the native checks with and without speculation, the bench line.
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "native_dist" > gpurun_out/r02z_pytest_gpu_2gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02z_pytest_gpu_2gpu.log
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_native_check.py 4 3 6 6 > gpurun_out/r02z_dist_check_n2.json 2> gpurun_out/r02z_dist_check_n2.err; echo "check n2 rc=$?"; tail -c 600 gpurun_out/r02z_dist_check_n2.json; tail -3 gpurun_out/r02z_dist_check_n2.err
timeout -k 5 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02z_bench_n2.json 2> gpurun_out/r02z_bench_n2.err; echo "bench n2 rc=$?"
QBGPU_DIST_SPECULATE=0 timeout -k 5 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02z_bench_n2_nospec.json 2> gpurun_out/r02z_bench_n2_nospec.err; echo "bench n2 nospec rc=$?"
python - <<PY
import json
for f in ('r02z_bench_n2','r02z_bench_n2_nospec'):
    d=json.loads([l for l in open('gpurun_out/'+f+'.json') if l.startswith('{')][-1])
    print(f, 'value', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'lanczos', d['lanczos']['iters_per_s'], d['lanczos']['steps'], d['lanczos']['E0'])
PY
tail -3 gpurun_out/r02z_bench_n2.err | grep -v OMP
