#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python scripts/dist_native_check.py 4 3 6 6 > gpurun_out/r02m_dist_check_n1.json 2> gpurun_out/r02m_dist_check_n1.err; echo "check n1 rc=$?"; tail -c 700 gpurun_out/r02m_dist_check_n1.json; tail -3 gpurun_out/r02m_dist_check_n1.err
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_native_check.py 4 3 6 6 > gpurun_out/r02m_dist_check_n2.json 2> gpurun_out/r02m_dist_check_n2.err; echo "check n2 rc=$?"; tail -c 900 gpurun_out/r02m_dist_check_n2.json; tail -5 gpurun_out/r02m_dist_check_n2.err
timeout -k 5 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02m_bench_n2.json 2> gpurun_out/r02m_bench_n2.err; echo "bench n2 rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02m_bench_n2.json') if l.startswith('{')][-1])
print('N=2 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
print(json.dumps(d['products'])); print(json.dumps(d.get('lanczos')))
PY
tail -5 gpurun_out/r02m_bench_n2.err
