#!/bin/bash
# Round 2 (8 GPUs): native drivers on 8 ranks (correctness), then the sharded bench at N = 8, 4.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02l_topo.txt 2>&1
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_native_check.py 4 3 6 6 > gpurun_out/r02l_dist_check_n8.json 2> gpurun_out/r02l_dist_check_n8.err; echo "check n8 rc=$?"; tail -c 600 gpurun_out/r02l_dist_check_n8.json
for N in 8 4; do
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02l_bench_n$N.json 2> gpurun_out/r02l_bench_n$N.err; echo "bench n$N rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02l_bench_n$N.json') if l.startswith('{')][-1])
print('N=$N value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
print(json.dumps(d['products'])); print(json.dumps(d.get('lanczos'))); print(d['config']['rows_per_rank'])
PY
tail -3 gpurun_out/r02l_bench_n$N.err
done
