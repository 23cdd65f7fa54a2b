#!/bin/bash
# Round 2 (8 GPUs): refined shards -- correctness on 8 ranks, the C++ example, then the sharded bench at N = 8, 4, 2.
mkdir -p gpurun_out
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_native_check.py 4 3 6 6 > gpurun_out/r02n_dist_check_n8.json 2> gpurun_out/r02n_dist_check_n8.err; echo "check n8 rc=$?"; tail -c 400 gpurun_out/r02n_dist_check_n8.json
g++ -std=c++17 -O2 -I include examples/dist_lanczos.cc -L quantum_basis_b200 -lqbgpu -Wl,-rpath,$PWD/quantum_basis_b200 -o /tmp/dist_lanczos && for N in 1 2 8; do timeout -k 5 120 /tmp/dist_lanczos $N; echo "example N=$N rc=$?"; done 2>&1 | tee gpurun_out/r02n_example_dist_lanczos.txt
for N in 8 4 2; do
for LANES in 1 4; do
QB_DIST_LANES=$LANES timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02n_bench_n${N}_lanes$LANES.json 2> gpurun_out/r02n_bench_n${N}_lanes$LANES.err; echo "bench n$N lanes $LANES rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02n_bench_n${N}_lanes$LANES.json') if l.startswith('{')][-1])
print('N=$N lanes=$LANES value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))
for k,v in d['products'].items(): print(' ', k, {kk:(round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk not in ('parts',)})
print(' lanczos', json.dumps(d.get('lanczos')))
PY
tail -2 gpurun_out/r02n_bench_n${N}_lanes$LANES.err | grep -v OMP_NUM | grep -v '\*\*\*'
[ $N != 8 ] && break
done
done
