#!/bin/bash
# Round 2 record, 8 GPUs: the native multi-GPU checks on 8 ranks, the bench line the driver will produce, the C++ example.
mkdir -p gpurun_out
N=8
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_native_check.py 4 3 6 6 > gpurun_out/r02y_dist_check_n8.json 2> gpurun_out/r02y_dist_check_n8.err; echo "check n8 rc=$?"; tail -c 200 gpurun_out/r02y_dist_check_n8.json
date +%s > gpurun_out/r02y_t0
QBGPU_VERBOSE=1 timeout -k 5 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02y_bench_n${N}.json 2> gpurun_out/r02y_bench_n${N}.err; echo "bench n$N rc=$? in $(( $(date +%s) - $(cat gpurun_out/r02y_t0) )) s"
grep "dist_lanczos" gpurun_out/r02y_bench_n8.err | tail -2
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02y_bench_n8.json') if l.startswith('{')][-1])
print('N=8 value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))
for k,v in d['products'].items(): print(' ', k, {kk:(round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk not in ('parts',)})
print(' lanczos', json.dumps(d.get('lanczos')))
PY
if [ -f examples/dist_lanczos ]; then timeout -k 5 120 ./examples/dist_lanczos 8 4 3 6 6 > gpurun_out/r02y_example_dist_lanczos_n8.txt 2>&1; echo "example rc=$?"; tail -2 gpurun_out/r02y_example_dist_lanczos_n8.txt; fi
