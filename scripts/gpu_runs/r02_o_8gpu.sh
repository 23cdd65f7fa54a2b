#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_native_check.py 4 3 6 6 > gpurun_out/r02o_dist_check_n8.json 2> gpurun_out/r02o_dist_check_n8.err; echo "check n8 rc=$?"; tail -c 300 gpurun_out/r02o_dist_check_n8.json
N=8
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02o_bench_n${N}.json 2> gpurun_out/r02o_bench_n${N}.err; echo "bench n$N rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02o_bench_n8.json') if l.startswith('{')][-1])
print('N=8 value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))
for k,v in d['products'].items(): print(' ', k, {kk:(round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk not in ('parts',)})
print(' lanczos', json.dumps(d.get('lanczos')))
PY
tail -2 gpurun_out/r02o_bench_n8.err | grep -v OMP_NUM | grep -v '\*\*\*'
