#!/bin/bash
mkdir -p gpurun_out
N=8
for MODES in rows3_needed_rows rows4_needed_rows rows6_needed_rows; do
QBGPU_VERBOSE=1 QB_DIST_MODES=$MODES timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02p_bench_n8_$MODES.json 2> gpurun_out/r02p_bench_n8_$MODES.err; echo "bench $MODES rc=$?"
grep "dist_lanczos" gpurun_out/r02p_bench_n8_$MODES.err | tail -2
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02p_bench_n8_$MODES.json') if l.startswith('{')][-1])
print('$MODES value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'lanczos it/s', round(d['lanczos']['iters_per_s'],1), 'cplx ms', round(d['products']['complex128']['ms_per_product'],3), 'segs', d['products']['fp64']['pull_segments_rank0'])
PY
done
