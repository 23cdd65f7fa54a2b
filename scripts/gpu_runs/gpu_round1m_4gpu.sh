#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 200 python -m pytest tests/test_gpu_parity.py::test_column_blocks_reproduce_the_fused_step -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/pt_blocks.log 2>&1; echo "rc=$? column blocks"; grep -E "^E  |passed|failed" gpurun_out/pt_blocks.log | head
timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/dist_check4.log 2>&1; echo "dist_check rc=$?"; grep -E "PASS|FAIL" gpurun_out/dist_check4.log | head -20
timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_4gpu_hubbard4x4.json 2> gpurun_out/bench_4gpu_hubbard4x4.err; echo "bench 4gpu rc=$?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_4gpu_hubbard4x4.json'))
print(d['value'], d['ms_per_step'], d['config']['exchange'], d['roofline']['local_product_ms'])
print({k: (v['iters_per_s'] if isinstance(v, dict) else v) for k, v in d['lanczos'].items()})
PY
