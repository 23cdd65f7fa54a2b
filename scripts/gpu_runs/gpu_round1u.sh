#!/bin/bash
# Refresh the multi-GPU bench line on N GPUs (driver-style launch).
mkdir -p gpurun_out
N=${1:-4}
timeout -k 5 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_hubbard4x4_n$N.json 2> gpurun_out/bench_hubbard4x4_n$N.err; echo "bench N=$N rc=$?"
tail -c 2500 gpurun_out/bench_hubbard4x4_n$N.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_hubbard4x4_n$N.err | tail -5
