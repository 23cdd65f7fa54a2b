#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus8.txt 2>&1
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/dist_check8.log 2>&1; echo "dist_check rc=$?"; grep -E "PASS|FAIL|Error|error" gpurun_out/dist_check8.log | head -20
timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu_hubbard4x4.json 2> gpurun_out/bench_8gpu_hubbard4x4.err; echo "bench 8gpu rc=$?"; tail -c 3500 gpurun_out/bench_8gpu_hubbard4x4.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_8gpu_hubbard4x4.err | tail -8
