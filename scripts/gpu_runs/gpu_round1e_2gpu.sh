#!/bin/bash
# 2-GPU pass: sharded correctness check, then the sharded bench on BASELINE config 3.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus2.txt 2>&1
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/dist_check.log 2>&1; echo "dist_check rc=$?"; grep -E "PASS|FAIL|Error|error" gpurun_out/dist_check.log | head -20
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_hubbard4x4.json 2> gpurun_out/bench_2gpu_hubbard4x4.err; echo "bench 2gpu rc=$?"; tail -c 3000 gpurun_out/bench_2gpu_hubbard4x4.json; tail -5 gpurun_out/bench_2gpu_hubbard4x4.err
