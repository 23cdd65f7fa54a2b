#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python scripts/kbench.py hubbard4x4 --real --ids 0 --fused > gpurun_out/kbench5_fused_real.txt 2>&1; grep -E "^variant|^fused" gpurun_out/kbench5_fused_real.txt
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/dist_check.log 2>&1; echo "dist_check rc=$?"; grep -E "PASS|FAIL|Error|error" gpurun_out/dist_check.log | head -20
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_hubbard4x4.json 2> gpurun_out/bench_2gpu_hubbard4x4.err; echo "bench 2gpu rc=$?"; tail -c 3500 gpurun_out/bench_2gpu_hubbard4x4.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_2gpu_hubbard4x4.err | tail -8
