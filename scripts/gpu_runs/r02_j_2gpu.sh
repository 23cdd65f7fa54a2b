#!/bin/bash
# Round 2 (2 GPUs): native multi-GPU drivers (csrc/dist.cu) -- correctness on 1 and 2 ranks, then the sharded bench.
mkdir -p gpurun_out
timeout -k 5 300 python scripts/dist_native_check.py > gpurun_out/r02j_dist_check_n1.json 2> gpurun_out/r02j_dist_check_n1.err; echo "check n1 rc=$?"; tail -c 1800 gpurun_out/r02j_dist_check_n1.json; tail -3 gpurun_out/r02j_dist_check_n1.err
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_native_check.py > gpurun_out/r02j_dist_check_n2.json 2> gpurun_out/r02j_dist_check_n2.err; echo "check n2 rc=$?"; tail -c 1800 gpurun_out/r02j_dist_check_n2.json; tail -5 gpurun_out/r02j_dist_check_n2.err
timeout -k 5 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02j_bench_n2.json 2> gpurun_out/r02j_bench_n2.err; echo "bench n2 rc=$?"; tail -c 3000 gpurun_out/r02j_bench_n2.json; tail -5 gpurun_out/r02j_bench_n2.err
