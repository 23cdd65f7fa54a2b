#!/bin/bash
mkdir -p gpurun_out
for SPEC in 1 0; do
QBGPU_DIST_SPECULATE=$SPEC timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$SPEC scripts/dist_native_check.py 4 3 6 6 > gpurun_out/r02zb_dist_check_n8_spec$SPEC.json 2> gpurun_out/r02zb_dist_check_n8_spec$SPEC.err; echo "check n8 spec=$SPEC rc=$?"; tail -c 100 gpurun_out/r02zb_dist_check_n8_spec$SPEC.json; grep "dist_native_check\]" gpurun_out/r02zb_dist_check_n8_spec$SPEC.err | head -8
done
