#!/bin/bash
# Peer-exchange variants (copy engines vs SM pulls, grouped column blocks) on the GPUs of this box.
mkdir -p gpurun_out
N=${1:-2}
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/peer_variants.py --steps 10 --warmup 3 $PV_ARGS > gpurun_out/peer_variants_n$N.jsonl 2> gpurun_out/peer_variants_n$N.err; echo "variants N=$N rc=$?"
cat gpurun_out/peer_variants_n$N.jsonl | cut -c1-420; tail -5 gpurun_out/peer_variants_n$N.err
