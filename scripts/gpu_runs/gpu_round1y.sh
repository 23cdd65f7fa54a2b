#!/bin/bash
# ncu --set full of the product kernel on the config-4 and config-2 workloads (one launch each, layout pinned).
mkdir -p gpurun_out
export QBGPU_FORCE_FORMAT=sell
for w in tri31_k10 heis_chain32_k0; do
  timeout -k 5 150 ncu --set full --clock-control none --import-source on -k regex:spmv_sjds -s 4 -c 1 -o gpurun_out/prof_spmv_sjds_$w python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-lanczos > gpurun_out/ncu_$w.log 2>&1; echo "ncu $w rc=$?"
done
ls -la gpurun_out/*.ncu-rep
