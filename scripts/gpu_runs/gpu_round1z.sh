#!/bin/bash
# L2 fetch-granularity hint: production kernel on config 3 with the default, 32-, 64- and 128-byte settings.
mkdir -p gpurun_out
: > gpurun_out/kbench_l2fetch.txt
for g in default 32 128; do
  if [ "$g" = default ]; then unset QBGPU_L2_FETCH_BYTES; else export QBGPU_L2_FETCH_BYTES=$g; fi
  for mode in "" "--real"; do
    echo "## L2 fetch granularity = $g $mode" >> gpurun_out/kbench_l2fetch.txt
    timeout -k 5 100 python scripts/kbench.py hubbard4x4 --ids 0 $mode 2>&1 | grep "variant" >> gpurun_out/kbench_l2fetch.txt
  done
done
cat gpurun_out/kbench_l2fetch.txt
