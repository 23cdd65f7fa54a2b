#!/bin/bash
mkdir -p gpurun_out
for t in tests/test_gpu_parity.py::test_column_blocks_reproduce_the_fused_step tests/test_gpu_parity.py::test_device_resident_thick_restart_lanczos; do
  QBGPU_VERBOSE=1 timeout -k 5 200 python -m pytest $t -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/pt_$(basename $t | tr ':' '_').log 2>&1; echo "rc=$? $t"; grep -E "^E  |passed|failed|trlan\]" gpurun_out/pt_$(basename $t | tr ':' '_').log | head -30
done
