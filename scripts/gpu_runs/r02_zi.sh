#!/bin/bash
mkdir -p gpurun_out
QBGPU_VERBOSE=1 timeout -k 5 200 python __graft_entry__.py smoke > gpurun_out/r02zi_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02zi_smoke.log
timeout -k 5 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "not hubbard4x4_momentum" > gpurun_out/r02zi_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02zi_pytest.log
grep -E "^E  |^FAILED" gpurun_out/r02zi_pytest.log | head
