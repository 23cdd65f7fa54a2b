#!/bin/bash
# L2 persisting set-aside experiment + Lanczos phase breakdown.
mkdir -p gpurun_out
export QBGPU_VERBOSE=1
IDS=0,19,20,33,36,40
for mb in 0 32 64 999; do
  if [ $mb = 0 ]; then export QBGPU_NO_L2_PERSIST=1; else unset QBGPU_NO_L2_PERSIST; export QBGPU_L2_PERSIST_MB=$mb; fi
  timeout -k 5 300 python scripts/kbench.py hubbard4x4 --ids $IDS --far 18,20,22 > gpurun_out/kbench3_hubbard_persist$mb.txt 2>&1; echo "== persist $mb MB rc=$?"; grep -E "qbgpu\]|^variant" gpurun_out/kbench3_hubbard_persist$mb.txt
done
unset QBGPU_L2_PERSIST_MB
timeout -k 5 300 python scripts/kbench.py hubbard4x4 --real --ids 0,20,31,33 --far 19,21 > gpurun_out/kbench3_hubbard_real_persist.txt 2>&1; grep -E "^variant" gpurun_out/kbench3_hubbard_real_persist.txt
timeout -k 5 300 python - > gpurun_out/lanczos_phases.txt 2>&1 <<'PY'
import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, '.')
import bench, quantum_basis_b200 as qb
L = qb.lib(); L.qbgpu_init(0)
M = bench.build_matrix(qb, "hubbard4x4")
n = M.dim
v = qb.DeviceVector(2 * n); L.qbgpu_vec_randomize_z(n, C.c_void_p(v.ptr), 1)
hess = np.zeros(200)
m = qb.lanczos(0, 40, 100, n, M, v, hess, "dnmcs")
print("steps", m)
PY
cat gpurun_out/lanczos_phases.txt | tail -5
