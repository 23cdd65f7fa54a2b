#!/bin/bash
mkdir -p gpurun_out
export QBGPU_VERBOSE=1
timeout -k 5 300 python scripts/kbench.py hubbard4x4 --real --ids 0 --fused > gpurun_out/kbench4_fused_real.txt 2>&1; grep -E "^variant|^fused" gpurun_out/kbench4_fused_real.txt
timeout -k 5 300 python scripts/kbench.py hubbard4x4 --ids 0 --fused > gpurun_out/kbench4_fused_complex.txt 2>&1; grep -E "^variant|^fused" gpurun_out/kbench4_fused_complex.txt
timeout -k 5 300 python - > gpurun_out/lanczos_phases2.txt 2>&1 <<'PY'
import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, '.')
import bench, quantum_basis_b200 as qb
L = qb.lib(); L.qbgpu_init(0)
M = bench.build_matrix(qb, "hubbard4x4")
n = M.dim
for rep in range(2):
    v = qb.DeviceVector(2 * n); L.qbgpu_vec_randomize_z(n, C.c_void_p(v.ptr), 1)
    hess = np.zeros(200)
    m = qb.lanczos(0, 40, 100, n, M, v, hess, "dnmcs")
    v.free()
PY
grep "qbgpu lanczos" gpurun_out/lanczos_phases2.txt
