#!/usr/bin/env python
"""ncu target: a few internal-order products on the species-order handles of a Hubbard workload (stored two-pass and/or
matrix-free), nothing else.  Usage: species_ncu_target.py <workload> <stored|matrix_free|ordinary> [reps] [real]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import quantum_basis_b200 as qb

workload, kind = sys.argv[1], sys.argv[2]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
real = len(sys.argv) > 4 and sys.argv[4] == "real"
L = qb.lib()
assert L.qbgpu_init(0) == 0
fam, p = bench.WORKLOADS[workload]
ns = p["Lx"] * p["Ly"]
one, zero = (C.c_double * 2)(1.0, 0.0), (C.c_double * 2)(0.0, 0.0)
if kind == "ordinary":
    M = bench.build_matrix(qb, workload, flags=2 | 8)
else:
    M = qb.hubbard(ns, p["nup"], p["ndn"], bench.square_bonds(p["Lx"], p["Ly"]), p["t"], p["U"], flags=128,
                   matrix_free=(kind == "matrix_free"))
n = M.info.n
if real:
    H = M.real_view()
    x = qb.vec_randomize(n, 1, dtype=np.float64, device=True)
    y = qb.DeviceVector(n, np.float64)
else:
    H = M
    x = qb.vec_randomize(n, 1, device=True)
    y = qb.DeviceVector(n)
for _ in range(reps):
    assert L.qbgpu_spmv_fused(H.handle, C.c_void_p(x.ptr), None, C.c_void_p(y.ptr), one, zero, zero, None) == 0
torch.cuda.synchronize()
print("done", workload, kind, n)
