#!/usr/bin/env python
"""ncu target: reference-order MultMv (device vectors, real content) on the stored species handle of a Hubbard workload."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import quantum_basis_b200 as qb
workload = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
L = qb.lib()
assert L.qbgpu_init(0) == 0
fam, p = bench.WORKLOADS[workload]
ns = p["Lx"] * p["Ly"]
M = qb.hubbard(ns, p["nup"], p["ndn"], bench.square_bonds(p["Lx"], p["Ly"]), p["t"], p["U"], flags=128)
n = M.info.n
x = qb.vec_randomize(n, 1, device=True)
y = qb.DeviceVector(n)
for _ in range(reps):
    M.MultMv(x, y)
torch.cuda.synchronize()
print("done")
