#!/usr/bin/env python
"""Isolate a mismatch of the complex two-pass species product: internal-order fused product of the stored species handle
(Hubbard 4x3) against the ordinary handle, for every combination of pass-1 / pass-2 kernel."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import quantum_basis_b200 as qb
from quantum_basis_b200.bench_support import square_bonds
L = qb.lib(); assert L.qbgpu_init(0) == 0
Lx, Ly, nu, nd = (int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (4, 3, 6, 6)
ns, bonds = Lx * Ly, square_bonds(Lx, Ly)
P = qb.hubbard(ns, nu, nd, bonds, 1.0, 1.1)
M = qb.hubbard(ns, nu, nd, bonds, 1.0, 1.1, flags=128)
n = P.dim
rng = np.random.default_rng(3)
x = rng.normal(size=n) + 1j * rng.normal(size=n)
yref = np.zeros(n, dtype=np.complex128); P.MultMv(x, yref)
perm = M.native_perm()
xi = np.empty_like(x); xi[perm] = x
want = np.empty_like(x); want[perm] = yref
xd, yd = qb.DeviceVector.from_numpy(xi), qb.DeviceVector(n)
one, zero = (C.c_double * 2)(1.0, 0.0), (C.c_double * 2)(0.0, 0.0)
loc, cross = M.species_parts()
for bulk in (2000, 2001, 2007, 2009, 2010):
    for blk in (3000, 3001, 3007):
        L.qbgpu_debug_set_variant(bulk); L.qbgpu_debug_set_variant(blk)
        yd.zero()
        assert L.qbgpu_spmv_fused(M.handle, C.c_void_p(xd.ptr), None, C.c_void_p(yd.ptr), one, zero, zero, None) == 0
        e = np.linalg.norm(yd.to_numpy() - want) / np.linalg.norm(want)
        # the parts one by one
        yd.zero()
        assert L.qbgpu_zmv(loc.handle, one, C.c_void_p(xd.ptr), zero, C.c_void_p(yd.ptr), 1) == 0
        y1 = yd.to_numpy()
        assert L.qbgpu_zmv(cross.handle, one, C.c_void_p(xd.ptr), one, C.c_void_p(yd.ptr), 1) == 0
        e2 = np.linalg.norm(yd.to_numpy() - want) / np.linalg.norm(want)
        print(f"bulk {bulk} block {blk}: fused {e:.2e}   parts {e2:.2e}   |local part| {np.linalg.norm(y1):.6f}", flush=True)
