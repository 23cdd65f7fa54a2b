#!/usr/bin/env python
"""Time the plain product (internal-order vectors, fused entry point) of the ordinary and the stored species handle of a
Hubbard workload across the configurations of the bulk-streamed kernel (qbgpu_debug_set_variant(2000 + mode)) and of the
block-local kernel (3000 + variant).  Usage: bulk_sweep.py <workload> [real]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import quantum_basis_b200 as qb

workload = sys.argv[1]
L = qb.lib()
assert L.qbgpu_init(0) == 0
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
assert L.qbgpu_set_stream(C.c_void_p(stream.cuda_stream)) == 0
fam, p = bench.WORKLOADS[workload]
one, zero = (C.c_double * 2)(1.0, 0.0), (C.c_double * 2)(0.0, 0.0)


def timed(fn, steps=8):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(stream)
    for _ in range(steps):
        fn()
    b.record(stream); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def sweep(tag, M, modes, setter_base):
    n = M.info.n
    for real in (False, True):
        H = M.real_view() if real else M
        x = qb.vec_randomize(n, 1, dtype=np.float64 if real else np.complex128, device=True)
        y = qb.DeviceVector(n, np.float64 if real else np.complex128)
        for m in modes:
            assert L.qbgpu_debug_set_variant(setter_base + m) == 0
            ms = timed(lambda: L.qbgpu_spmv_fused(H.handle, C.c_void_p(x.ptr), None, C.c_void_p(y.ptr), one, zero, zero, None))
            print(f"{tag} {'fp64' if real else 'cplx'} variant {setter_base + m}: {ms:.3f} ms", flush=True)
        x.free(); y.free()


if "--species-only" not in sys.argv:
    M = bench.build_matrix(qb, workload, flags=2 | 8)
    sweep("ordinary", M, [0, 1, 2, 3, 4, 6, 9], 2000)
    M.destroy()
if fam == "hubbard":
    ns = p["Lx"] * p["Ly"]
    M = qb.hubbard(ns, p["nup"], p["ndn"], bench.square_bonds(p["Lx"], p["Ly"]), p["t"], p["U"], flags=128)
    if "--quick" in sys.argv:
        L.qbgpu_debug_set_variant(3001)
        sweep("species-stored(production)", M, [10], 2000)
        sys.exit(0)
    L.qbgpu_debug_set_variant(2001)
    sweep("species-stored(block variants; cross=bulk mode 1)", M, [1, 3, 4, 6, 7], 3000)
    L.qbgpu_debug_set_variant(3001)
    sweep("species-stored(block variant 1; cross bulk modes)", M, [0, 1, 2, 3, 4, 5, 6, 7, 8, 9], 2000)
