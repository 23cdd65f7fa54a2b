#!/usr/bin/env python
"""Pass 1 of the stored species handle (the block-local "local" part) alone, fp64 vectors: the register-fed block kernel
(variant 1) against the ring-fed one (sjds_block_bulk_kernel, variants 10..15) -- time per product and bit-equality of y, with
and without the z / gamma terms of the epilogue.  Usage: block_bulk_sweep.py <workload> [variants...]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import quantum_basis_b200 as qb
from quantum_basis_b200.csr import check as _check

workload = sys.argv[1]
variants = [int(v) for v in sys.argv[2:]] or [1, 10, 11, 12, 13, 14, 15]
L = qb.lib()
assert L.qbgpu_init(0) == 0
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
assert L.qbgpu_set_stream(C.c_void_p(stream.cuda_stream)) == 0
fam, p = bench.WORKLOADS[workload]
ns = p["Lx"] * p["Ly"]
M = qb.hubbard(ns, p["nup"], p["ndn"], bench.square_bonds(p["Lx"], p["Ly"]), p["t"], p["U"], flags=128)
R = M.real_view()
loc, cross = R.species_parts()
n = M.info.n
x = qb.vec_randomize(n, 1, dtype=np.float64, device=True)
z = qb.vec_randomize(n, 5, dtype=np.float64, device=True)
one, zero = (C.c_double * 2)(1.0, 0.0), (C.c_double * 2)(0.0, 0.0)
gam, bet = (C.c_double * 2)(0.37, 0.0), (C.c_double * 2)(-1.25, 0.0)


def timed(fn, steps=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(stream)
    for _ in range(steps):
        fn()
    b.record(stream); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


ref = {}
for v in variants:
    assert L.qbgpu_debug_set_variant(3000 + v) == 0
    y = qb.DeviceVector(n, np.float64)
    y2 = qb.DeviceVector(n, np.float64)
    ms = timed(lambda: _check(L.qbgpu_spmv_fused(loc.handle, C.c_void_p(x.ptr), None, C.c_void_p(y.ptr), one, zero, zero, None)))
    _check(L.qbgpu_spmv_fused(loc.handle, C.c_void_p(x.ptr), C.c_void_p(z.ptr), C.c_void_p(y2.ptr), one, gam, bet, None))
    torch.cuda.synchronize()
    a, b = y.to_numpy(), y2.to_numpy()
    if not ref:
        ref = {"a": a, "b": b}
        same = "reference"
    else:
        same = f"bit-identical {np.array_equal(a, ref['a'])} / with z,gamma {np.array_equal(b, ref['b'])} (max diff {np.abs(a - ref['a']).max():.3g}, {np.abs(b - ref['b']).max():.3g})"
    print(f"local part, fp64, block variant {v}: {ms:.3f} ms   {same}", flush=True)
    y.free(); y2.free()
# the whole reference-order product with the best variant is timed by bench.py (QBGPU_BLOCK_SMEM=<variant>)
