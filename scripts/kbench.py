#!/usr/bin/env python
"""Kernel-variant micro-benchmark (tuning aid, not the product bench): times every experimental instantiation of the
sliced-jagged H*v kernel (qbgpu_debug_set_variant) and the CSR-vector lane widths on one workload.
Usage: python scripts/kbench.py <workload> [--real] [--ids a,b,c] [--reps N]"""
import argparse
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
import quantum_basis_b200 as qb  # noqa: E402


EXTRA = {33: "minb=4 UL=0 U=4 S=L2ef X=adaptive", 34: "minb=4 UL=1 U=8 S=L2ef X=adaptive", 35: "minb=3 UL=0 U=4 S=L2ef X=L2el",
         36: "minb=5 UL=0 U=4 S=L2ef X=L2el", 37: "minb=6 UL=0 U=4 S=L2ef X=L2el", 38: "minb=6 UL=0 U=2 S=L2ef X=L2el",
         39: "minb=4 UL=0 U=6 S=L2ef X=L2el", 40: "minb=5 UL=0 U=4 S=L2ef X=adaptive", 41: "minb=3 UL=1 U=8 S=L2ef X=nc",
         42: "minb=5 UL=1 U=8 S=L2ef X=nc", 43: "minb=3 UL=1 U=12 S=L2ef X=nc", 44: "minb=4 UL=0 U=6 S=L2ef X=adaptive"}


def describe(v):
    if v == 0:
        return "production"
    if v in EXTRA:
        return EXTRA[v]
    w = v - 1
    return f"minb={2 if w // 16 == 0 else 4} UL={(w // 8) % 2} U={4 if (w // 4) % 2 == 0 else 8} S={'cs' if (w // 2) % 2 == 0 else 'L2ef'} X={'nc' if w % 2 == 0 else 'L2el'}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("--real", action="store_true", help="fp64 vectors (a `d` handle) instead of complex")
    ap.add_argument("--ids", default="")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--csr", action="store_true", help="also time the CSR-vector kernel at every lane width")
    ap.add_argument("--matfree", action="store_true", help="time the matrix-free product instead of a stored layout")
    ap.add_argument("--terms", action="store_true", help="with --matfree: the term-coded variant (QBGPU_MATFREE_TERMS)")
    ap.add_argument("--fused", action="store_true", help="time the fused-epilogue variants of the production kernel")
    ap.add_argument("--dict", action="store_true", help="1-byte value codes (QBGPU_VALUE_DICT)")
    ap.add_argument("--far", default="", help="comma list of log2(far_rows) to sweep for the adaptive-policy variants")
    a = ap.parse_args()
    L = qb.lib()
    assert L.qbgpu_init(0) == 0
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    L.qbgpu_set_stream(C.c_void_p(stream.cuda_stream))
    fam, p = bench.WORKLOADS[a.workload]
    t0 = time.time()
    h = C.c_void_p()
    cplx = 0 if a.real else 1
    flags = 8 | 2 | (16 if a.dict else 0) | (64 if a.terms else 0)          # sliced-jagged, no autotune
    if fam == "hubbard":
        bonds = np.array(bench.square_bonds(p["Lx"], p["Ly"]), dtype=np.int32).ravel()
        f = L.qbgpu_create_matfree_hubbard if a.matfree else L.qbgpu_build_hubbard
        rc = f(C.byref(h), p["Lx"] * p["Ly"], p["nup"], p["ndn"], len(bonds) // 2, bonds.ctypes.data, p["t"], p["U"], cplx, flags, 0, -1)
    else:
        n = p["L"]
        bonds = np.array([(x, (x + 1) % n) for x in range(n)], dtype=np.int32).ravel()
        f = L.qbgpu_create_matfree_heisenberg if a.matfree else L.qbgpu_build_heisenberg
        rc = f(C.byref(h), n, n // 2, n, bonds.ctypes.data, 1.0, cplx, flags, 0, -1)
    assert rc == 0, L.qbgpu_last_error()
    M = qb.csr_mat._adopt(h, bool(cplx))
    torch.cuda.synchronize()
    inf = M.info
    n, Z = inf.n, inf.nnz_stored
    if a.matfree:
        Z = 2 * bench.workload_upper_nnz(a.workload) - n        # entries regenerated per product
    s_vec = 16 if cplx else 8
    s_val = 1 if inf.value_dict else 8
    B = bench.algorithmic_bytes(Z, n, n, s_val, s_vec)
    if a.terms:
        print(f"# term-coded: {inf.nnz_input} bytes of codes (padding included) = {inf.nnz_input / Z:.3f} B per entry, handle {inf.device_bytes/1e9:.2f} GB", flush=True)
    print(f"# {a.workload}: n={n} Z={Z} B_spmv={B/1e9:.2f} GB (S_val={s_val},S_vec={s_vec}) dict={inf.value_dict} build {time.time()-t0:.1f}s", flush=True)
    dt = np.complex128 if cplx else np.float64
    x = qb.vec_randomize(n, 1, dtype=dt, device=True)
    y = qb.DeviceVector(n, dt)
    ids = [int(t) for t in a.ids.split(",")] if a.ids else list(range(0, 33))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    res = []
    fars = [int(t) for t in a.far.split(",")] if a.far else [21]
    runs = []
    for v in ids:
        if "adaptive" in describe(v):
            runs += [(v, f) for f in fars]
        else:
            runs.append((v, 21))
    for v, f in runs:
        L.qbgpu_debug_set_far_rows(1 << f)
        L.qbgpu_debug_set_variant(v)
        for _ in range(2):
            M.MultMv(x, y)
        e0.record(stream)
        for _ in range(a.reps):
            M.MultMv(x, y)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        res.append((ms, v))
        print(f"variant {v:2d} far=2^{f:<2d} [{describe(v):40s}] {ms:8.3f} ms  {B/ms/1e6:7.0f} GB/s  {B/ms/1e6/6451.8:5.3f} of measured peak", flush=True)
    L.qbgpu_debug_set_variant(0)
    if a.fused:
        dots = qb.DeviceVector(4, np.float64)
        one = (C.c_double * 2)(1.0, 0.0); zero = (C.c_double * 2)(0.0, 0.0); half = (C.c_double * 2)(-0.5, 0.0)
        cases = [("plain y=Hx", zero, zero, False, False), ("beta*z (z=y)", zero, half, True, False), ("dots", zero, zero, False, True),
                 ("beta*z + dots (Lanczos step a)", zero, half, True, True), ("gamma*x + dots (CG)", half, zero, False, True)]
        dot_cfgs = [0, 45, 46, 47, 48, 49, 50]
        names = {0: "production", 45: "U8 UL1 minb3", 46: "U8 UL1 minb2", 47: "U4 UL0 minb4", 48: "U4 UL0 minb3", 49: "U6 UL1 minb3", 50: "U4 UL1 minb4"}
        for label, gam, bet, usez, used, *rest in [(c + (0,)) for c in cases] + [(f"step a, dots cfg {names[v]}", zero, half, True, True, v) for v in dot_cfgs[1:]]:
            L.qbgpu_debug_set_variant(rest[0])
            def call():
                rc = L.qbgpu_spmv_fused(M.handle, C.c_void_p(x.ptr), C.c_void_p(y.ptr) if usez else None, C.c_void_p(y.ptr), one, gam, bet,
                                        C.c_void_p(dots.ptr) if used else None)
                assert rc == 0, L.qbgpu_last_error()
            for _ in range(2):
                call()
            e0.record(stream)
            for _ in range(a.reps):
                call()
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.reps
            print(f"fused [{label:32s}] {ms:8.3f} ms", flush=True)
        L.qbgpu_debug_set_variant(0)
    res.sort()
    print("# best:", ", ".join(f"{v}:{ms:.3f}" for ms, v in res[:5]))
    if a.csr:
        M.destroy()
        for lanes_flag in (4 | 2,):
            h2 = C.c_void_p()
            os.environ["QBGPU_VERBOSE"] = "1"
            if fam == "hubbard":
                rc = L.qbgpu_build_hubbard(C.byref(h2), p["Lx"] * p["Ly"], p["nup"], p["ndn"], len(bonds) // 2, bonds.ctypes.data, p["t"], p["U"], cplx, 4, 0, -1)
            else:
                rc = L.qbgpu_build_heisenberg(C.byref(h2), p["L"], p["L"] // 2, p["L"], bonds.ctypes.data, 1.0, cplx, 4, 0, -1)
            assert rc == 0, L.qbgpu_last_error()


if __name__ == "__main__":
    main()
