#!/usr/bin/env python
"""Correctness of the native multi-GPU drivers (csrc/dist.cu, qbgpu_dist_*) against the single-GPU entry points, on N ranks:

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_native_check.py [Lx Ly NUP NDN]

also runs with N = 1 (plain python: the same loops without peers).  Every rank builds its species-order shard of the Hubbard
model AND the full single-GPU handle (small sizes), and compares: the sharded product, Lanczos (steps, E0, coefficients),
eigenvec_CG (residual of the eigenvector), energy_scale, KPM moments; then the same for an ordinary row shard (Heisenberg chain).
Prints one JSON line per rank-0 check and exits non-zero on a mismatch."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import quantum_basis_b200 as qb
from quantum_basis_b200 import dist as qd
from quantum_basis_b200.bench_support import square_bonds


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    L = qb.lib()
    assert L.qbgpu_init(local_rank) == 0
    # one stream for torch and the library (like bench.py): torch allocations, fills and copies are then ordered with the kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert L.qbgpu_set_stream(C.c_void_p(stream.cuda_stream)) == 0
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

        def exchange(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
    else:
        exchange = lambda b: [b]          # noqa: E731
    a = sys.argv[1:]
    Lx, Ly, nup, ndn = (int(a[0]), int(a[1]), int(a[2]), int(a[3])) if len(a) >= 4 else (4, 3, 6, 6)
    ns, bonds = Lx * Ly, square_bonds(Lx, Ly)
    ok = True
    out = {"world": world, "model": f"hubbard {Lx}x{Ly} {nup},{ndn}"}

    def check(name, cond, detail):
        nonlocal ok
        ok = ok and bool(cond)
        out[name] = detail if cond else {"FAILED": detail}
        if not cond:
            print(f"[dist_native_check] rank {rank}: {name} FAILED: {detail}", file=sys.stderr, flush=True)

    # ------------------------------------------------------------------ species-order shards, fp64 and complex vectors
    n = L.qbgpu_dim_hubbard(ns, nup, ndn)
    bounds, d_dn = qd.species_nnz_balanced_bounds(ns, nup, ndn, bonds, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    nloc = hi - lo
    full = qb.hubbard(ns, nup, ndn, bonds, 1.0, 1.1, flags=128)             # single-GPU handle (reference-order boundary)
    perm = full.native_perm()
    x_ref = qb.vec_randomize(n, 1)                                           # host, reference order
    y_ref = np.zeros(n, dtype=np.complex128)
    full.MultMv(x_ref, y_ref)
    x_int = np.empty_like(x_ref); x_int[perm] = x_ref
    y_int = np.empty_like(y_ref); y_int[perm] = y_ref
    v = np.zeros(2 * n, dtype=np.complex128); v[:n] = x_ref
    hess1 = np.zeros(2000)
    m1 = qb.lanczos(0, 999, 1000, n, full, v, hess1, "sr_val0")
    e1 = qb.hess_eigen(hess1, 1000, m1)[0][0]
    lo1, hi1 = qb.energy_scale(n, full, np.zeros(2 * n, dtype=np.complex128), 0.1, 40)
    mu1 = qb.kpm_moments(full, x_ref, lo1, hi1, 64)
    ref_rows = torch.empty(max(nloc, 1), dtype=torch.int32, device="cuda")
    assert L.qbgpu_species_ref_rows(ns, nup, ndn, lo, hi, C.c_void_p(ref_rows.data_ptr())) == 0, L.qbgpu_last_error()
    M = qb.hubbard(ns, nup, ndn, bonds, 1.0, 1.1, flags=128, rows=(lo, hi))
    for tag, cplx in (("fp64", False), ("complex", True)):
        # refined shard (early / late parts, pull plan) unless QB_DIST_PLAIN=1: then the plain (local, cross) decomposition
        shard = qd.SpeciesShard(qb, M, ns, nup, ndn, bonds, bounds, rank, world, real=not cplx, gap=2)
        loc, cross = shard.local, shard.cross
        D = qd.NativeDist(qb, n, bounds, rank, world, cplx, exchange)
        mode = os.environ.get("QB_DIST_MODE", "rows2_needed_rows" if cplx else "refined_needed_rows")
        shard.install(D, mode, lanes=int(os.environ.get("QB_DIST_LANES", "1")))
        out[f"{tag}_shard"] = {"mode": mode, "early": len(shard.installed[0]), "late": len(shard.installed[1]), "segments": len(shard.plan)}
        D.randomize(0, 1, ref_rows.data_ptr())
        got = D.download_own(0)
        check(f"{tag}_start_vector", np.abs(got - (x_int[lo:hi] if cplx else x_int[lo:hi].real)).max() <= 1e-15, "rows of vec_randomize(1) in the internal order")
        y = torch.zeros((2 if cplx else 1) * max(nloc, 1), dtype=torch.float64, device="cuda")
        D.mv(loc, cross, 0, y.data_ptr(), barrier=True)
        yh = y.cpu().numpy()
        yh = yh.view(np.complex128)[:nloc] if cplx else yh[:nloc]
        want = y_int[lo:hi] if cplx else y_int[lo:hi].real
        err = np.linalg.norm(yh - want) / max(np.linalg.norm(want), 1e-300)
        check(f"{tag}_product", err <= 1e-13, {"rel_l2_vs_single_gpu": float(err)})
        D.randomize(0, 1, ref_rows.data_ptr())
        m, hess = D.lanczos(loc, cross, 999, 1000, "sr_val0")
        e = qb.hess_eigen(hess, 1000, m)[0][0]
        check(f"{tag}_lanczos", abs(m - m1) <= 1 and abs(e - e1) <= 1e-10 * abs(e1) and np.abs(hess[1000:1020] - hess1[1000:1020]).max() < 1e-10,
              {"steps": m, "steps_single": int(m1), "E0": float(e), "E0_single": float(e1), "max_da_first20": float(np.abs(hess[1000:1020] - hess1[1000:1020]).max())})
        lo2, hi2 = D.energy_scale(loc, cross, 0.1, 40, ref_rows.data_ptr())
        check(f"{tag}_energy_scale", abs(lo2 - lo1) < 1e-9 and abs(hi2 - hi1) < 1e-9, {"lo": lo2, "hi": hi2, "lo_single": lo1, "hi_single": hi1})
        D.randomize(0, 1, ref_rows.data_ptr())
        mu = D.kpm_moments(loc, cross, lo1, hi1, 64)
        check(f"{tag}_kpm", np.abs(mu - mu1).max() < 1e-9, {"max_abs_diff_vs_single": float(np.abs(mu - mu1).max())})
        # eigenvec_CG: the ground-state vector of E0; checked by its residual |H v - E0 v| and its norm
        es = 8 * (2 if cplx else 1)
        bufs = [torch.zeros((2 if cplx else 1) * max(nloc, 1), dtype=torch.float64, device="cuda") for _ in range(4)]
        D.randomize(0, 1, ref_rows.data_ptr())
        assert L.qbgpu_memcpy_d2d(C.c_void_p(bufs[0].data_ptr()), C.c_void_p(D.own(0)), nloc * es) == 0
        mc, accu = D.eigenvec_cg(loc, cross, e1, *[b.data_ptr() for b in bufs])
        assert L.qbgpu_memcpy_d2d(C.c_void_p(D.own(0)), C.c_void_p(bufs[0].data_ptr()), nloc * es) == 0
        D.mv(loc, cross, 0, y.data_ptr(), barrier=True)
        hv = y.cpu().numpy(); vv = bufs[0].cpu().numpy()
        if cplx:
            hv, vv = hv.view(np.complex128), vv.view(np.complex128)
        part = torch.tensor([np.linalg.norm(hv[:nloc] - e1 * vv[:nloc]) ** 2, np.linalg.norm(vv[:nloc]) ** 2], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(part)
        res, nrm = float(part[0].sqrt().item()), float(part[1].sqrt().item())
        check(f"{tag}_eigenvec_cg", accu < 2e-12 and res < 1e-9 and abs(nrm - 1.0) < 1e-10, {"steps": mc, "accuracy": accu, "residual": res, "norm": nrm})
        D.destroy(); shard.destroy()
    M.destroy(); full.destroy()

    # ------------------------------------------------------------------ an ordinary row shard (reference order, no local part)
    Lc = 16
    cb = [(q, (q + 1) % Lc) for q in range(Lc)]
    nh = L.qbgpu_dim_heisenberg(Lc, Lc // 2)
    bh, _ = qd.equal_row_bounds(nh, world)
    fullh = qb.heisenberg(Lc, Lc // 2, cb, 1.0)
    xh = qb.vec_randomize(nh, 1)
    yh_ref = np.zeros(nh, dtype=np.complex128)
    fullh.MultMv(xh, yh_ref)
    vh = np.zeros(2 * nh, dtype=np.complex128); vh[:nh] = xh
    hh = np.zeros(2000)
    mh = qb.lanczos(0, 999, 1000, nh, fullh, vh, hh, "sr_val0")
    eh = qb.hess_eigen(hh, 1000, mh)[0][0]
    S = qb.heisenberg(Lc, Lc // 2, cb, 1.0, rows=(bh[rank], bh[rank + 1]))
    D = qd.NativeDist(qb, nh, bh, rank, world, True, exchange)
    D.randomize(0, 1, None)
    nl = bh[rank + 1] - bh[rank]
    y = torch.zeros(2 * max(nl, 1), dtype=torch.float64, device="cuda")
    D.mv(None, S, 0, y.data_ptr(), barrier=True)
    got = y.cpu().numpy().view(np.complex128)[:nl]
    err = np.linalg.norm(got - yh_ref[bh[rank]:bh[rank + 1]]) / np.linalg.norm(yh_ref)
    check("ordinary_shard_product", err <= 1e-13, {"rel_l2_vs_single_gpu": float(err)})
    D.randomize(0, 1, None)
    m, hess = D.lanczos(None, S, 999, 1000, "sr_val0")
    e = qb.hess_eigen(hess, 1000, m)[0][0]
    check("ordinary_shard_lanczos", abs(m - mh) <= 1 and abs(e - eh) <= 1e-10 * abs(eh), {"steps": m, "steps_single": int(mh), "E0": float(e), "E0_single": float(eh)})
    D.destroy()

    flag = torch.tensor([0 if ok else 1], dtype=torch.int32, device="cuda")
    if world > 1:
        dist.all_reduce(flag)
    out["all_ranks_ok"] = int(flag.item()) == 0
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 0 else 1)


if __name__ == "__main__":
    main()
