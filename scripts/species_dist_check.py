#!/usr/bin/env python
"""Multi-GPU check and first timing of the matrix-free species-order product (run under torchrun, one rank per GPU).
NOT YET RUN ON HARDWARE (written after round 1's GPU budget was spent); the single-GPU kernels must pass
tests/test_zz_gpu_species.py first.

Every rank holds a shard of whole up configurations (dist.ROW_ALIGN = species_row_align(D_dn)) cut into one column part per
owner; the three exchanges of dist.py (all-gather, pipelined broadcasts, peer pulls) multiply them; vectors are in the
internal order.  Checked against the unsharded handle on the same device; then timed.

  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/species_dist_check.py [Lx Ly nup ndn]
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import bench  # noqa: E402
import quantum_basis_b200 as qb  # noqa: E402
from quantum_basis_b200 import dist as qd  # noqa: E402

SPECIES = 128


def main():
    args = [int(a) for a in sys.argv[1:5]] or [4, 3, 6, 6]
    Lx, Ly, nup, ndn = args
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    L = qb.lib()
    assert L.qbgpu_init(lr) == 0
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    L.qbgpu_set_stream(C.c_void_p(stream.cuda_stream))
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    from math import comb
    ns, bonds = Lx * Ly, bench.square_bonds(Lx, Ly)
    d_dn = comb(ns, ndn)
    n = comb(ns, nup) * d_dn
    qd.ROW_ALIGN = qd.species_row_align(d_dn)
    bounds, chunk = qd.equal_row_bounds(n, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    mk = lambda rows=None: qb.hubbard(ns, nup, ndn, bonds, 1.0, 1.1, matrix_free=True, flags=SPECIES, rows=rows)   # noqa: E731
    M, Mfull = mk((lo, hi)), mk()
    comm = qd.TorchComm()
    col_bounds = [min(n, q * chunk) for q in range(world)] + [n]
    parts = qd.DeviceKernels.split(qb, M, col_bounds)
    ok = True

    def report(name, good, detail=""):
        nonlocal ok
        ok = ok and good
        if rank == 0:
            print(("PASS " if good else "FAIL ") + name + " " + detail, flush=True)

    rng = np.random.default_rng(5)
    xfull = rng.normal(size=n) + 1j * rng.normal(size=n)        # internal order throughout
    xfull /= np.linalg.norm(xfull)
    one, zero = (C.c_double * 2)(1.0, 0.0), (C.c_double * 2)(0.0, 0.0)
    xd, yd = qb.DeviceVector.from_numpy(xfull), qb.DeviceVector(n)
    assert L.qbgpu_spmv_fused(Mfull.handle, C.c_void_p(xd.ptr), None, C.c_void_p(yd.ptr), one, zero, zero, None) == 0
    yref = yd.to_numpy()
    kern = qd.DeviceKernels(qb, M, real=False, parts=parts)
    times = {}
    for name, Op in (("allgather", qd.ShardedOperator), ("pipelined", qd.PipelinedOperator)):
        op = Op(kern, n, rank, world, comm)
        xl = kern.alloc(chunk); yl = kern.alloc(chunk)
        xl[:2 * (hi - lo)] = torch.from_numpy(np.ascontiguousarray(xfull[lo:hi]).view(np.float64)).cuda()
        op.matvec(xl, yl)
        torch.cuda.synchronize()
        y = yl.cpu().numpy().view(np.complex128)[: hi - lo]
        err = np.linalg.norm(y - yref[lo:hi]) / max(1e-300, np.linalg.norm(yref[lo:hi])) if hi > lo else 0.0
        e = torch.tensor([err], device="cuda"); dist.all_reduce(e, op=dist.ReduceOp.MAX)
        report(f"species shards ({name}, complex x)", e.item() < 1e-13, f"rel_l2={e.item():.2e}")
        times[name] = qd._timed(torch, dist, stream, lambda: op.matvec(xl, yl), 10, 3)
    try:
        opC = qd.PeerExchangeOperator(qb, kern, n, rank, world, comm, torch)
        opC.own(0).upload(np.ascontiguousarray(xfull[lo:hi]))
        comm.all_reduce(opC.token)
        yl = kern.alloc(chunk)
        opC.matvec(0, yl)
        torch.cuda.synchronize()
        y = yl.cpu().numpy().view(np.complex128)[: hi - lo]
        err = np.linalg.norm(y - yref[lo:hi]) / max(1e-300, np.linalg.norm(yref[lo:hi])) if hi > lo else 0.0
        e = torch.tensor([err], device="cuda"); dist.all_reduce(e, op=dist.ReduceOp.MAX)
        report("species shards (peer pull, complex x)", e.item() < 1e-13, f"rel_l2={e.item():.2e}")
        times["peer_pull"] = qd._timed(torch, dist, stream, lambda: opC.matvec(0, yl), 10, 3)
    except AssertionError as ex:
        report("peer pull exchange", False, f"unavailable or failed: {ex}")
    if rank == 0:
        print("ms per product:", {k: round(v, 3) for k, v in times.items()}, f"(n = {n}, {world} ranks)", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
