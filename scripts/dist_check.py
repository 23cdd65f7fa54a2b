#!/usr/bin/env python
"""Multi-GPU correctness check (run under torchrun, one rank per GPU): the row-sharded product and the sharded Lanczos of
quantum_basis_b200/dist.py -- both exchange schemes, complex and real-mode vectors -- against the single-GPU library on
the same matrix (Heisenberg chain L=24, Sz=0: dim 2,704,156).  Prints PASS/FAIL lines; exit code 1 on any failure."""
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


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    L = qb.lib()
    assert L.qbgpu_init(lr) == 0
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    L.qbgpu_set_stream(C.c_void_p(stream.cuda_stream))
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    wl = "heis_chain24"
    n = L.qbgpu_dim_heisenberg(24, 12)
    bounds, chunk = qd.equal_row_bounds(n, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    M = bench.build_matrix(qb, wl, row_range=(lo, hi))
    Mfull = bench.build_matrix(qb, wl)                       # every rank keeps the whole matrix as the checker
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
    xfull = rng.normal(size=n) + 1j * rng.normal(size=n)
    xfull /= np.linalg.norm(xfull)
    yref = np.zeros(n, dtype=np.complex128)
    Mfull.MultMv(qb.DeviceVector.from_numpy(xfull), (ydev := qb.DeviceVector(n)))
    yref = ydev.to_numpy()
    x0 = qb.vec_randomize(n, 1)
    v = np.zeros(2 * n, dtype=np.complex128); v[:n] = x0
    hess = np.zeros(200)
    m = qb.lanczos(0, 60, 100, n, Mfull, v, hess, "dnmcs")
    kern = qd.DeviceKernels(qb, M, real=False, parts=parts)
    for name, Op in (("allgather", qd.ShardedOperator), ("pipelined", qd.PipelinedOperator)):
        op = Op(kern, n, rank, world, comm)
        xl = kern.alloc(chunk); yl = kern.alloc(chunk)
        xl[:2 * (hi - lo)] = torch.from_numpy(np.ascontiguousarray(xfull[lo:hi]).view(np.float64)).cuda()
        op.matvec(xl, yl)
        torch.cuda.synchronize()
        y = yl.cpu().numpy().view(np.complex128)[: hi - lo]
        err = np.linalg.norm(y - yref[lo:hi]) / np.linalg.norm(yref[lo:hi])
        e = torch.tensor([err], device="cuda"); dist.all_reduce(e, op=dist.ReduceOp.MAX)
        report(f"sharded product ({name}, complex x)", e.item() < 1e-13, f"rel_l2={e.item():.2e}")

    # peer-memory pull exchange (CUDA IPC + copy engines)
    try:
        opC = qd.PeerExchangeOperator(qb, kern, n, rank, world, comm, torch)
        opC.own(0).upload(np.ascontiguousarray(xfull[lo:hi]))
        comm.all_reduce(opC.token)
        yl = kern.alloc(chunk)
        for _ in range(3):
            opC.matvec(0, yl)
        torch.cuda.synchronize()
        y = yl.cpu().numpy().view(np.complex128)[: hi - lo]
        err = np.linalg.norm(y - yref[lo:hi]) / np.linalg.norm(yref[lo:hi])
        e = torch.tensor([err], device="cuda"); dist.all_reduce(e, op=dist.ReduceOp.MAX)
        report("sharded product (peer pull, complex x)", e.item() < 1e-13, f"rel_l2={e.item():.2e}")
        for real in (True, False):
            kr = qd.DeviceKernels(qb, M, real=real, parts=parts)
            oC = qd.PeerExchangeOperator(qb, kr, n, rank, world, comm, torch)
            oC.own(0).upload(np.ascontiguousarray(x0[lo:hi].real if real else x0[lo:hi]))
            comm.all_reduce(oC.token)
            state = torch.zeros(8, dtype=torch.float64, device="cuda"); state[0] = 1.0
            a_dev = torch.zeros(100, dtype=torch.float64, device="cuda"); b_dev = torch.zeros(100, dtype=torch.float64, device="cuda")
            qd.peer_lanczos(oC, 30, state, a_dev, b_dev)
            torch.cuda.synchronize()
            a = a_dev.cpu().numpy(); b = b_dev.cpu().numpy()
            da = np.abs(a[:25] - hess[100:125]).max(); db = np.abs(b[1:25] - hess[1:25]).max()
            report(f"sharded Lanczos (peer pull, {'fp64' if real else 'complex'} vectors) vs single GPU", da < 1e-11 and db < 1e-11, f"max|da|={da:.1e} max|db|={db:.1e}")
    except AssertionError as ex:
        report("peer pull exchange", False, f"unavailable or failed: {ex}")

    # Lanczos: single-GPU fused loop (real mode) vs sharded loops, real and complex vectors
    for real in (True, False):
        kr = qd.DeviceKernels(qb, M, real=real, parts=parts)
        for name, Op, fn in (("allgather", qd.ShardedOperator, qd.sharded_lanczos), ("pipelined", qd.PipelinedOperator, qd.pipelined_lanczos)):
            op = Op(kr, n, rank, world, comm)
            u0 = kr.alloc(chunk); u1 = kr.alloc(chunk)
            if real:
                u0[: hi - lo] = torch.from_numpy(np.ascontiguousarray(x0[lo:hi].real)).cuda()
            else:
                u0[: 2 * (hi - lo)] = torch.from_numpy(np.ascontiguousarray(x0[lo:hi]).view(np.float64)).cuda()
            state = torch.zeros(8, dtype=torch.float64, device="cuda"); state[0] = 1.0
            a_dev = torch.zeros(100, dtype=torch.float64, device="cuda"); b_dev = torch.zeros(100, dtype=torch.float64, device="cuda")
            fn(op, u0, u1, 100, 30, state, a_dev, b_dev)
            torch.cuda.synchronize()
            a = a_dev.cpu().numpy(); b = b_dev.cpu().numpy()
            da = np.abs(a[:25] - hess[100:125]).max(); db = np.abs(b[1:25] - hess[1:25]).max()
            report(f"sharded Lanczos ({name}, {'fp64' if real else 'complex'} vectors) vs single GPU", da < 1e-11 and db < 1e-11, f"max|da|={da:.1e} max|db|={db:.1e}")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
