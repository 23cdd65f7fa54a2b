"""World-size-2 (and 3) CPU tests of the multi-GPU orchestration in quantum_basis_b200/dist.py over gloo.

The device kernels cannot run here, so the `kernels` provider is an oracle-backed numpy restatement of the three fused
Lanczos passes with exactly the semantics of include/qbgpu.h (qbgpu_lanczos_step_a/b/c, qbgpu_zmv on a row shard).
What is under test is the host logic that is shared with the GPU run: the row partition, the padded all-gather layout,
where the two scalar all-reduces sit, and the scale rotation -- the sharded recurrence must reproduce the reference's
Lanczos coefficients on the whole matrix.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


class OracleKernels:
    def __init__(self, F_rows, lo, hi, n):
        self.F, self.lo, self.hi, self.n = F_rows, lo, hi, n

    def alloc(self, nentries):
        return torch.zeros(2 * nentries, dtype=torch.float64)

    def slot(self, state, i):
        return state[i:i + 1]

    @staticmethod
    def _c(t):
        return t.numpy().view(np.complex128)

    def multmv(self, x_full, y_local):
        self._c(y_local)[: self.hi - self.lo] = self.F @ self._c(x_full)[: self.n]

    def lanczos_step_a(self, x_full, uz, state):
        sx, sz, bprev = state[0].item(), state[1].item(), state[2].item()
        nl = self.hi - self.lo
        x = self._c(x_full)
        z = self._c(uz)
        w = sx * (self.F @ x[: self.n])
        if bprev * sz != 0.0:
            w = w - bprev * sz * z[:nl]
        z[:nl] = w
        chunk = x_full.numel() // 2 // dist.get_world_size()
        xl = x[dist.get_rank() * chunk: dist.get_rank() * chunk + nl]
        state[3] = float(np.real(np.vdot(sx * xl, w)))

    def segment(self, t, first_entry, nentries):
        return t[2 * first_entry: 2 * (first_entry + nentries)]

    def set_col_bounds(self, col_bounds):
        self.cb = col_bounds
        Fc = self.F.tocsc()
        self.Fp = [Fc[:, col_bounds[p]:col_bounds[p + 1]].tocsr() for p in range(len(col_bounds) - 1)]

    def multmv_part(self, p, x_full, y_local, accumulate):
        nl = self.hi - self.lo
        part = self.Fp[p] @ self._c(x_full)[self.cb[p]:self.cb[p + 1]]
        y = self._c(y_local)
        y[:nl] = y[:nl] + part if accumulate else part

    def lanczos_step_a_part(self, p, x_full, uz, state, first, last):
        sx, sz, bprev = state[0].item(), state[1].item(), state[2].item()
        nl = self.hi - self.lo
        x = self._c(x_full)
        z = self._c(uz)
        w = sx * (self.Fp[p] @ x[self.cb[p]:self.cb[p + 1]])
        if first:
            if bprev * sz != 0.0:
                w = w - bprev * sz * z[:nl]
        else:
            w = w + z[:nl]
        z[:nl] = w
        if last:
            chunk = x_full.numel() // 2 // dist.get_world_size()
            xl = x[dist.get_rank() * chunk: dist.get_rank() * chunk + nl]
            state[3] = float(np.real(np.vdot(sx * xl, w)))

    def lanczos_step_b(self, ux, uz, state):
        nl = self.hi - self.lo
        z = self._c(uz)
        z[:nl] = z[:nl] - state[3].item() * state[0].item() * self._c(ux)[:nl]
        state[6] = float(np.vdot(z[:nl], z[:nl]).real)

    def lanczos_step_c(self, state, a_dev, b_dev, m):
        b = float(np.sqrt(state[6].item()))
        a_dev[m - 1] = state[3].item()
        b_dev[m] = b
        sx_old = state[0].item()
        state[0] = 1.0 / b
        state[1] = sx_old
        state[2] = b


def _worker(rank, world, port, name, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as O
    from quantum_basis_b200 import dist as qdist
    A, meta, ex = O.load_golden(name)
    n = A.dim
    F = A.to_scipy_full()
    bounds, chunk = qdist.equal_row_bounds(n, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    kern = OracleKernels(F[lo:hi], lo, hi, n)
    op = qdist.ShardedOperator(kern, n, rank, world, qdist.TorchComm())
    assert (op.lo, op.hi, op.chunk) == (lo, hi, chunk)
    x = O.vec_randomize(n, 1)
    x_loc = kern.alloc(chunk); y_loc = kern.alloc(chunk)
    OracleKernels._c(x_loc)[: hi - lo] = x[lo:hi]
    op.matvec(x_loc, y_loc)
    y = OracleKernels._c(y_loc)[: hi - lo].copy()
    err_mv = np.linalg.norm(y - ex["y1"][lo:hi]) / np.linalg.norm(ex["y1"][lo:hi])
    steps = 25
    state = torch.zeros(8, dtype=torch.float64); state[0] = 1.0
    a_dev = torch.zeros(64, dtype=torch.float64); b_dev = torch.zeros(64, dtype=torch.float64)
    qdist.sharded_lanczos(op, x_loc.clone(), kern.alloc(chunk), 64, steps, state, a_dev, b_dev)
    # the pipelined variant (one broadcast per owner, one column block per owner) must give the same numbers
    pop = qdist.PipelinedOperator(kern, n, rank, world, qdist.TorchComm())
    kern.set_col_bounds(pop.col_bounds)
    y2 = kern.alloc(chunk)
    pop.matvec(x_loc, y2)
    err_pipe = np.linalg.norm(OracleKernels._c(y2)[: hi - lo] - ex["y1"][lo:hi]) / np.linalg.norm(ex["y1"][lo:hi])
    state2 = torch.zeros(8, dtype=torch.float64); state2[0] = 1.0
    a2 = torch.zeros(64, dtype=torch.float64); b2 = torch.zeros(64, dtype=torch.float64)
    qdist.pipelined_lanczos(pop, x_loc.clone(), kern.alloc(chunk), 64, steps, state2, a2, b2)
    err_mv = max(err_mv, err_pipe, float((a2 - a_dev).abs().max()) * 1e-2, float((b2 - b_dev).abs().max()) * 1e-2)
    q.put((rank, err_mv, a_dev.numpy().copy(), b_dev.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,name", [(2, "tri4x4_k01"), (3, "hubbard4x2"), (2, "honeycomb3x2_general")])
def test_sharded_product_and_lanczos_over_gloo(world, name, oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    A, meta, ex = oracle.load_golden(name)
    m, a, b, _ = oracle.lanczos(A, oracle.vec_randomize(A.dim, 1), 64, "dnmcs")
    for rank, err_mv, a_s, b_s in res:
        assert err_mv < 1e-13
        assert np.abs(a_s[:15] - a[:15]).max() < 1e-11          # every rank holds the same all-reduced coefficients
        assert np.abs(b_s[:15] - b[:15]).max() < 1e-11
        assert np.abs(a_s[:20] - ex["dn_a"][:20]).max() < 1e-10  # and they are the compiled reference's


def test_equal_row_bounds():
    from quantum_basis_b200.dist import equal_row_bounds
    for n, parts in ((10, 3), (165636900, 8), (7, 8), (16, 4)):
        b, chunk = equal_row_bounds(n, parts)
        assert b[0] == 0 and b[-1] == n and len(b) == parts + 1
        assert all(0 <= b[i + 1] - b[i] <= chunk for i in range(parts))
        assert chunk * parts >= n


def test_matching_rounds_are_conflict_free_and_cover_every_needed_transfer():
    """The pull schedule of PeerExchangeOperator(schedule="matching"): in every round a source serves at most one reader
    and a reader pulls from at most one source; every needed (reader, source) pair appears exactly once; a d-regular
    need graph takes exactly d rounds (Hubbard 4x4 on 8 ranks: 5 instead of the ring's 7)."""
    from quantum_basis_b200.dist import matching_rounds
    allowed = {0: {0, 1, 2}, 1: {0, 1, 3}, 2: {0, 2, 3}, 3: {1, 2, 3}}
    cases = {
        "hubbard8": [[(p // 2 in allowed[r // 2]) and p != r for p in range(8)] for r in range(8)],
        "full4": [[p != r for p in range(4)] for r in range(4)],
        "irregular": [[False, True, True, False, True], [True, False, False, False, False], [False, True, False, True, False],
                      [False, False, False, False, False], [True, True, True, True, False]],
    }
    for name, needs in cases.items():
        rounds = matching_rounds(needs)
        seen = set()
        for rd in rounds:
            srcs = [p for p in rd if p is not None]
            assert len(srcs) == len(set(srcs)), name
            for r, p in enumerate(rd):
                if p is not None:
                    assert needs[r][p] and (r, p) not in seen, name
                    seen.add((r, p))
        assert seen == {(r, p) for r in range(len(needs)) for p in range(len(needs)) if needs[r][p]}, name
    assert len(matching_rounds(cases["hubbard8"])) == 5 and len(matching_rounds(cases["full4"])) == 3
