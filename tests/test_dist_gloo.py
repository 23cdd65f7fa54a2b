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


# ------------------------------------------------------------------ PeerExchangeOperator, simulated in one process
class _FakePeerLib:
    """The handful of C-ABI entry points PeerExchangeOperator uses, on host memory: every simulated rank is a thread of
    this process, so a "peer mapping" is just the owner's address and a pull is a memmove."""
    import ctypes as _C

    def __init__(self):
        self.keep = []
        self.pulled = {}                                           # thread rank -> list of (slot) pulled in the last product

    def qbgpu_last_error(self):
        return b"fake"

    def qbgpu_malloc(self, pref, nbytes):
        import ctypes as C
        buf = (C.c_char * max(1, int(nbytes)))()
        self.keep.append(buf)
        pref._obj.value = C.addressof(buf)
        return 0

    def qbgpu_memset0(self, p, nbytes):
        import ctypes as C
        C.memset(p.value, 0, int(nbytes))
        return 0

    def qbgpu_memcpy_h2d(self, dst, src, nbytes):
        import ctypes as C
        C.memmove(dst.value, src.value, int(nbytes))
        return 0

    qbgpu_memcpy_d2h = qbgpu_memcpy_h2d

    def qbgpu_ipc_export(self, p, handle):
        import ctypes as C
        import struct
        C.memmove(handle, struct.pack("<Q", p.value) + bytes(56), 64)
        return 0

    def qbgpu_ipc_open(self, handle, out):
        import struct
        out._obj.value = struct.unpack("<Q", bytes(handle)[:8])[0]
        return 0

    def qbgpu_peer_pull_async(self, lane, slot, dst, src, nbytes):
        import ctypes as C
        import threading
        C.memmove(dst.value, src.value, int(nbytes))
        self.pulled.setdefault(threading.current_thread().name, []).append(slot)
        return 0

    def qbgpu_peer_wait(self, slot):
        return 0


def _simulate_peer_exchange(world, n, F, schedule):
    """Run PeerExchangeOperator.matvec on `world` threads; returns (y per rank, pulled owners per rank)."""
    import ctypes as C
    import threading
    import types
    import scipy.sparse as sp
    from quantum_basis_b200 import dist as qd

    lib = _FakePeerLib()
    fake_qb = types.SimpleNamespace(lib=lambda: lib)
    barrier = threading.Barrier(world)
    slots = {}
    lock = threading.Lock()
    tl = threading.local()

    def all_gather_object(out, obj):
        with lock:
            tl.calls = getattr(tl, "calls", 0) + 1
            slots.setdefault(tl.calls, {})[tl.rank] = obj
        barrier.wait()
        for r in range(world):
            out[r] = slots[tl.calls][r]
        barrier.wait()

    class Comm:
        def all_reduce(self, t):
            barrier.wait()

    class Kern:
        ncomp = 2

        def __init__(self, rows, lo, hi, col_bounds):
            Fc = rows.tocsc()
            self.lo, self.hi, self.cb = lo, hi, col_bounds
            self.blocks = [Fc[:, col_bounds[p]:col_bounds[p + 1]].tocsr() for p in range(len(col_bounds) - 1)]
            self.parts = [types.SimpleNamespace(info=types.SimpleNamespace(nnz_stored=b.nnz)) for b in self.blocks]

        def multmv_part(self, p, x_full, y_local, accumulate):
            nb = (self.cb[p + 1] - self.cb[p]) * 16
            raw = (C.c_char * nb).from_address(x_full.ptr + self.cb[p] * 16)
            x = np.frombuffer(raw, dtype=np.complex128)
            part = self.blocks[p] @ x
            y_local[:] = y_local + part if accumulate else part

    x = (np.arange(n) + 1.0) * np.exp(0.3j * np.arange(n))
    bounds, chunk = qd.equal_row_bounds(n, world)
    ys, errs = [None] * world, []
    saved = qd.os.environ.get("QB_PEER_SCHEDULE")
    qd.os.environ["QB_PEER_SCHEDULE"] = schedule
    import torch.distributed as tdist
    orig = tdist.all_gather_object
    tdist.all_gather_object = all_gather_object

    def run(rank):
        try:
            tl.rank = rank
            lo, hi = bounds[rank], bounds[rank + 1]
            cb = [min(n, q * chunk) for q in range(world)] + [n]
            kern = Kern(F[lo:hi], lo, hi, cb)
            host_torch = types.SimpleNamespace(float64=torch.float64,
                                               zeros=lambda *a, **k: torch.zeros(*a, **{q: v for q, v in k.items() if q != "device"}))
            op = qd.PeerExchangeOperator(fake_qb, kern, n, rank, world, Comm(), host_torch)
            for b in range(2):                                       # poison everything that is not the own slice
                raw = (C.c_char * (chunk * world * 16)).from_address(op.X[b].ptr)
                np.frombuffer(raw, dtype=np.complex128)[:] = np.nan
            op.own(0).upload(np.ascontiguousarray(x[lo:hi]))
            y = np.zeros(hi - lo, dtype=np.complex128)
            op.matvec(0, y)
            ys[rank] = (y, op.schedule, op.skipped_blocks)
        except Exception as e:                                       # surface thread failures in the main thread
            errs.append((rank, repr(e)))
            try:
                barrier.abort()
            except Exception:
                pass

    threads = [threading.Thread(target=run, args=(r,), name=f"rank{r}") for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(60)
    tdist.all_gather_object = orig
    if saved is None:
        qd.os.environ.pop("QB_PEER_SCHEDULE", None)
    else:
        qd.os.environ["QB_PEER_SCHEDULE"] = saved
    assert not errs, errs
    return ys, lib.pulled, bounds


@pytest.mark.parametrize("schedule", ["matching", "ring"])
def test_peer_exchange_operator_logic_with_empty_blocks(schedule):
    """Four simulated ranks, a block structure in which rank r never touches the columns of rank (r+2)%4: the product is
    exact although those slices are poisoned with NaN (their blocks are skipped), the matching schedule does not even
    pull them, and the ring schedule pulls everything but multiplies only what exists."""
    import scipy.sparse as sp
    world, n = 4, 128
    rng = np.random.default_rng(5)
    dense = np.zeros((n, n), dtype=np.complex128)
    for r in range(world):
        for p in range(world):
            if (p - r) % world == 2:
                continue
            blk = (rng.random((32, 32)) < 0.15) * (rng.normal(size=(32, 32)) + 1j * rng.normal(size=(32, 32)))
            dense[32 * r:32 * r + 32, 32 * p:32 * p + 32] = blk
    F = sp.csr_matrix(dense)
    ys, pulled, bounds = _simulate_peer_exchange(world, n, F, schedule)
    x = (np.arange(n) + 1.0) * np.exp(0.3j * np.arange(n))
    ref = dense @ x
    for r in range(world):
        y, sched, skipped = ys[r]
        assert skipped == 1
        assert np.allclose(y, ref[bounds[r]:bounds[r + 1]], rtol=0, atol=1e-12) and not np.isnan(y).any()
        got = sorted(pulled[f"rank{r}"])
        if schedule == "matching":
            assert sched.startswith("matching") and got == sorted(p for p in range(world) if p != r and (p - r) % world != 2)
        else:
            assert sched == "ring" and got == sorted(p for p in range(world) if p != r)


# ------------------------------------------------------------------------------------------------------------------
# The same orchestration over shards of a species-order handle (csrc/species.cu): rows per owner are whole up
# configurations (dist.ROW_ALIGN = species_row_align(D_dn)) and the vectors are in the internal order.  The kernels are the
# oracle-backed ones above on P H P^T; what is under test is that nothing in the operators depends on the 32-row alignment.
def _species_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import scipy.sparse as sp
    import oracle_lib as O
    import species_builders as SB
    from quantum_basis_b200 import dist as qdist
    A, meta, ex = O.load_golden("hubbard4x2")                    # 4x2, N_up = N_dn = 4: D_up = D_dn = 70
    n = A.dim
    perm = SB.species_perm(8, 4, 4)
    P = sp.csr_matrix((np.ones(n), (perm, np.arange(n))), shape=(n, n))
    F = (P @ A.to_scipy_full() @ P.T).tocsr()
    d_dn = 70
    qdist.ROW_ALIGN = qdist.species_row_align(d_dn)
    bounds, chunk = qdist.equal_row_bounds(n, world)
    assert chunk % d_dn == 0 and chunk % 4 == 0 and all(b % d_dn == 0 for b in bounds)
    lo, hi = bounds[rank], bounds[rank + 1]
    kern = OracleKernels(F[lo:hi], lo, hi, n)
    x = np.empty(n, dtype=np.complex128); x[perm] = O.vec_randomize(n, 1)
    want = np.empty(n, dtype=np.complex128); want[perm] = ex["y1"]
    x_loc = kern.alloc(chunk)
    OracleKernels._c(x_loc)[: hi - lo] = x[lo:hi]
    errs = []
    op = qdist.ShardedOperator(kern, n, rank, world, qdist.TorchComm())
    y_loc = kern.alloc(chunk)
    op.matvec(x_loc, y_loc)
    errs.append(np.linalg.norm(OracleKernels._c(y_loc)[: hi - lo] - want[lo:hi]) / np.linalg.norm(want[lo:hi]))
    pop = qdist.PipelinedOperator(kern, n, rank, world, qdist.TorchComm())
    kern.set_col_bounds(pop.col_bounds)
    y2 = kern.alloc(chunk)
    pop.matvec(x_loc, y2)
    errs.append(np.linalg.norm(OracleKernels._c(y2)[: hi - lo] - want[lo:hi]) / np.linalg.norm(want[lo:hi]))
    state = torch.zeros(8, dtype=torch.float64); state[0] = 1.0
    a_dev = torch.zeros(64, dtype=torch.float64); b_dev = torch.zeros(64, dtype=torch.float64)
    qdist.pipelined_lanczos(pop, x_loc.clone(), kern.alloc(chunk), 64, 25, state, a_dev, b_dev)
    q.put((rank, max(errs), a_dev.numpy().copy(), b_dev.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_species_aligned_shards_over_gloo(world, oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_species_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    A, meta, ex = oracle.load_golden("hubbard4x2")
    for rank, err, a_s, b_s in res:
        assert err < 1e-13
        assert np.abs(a_s[:20] - ex["dn_a"][:20]).max() < 1e-10  # the compiled reference's coefficients: the order does not matter
        assert np.abs(b_s[1:20] - ex["dn_b"][1:20]).max() < 1e-10


def test_species_row_align():
    from quantum_basis_b200.dist import species_row_align, equal_row_bounds
    assert species_row_align(12870) == 25740 and species_row_align(70) == 140 and species_row_align(924) == 924
    b, chunk = equal_row_bounds(165636900, 8, align=species_row_align(12870))
    assert chunk % 12870 == 0 and b[-1] == 165636900 and all(x % 12870 == 0 for x in b)


# ---------------------------------------------------------------------------------------------------------------------
# Host-side logic of the native multi-GPU path (csrc/dist.cu behind quantum_basis_b200.dist.NativeDist): the shard bounds and
# the one-off exchange of the IPC handles.  The device side is covered on GPUs by scripts/dist_native_check.py.
def test_species_nnz_balanced_bounds():
    import lin_builders as LB
    import species_builders as SB
    from quantum_basis_b200 import dist as qd
    for (Lx, Ly, nu, nd) in ((4, 2, 4, 4), (4, 3, 6, 6), (3, 3, 4, 5)):
        ns, bonds = Lx * Ly, LB.square_bonds(Lx, Ly)
        Du, Dd = SB.configurations(ns, nu).size, SB.configurations(ns, nd).size
        lp, cp = SB.species_parts(ns, nu, nd, bonds) if Du * Dd < 200000 else (None, None)
        for world in (1, 2, 3, 8):
            b, dd = qd.species_nnz_balanced_bounds(ns, nu, nd, bonds, world)
            assert dd == Dd and b[0] == 0 and b[-1] == Du * Dd and len(b) == world + 1
            assert all(x % Dd == 0 for x in b) and all(b[k] <= b[k + 1] for k in range(world))
            if lp is not None and world > 1:
                nnz = (np.diff(lp.indptr) + np.diff(cp.indptr)).astype(np.int64)
                per = [int(nnz[b[k]:b[k + 1]].sum()) for k in range(world)]
                # balanced to within the entries of two up configurations
                assert max(per) - min(per) <= 2 * int(nnz.reshape(Du, Dd).sum(axis=1).max()), (world, per)
    w = np.arange(1 << 10)
    assert np.array_equal(qd.species_order_key(w, 10), SB._order_key(w, 10))


def _handle_exchange_worker(rank, world, port, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    mine = bytes([rank]) * 64                       # stands for the 64-byte CUDA IPC handle of this rank's buffers
    out = [None] * world
    dist.all_gather_object(out, mine)
    blob = b"".join(out)
    q.put((rank, len(blob), [blob[64 * p] for p in range(world)]))
    dist.destroy_process_group()


def test_ipc_handle_exchange_over_gloo():
    """what NativeDist does once at start-up: every rank contributes 64 bytes, every rank receives all of them rank-major"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 29641
    ps = [ctx.Process(target=_handle_exchange_worker, args=(r, world, port, q)) for r in range(world)]
    for p_ in ps:
        p_.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p_ in ps:
        p_.join(timeout=60)
    assert got == [(0, 128, [0, 1]), (1, 128, [0, 1])]
