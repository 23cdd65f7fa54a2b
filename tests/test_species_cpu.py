"""CPU tests of the species-order layouts of the Hubbard model (quantum_basis_b200/csrc/species.cu).

Two links, both without a GPU:
  1. the factorisation itself -- species order, the (local, cross) split of H, signs as (own configuration) x (parity of the
     other species on a site interval) -- restated in numpy (tests/species_builders.py) and compared entry for entry with
     tests/lin_builders.py, which is pinned bit for bit to matrices assembled by the compiled reference;
  2. the library's index logic -- the hop tables built in C++, the permutation from the reference's Lin order, the generator
     of the two stored parts, the slice order, and the two matrix-free passes with their warp-item decomposition -- executed
     on the host through qbgpu_debug_species_host (the same __host__ __device__ row functions the kernels call) and compared
     with the restatement.
What only the device can show (the kernels' launch geometry, reductions, timing) is left to tests/test_gpu_species.py.
"""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

import lin_builders as lb
import species_builders as sb
from quantum_basis_b200 import _lib

CASES = [  # Lx, Ly, nup, ndn   (4x2 has doubled y-bonds: multiplicity 2; 3x3 an odd site count)
    (4, 2, 3, 5), (3, 3, 4, 5), (4, 2, 4, 4), (2, 2, 1, 2), (3, 2, 3, 3), (3, 3, 2, 6),
]


def _full_from_upper(n, ia, ja, val):
    A = sp.csr_matrix((val, ja, ia), shape=(n, n))
    return (A + sp.triu(A, 1).conj().T).tocsr()


def _same(A, B):
    D = (A - B).tocsr()
    D.eliminate_zeros()
    return D.nnz == 0


@pytest.mark.parametrize("Lx,Ly,nup,ndn", CASES)
def test_species_split_equals_the_reference_matrix(Lx, Ly, nup, ndn):
    ns, bonds = Lx * Ly, lb.square_bonds(Lx, Ly)
    n, ia, ja, val = lb.hubbard_upper_csr(ns, nup, ndn, bonds, 1.0, 1.1)
    H = _full_from_upper(n, ia, ja, val.real)
    perm = sb.species_perm(ns, nup, ndn)
    assert sorted(perm.tolist()) == list(range(n))
    P = sp.csr_matrix((np.ones(n), (perm, np.arange(n))), shape=(n, n))
    local, cross = sb.species_parts(ns, nup, ndn, bonds, 1.0, 1.1)
    assert _same(P @ H @ P.T, local + cross)                      # exact: every value is +-t*w or U*k
    # the local part never leaves the block of one up configuration; the cross part never changes the down index
    Dd = sb.configurations(ns, ndn).size
    lc = local.tocoo()
    assert np.all(lc.row // Dd == lc.col // Dd)
    cc = cross.tocoo()
    assert np.all(cc.row % Dd == cc.col % Dd)


def _run_host(ns, nup, ndn, bonds, t, U, tile, x=None):
    L = _lib.lib()
    b = np.ascontiguousarray(np.asarray(bonds, dtype=np.int32).reshape(-1, 2))
    sizes = np.zeros(4, dtype=np.int64)
    null = C.c_void_p(0)
    p = lambda a: C.c_void_p(a.ctypes.data)   # noqa: E731
    _lib.check(L.qbgpu_debug_species_host(ns, nup, ndn, b.shape[0], p(b), t, U, tile, p(sizes), *([null] * 11)))
    Du, Dd, tu, td = (int(v) for v in sizes)
    n = Du * Dd
    out = dict(Du=Du, Dd=Dd, n=n, perm=np.empty(n, np.int32),
               rpl=np.empty(n + 1, np.int64), cl=np.empty(Du * (td + Dd), np.int32), vl=np.empty(Du * (td + Dd)),
               rpc=np.empty(n + 1, np.int64), cc=np.empty(Dd * tu, np.int32), vc=np.empty(Dd * tu),
               order=np.empty((n + 31) // 32, np.int32), y=np.zeros(n), touched=np.zeros(n, np.int32))
    xx = np.ascontiguousarray(x) if x is not None else np.zeros(n)
    _lib.check(L.qbgpu_debug_species_host(ns, nup, ndn, b.shape[0], p(b), t, U, tile, p(sizes), p(out["perm"]),
                                          p(out["rpl"]), p(out["cl"]), p(out["vl"]), p(out["rpc"]), p(out["cc"]), p(out["vc"]),
                                          p(out["order"]), p(xx), p(out["y"]), p(out["touched"])))
    return out


@pytest.mark.parametrize("Lx,Ly,nup,ndn", CASES)
def test_library_index_logic_on_the_host(Lx, Ly, nup, ndn):
    ns, bonds = Lx * Ly, lb.square_bonds(Lx, Ly)
    t, U = 1.0, 1.1
    rng = np.random.default_rng(7)
    n = sb.configurations(ns, nup).size * sb.configurations(ns, ndn).size
    x = rng.standard_normal(n)
    for tile in (32, 64):
        o = _run_host(ns, nup, ndn, bonds, t, U, tile, x)
        assert o["n"] == n
        # the permutation out of the reference's Lin order
        assert np.array_equal(o["perm"], sb.species_perm(ns, nup, ndn))
        # the two stored parts: same sparsity, same values, columns ascending inside every row
        local, cross = sb.species_parts(ns, nup, ndn, bonds, t, U)
        Ll = sp.csr_matrix((o["vl"], o["cl"], o["rpl"]), shape=(n, n))
        Lc = sp.csr_matrix((o["vc"], o["cc"], o["rpc"]), shape=(n, n))
        assert _same(Ll, local) and _same(Lc, cross)
        for rp, col in ((o["rpl"], o["cl"]), (o["rpc"], o["cc"])):
            inner = np.ones(col.size, dtype=bool)
            inner[rp[:-1][rp[:-1] < col.size]] = False             # first entry of each row
            assert np.all(np.diff(col.astype(np.int64))[inner[1:]] > 0)
        # the diagonal is stored in every row of the local part, even when it is zero (src/sparse.cc:44-54)
        rows = np.repeat(np.arange(n), np.diff(o["rpl"]))
        assert np.count_nonzero(rows == o["cl"]) == n
        # slice order: a permutation of the slices, tiles never decreasing, ascending inside a tile
        order = o["order"].astype(np.int64)
        assert sorted(order.tolist()) == list(range((n + 31) // 32))
        tiles = ((order * 32) % o["Dd"]) // tile
        assert np.all(np.diff(tiles) >= 0)
        assert np.all(np.diff(order)[np.diff(tiles) == 0] > 0)
        assert np.array_equal(order, sb.slice_order(o["Du"], o["Dd"], tile))
        # the two matrix-free passes: every row written exactly once by the cross pass, product equal to the matrix's
        assert np.all(o["touched"] == 1)
        y_ref = (local + cross) @ x
        assert np.linalg.norm(o["y"] - y_ref) <= 1e-13 * np.linalg.norm(y_ref)


def test_matrix_free_passes_reproduce_the_reference_order_product():
    """End to end on the host: x in the reference's order -> permute -> two passes -> permute back == H_ref x."""
    Lx, Ly, nup, ndn = 4, 2, 3, 5
    ns, bonds = Lx * Ly, lb.square_bonds(Lx, Ly)
    n, ia, ja, val = lb.hubbard_upper_csr(ns, nup, ndn, bonds, 1.0, 1.1)
    H = _full_from_upper(n, ia, ja, val.real)
    x = np.random.default_rng(3).standard_normal(n)
    perm = sb.species_perm(ns, nup, ndn)
    x_int = np.empty(n)
    x_int[perm] = x                                               # vec_to_native: dst[perm[r]] = src[r]
    o = _run_host(ns, nup, ndn, bonds, 1.0, 1.1, 32, x_int)
    y = o["y"][perm]                                              # vec_from_native: dst[r] = src[perm[r]]
    y_ref = H @ x
    assert np.linalg.norm(y - y_ref) <= 1e-13 * np.linalg.norm(y_ref)


@pytest.mark.parametrize("Lx,Ly,nup,ndn", [(4, 2, 3, 5), (3, 3, 4, 5), (4, 2, 4, 4)])
def test_row_shards_and_column_parts_on_the_host(Lx, Ly, nup, ndn):
    """The multi-GPU pattern of dist.py on the matrix-free species product: every rank owns a range of up configurations and
    multiplies column part p when the slice of rank p has arrived -- own part first (peer pulls) or in rank order (pipelined
    broadcasts, where a part WITHOUT the local pass opens the product).  Same __host__ __device__ functions as the kernels."""
    ns, bonds = Lx * Ly, lb.square_bonds(Lx, Ly)
    t, U = 1.0, 1.1
    local, cross = sb.species_parts(ns, nup, ndn, bonds, t, U)
    H = (local + cross).tocsr()
    Du, Dd = sb.configurations(ns, nup).size, sb.configurations(ns, ndn).size
    x = np.random.default_rng(5).standard_normal(Du * Dd)
    L = _lib.lib()
    b = np.ascontiguousarray(np.asarray(bonds, dtype=np.int32).reshape(-1, 2))
    p = lambda a: C.c_void_p(a.ctypes.data)   # noqa: E731
    for world in (2, 3, 5):
        chunk = -(-Du // world)
        pb = np.array([min(Du, q * chunk) for q in range(world)] + [Du], dtype=np.int64)
        for rank in range(world):
            u_lo, u_hi = int(pb[rank]), int(pb[rank + 1])
            want = (H @ x)[u_lo * Dd:u_hi * Dd]
            for order in ([rank] + [q for q in range(world) if q != rank], list(range(world)), list(range(world))[::-1]):
                o = np.array(order, dtype=np.int32)
                y = np.full(max(1, (u_hi - u_lo) * Dd), 7.0)                # must be overwritten, not accumulated into
                _lib.check(L.qbgpu_debug_species_parts_host(ns, nup, ndn, b.shape[0], p(b), t, U, 32, u_lo, u_hi, world, p(pb), p(o), p(x), p(y)))
                if u_hi > u_lo:
                    assert np.linalg.norm(y[:want.size] - want) <= 1e-13 * np.linalg.norm(want), (world, rank, order)


def test_tables_at_config3_scale_reproduce_the_stored_entry_count():
    """BASELINE config 3 (4x4, N_up = N_dn = 8): the two hop tables describe exactly the 5,819,376,420 entries of the expanded
    matrix -- D_up (dn hops + D_dn diagonals) + D_dn (up hops) -- and are built in a fraction of a second."""
    L = _lib.lib()
    b = np.ascontiguousarray(np.asarray(lb.square_bonds(4, 4), dtype=np.int32).reshape(-1, 2))
    sizes = np.zeros(4, dtype=np.int64)
    null = C.c_void_p(0)
    _lib.check(L.qbgpu_debug_species_host(16, 8, 8, b.shape[0], C.c_void_p(b.ctypes.data), 1.0, 1.1, 128, C.c_void_p(sizes.ctypes.data), *([null] * 11)))
    Du, Dd, tu, td = (int(v) for v in sizes)
    assert (Du, Dd) == (12870, 12870) and tu == td == 219648
    assert Du * (td + Dd) + Dd * tu == 5819376420


def test_bad_arguments_fail_loudly():
    L = _lib.lib()
    b = np.array([[0, 1]], dtype=np.int32)
    sizes = np.zeros(4, dtype=np.int64)
    null = C.c_void_p(0)
    assert L.qbgpu_debug_species_host(1, 1, 1, 1, C.c_void_p(b.ctypes.data), 1.0, 1.0, 32, C.c_void_p(sizes.ctypes.data), *([null] * 11)) != 0
    assert L.qbgpu_debug_species_host(26, 1, 1, 1, C.c_void_p(b.ctypes.data), 1.0, 1.0, 32, C.c_void_p(sizes.ctypes.data), *([null] * 11)) != 0
