"""GPU tests of the species-order handles of the Hubbard model (QBGPU_SPECIES_ORDER; quantum_basis_b200/csrc/species.cu).

First hardware run: round 2 (gpurun_out/r02_pytest_species.log, all green).  The index logic (hop tables, permutation,
generator of the two stored parts, slice order, the matrix-free passes) is also checked on the host by
tests/test_species_cpu.py through the same __host__ __device__ functions.

What is compared: the species-order product, through the reference-shaped calls (vectors in the reference's order), against
the CPU restatement of the reference's product on the reference-identical matrix (1e-12, BASELINE.json) and against the
ordinary device handle; the Krylov drivers (E0 to 1e-10, KPM moments to 1e-9) against the ordinary handle.
"""
import os

import numpy as np
import pytest

import lin_builders as B
import species_builders as SB
import quantum_basis_b200 as qb
from quantum_basis_b200 import _lib
from gpu_species_common import SPECIES, TOL_MV, TOL_E0, TOL_KPM, CASES, rel_l2, _case

pytestmark = pytest.mark.gpu

@pytest.mark.parametrize("name", ["hub4x2_35", "hub3x3_45"])
@pytest.mark.parametrize("matrix_free", [False, True])
def test_permutation_and_info(name, matrix_free):
    ns, nu, nd, bonds, U, mk = _case(name)
    M = mk(mf=matrix_free)
    assert M.has_internal_order()
    assert np.array_equal(M.native_perm(), SB.species_perm(ns, nu, nd))
    plain = mk(flags=0)
    assert not plain.has_internal_order()
    inf = M.info
    assert inf.n == plain.dim
    if matrix_free:
        assert inf.format == 32 and inf.nnz_stored == 0
    else:
        assert inf.nnz_stored == plain.info.nnz_stored and inf.nnz_input == plain.info.nnz_input    # same entries, split in two parts


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("matrix_free", [False, True])
def test_product_in_the_reference_order(oracle, name, matrix_free):
    from oracle_lib import Csr
    ns, nu, nd, bonds, U, mk = _case(name)
    n, ia, ja, val = B.hubbard_upper_csr(ns, nu, nd, bonds, 1.0, U)
    A = Csr(n, ia, ja, val, True)
    rng = np.random.default_rng(11)
    x = rng.normal(size=n) + 1j * rng.normal(size=n)
    M = mk(mf=matrix_free)
    y = np.full(n, 3.0 - 2.0j)
    M.MultMv(x, y)                                                    # host vectors, complex
    assert rel_l2(y, oracle.spmv_ld(A, x)) <= TOL_MV
    ys = np.zeros(n, dtype=np.complex128)
    mk(flags=0).MultMv(x, ys)
    assert rel_l2(y, ys) <= 1e-14
    y2 = y.copy()
    M.MultMv2(x, y2)                                                  # y += H x
    assert rel_l2(y2, 2 * y) <= 1e-14
    xd, yd = qb.DeviceVector.from_numpy(x), qb.DeviceVector(n)        # device vectors
    M.MultMv(xd, yd)
    assert np.array_equal(yd.to_numpy(), y)
    Md = mk(cx=False, mf=matrix_free)                                 # fp64 handle
    xr = rng.normal(size=n); yr = np.zeros(n)
    Md.MultMv(xr, yr)
    assert rel_l2(yr, oracle.spmv_ld(A, xr).real) <= TOL_MV


@pytest.mark.parametrize("matrix_free", [False, True])
def test_internal_order_entry_points(matrix_free):
    """to_native / from_native and the fused product (alpha, gamma, beta*z, running dots) in the internal order."""
    import ctypes as C
    ns, nu, nd, bonds, U, mk = _case("hub4x2_44")
    M, P = mk(mf=matrix_free), mk(flags=0)
    n = M.dim
    perm = M.native_perm()
    rng = np.random.default_rng(5)
    x = rng.normal(size=n) + 1j * rng.normal(size=n)
    z = rng.normal(size=n) + 1j * rng.normal(size=n)
    xd = qb.DeviceVector.from_numpy(x)
    xn = M.to_native(xd)
    x_int = np.empty_like(x); x_int[perm] = x
    assert np.array_equal(xn.to_numpy(), x_int)
    assert np.array_equal(M.from_native(xn).to_numpy(), x)
    # y = alpha H x + gamma x + beta z with dots, internal order, against the ordinary handle in the reference's order
    L = qb.lib()
    z_int = np.empty_like(z); z_int[perm] = z
    zn, yn = qb.DeviceVector.from_numpy(z_int), qb.DeviceVector(n)
    dots = qb.DeviceVector(4, np.float64)
    al, ga, be = (C.c_double * 2)(0.7, 0.0), (C.c_double * 2)(-0.3, 0.0), (C.c_double * 2)(1.5, 0.0)
    _lib.check(L.qbgpu_spmv_fused(M.handle, C.c_void_p(xn.ptr), C.c_void_p(zn.ptr), C.c_void_p(yn.ptr), al, ga, be, C.c_void_p(dots.ptr)))
    hx = np.zeros(n, dtype=np.complex128)
    P.MultMv(x, hx)
    want = 0.7 * hx - 0.3 * x + 1.5 * z
    got = yn.to_numpy()[perm]
    assert rel_l2(got, want) <= 1e-13
    d = dots.to_numpy()
    assert abs(d[0] + 1j * d[1] - np.vdot(x, want)) <= 1e-11 * abs(np.vdot(x, want)) + 1e-11
    assert abs(d[2] - np.vdot(want, want).real) <= 1e-12 * np.vdot(want, want).real
    # in place (z aliases y), the Lanczos / Chebyshev calling pattern
    _lib.check(L.qbgpu_spmv_fused(M.handle, C.c_void_p(xn.ptr), C.c_void_p(zn.ptr), C.c_void_p(zn.ptr), al, ga, be, None))
    assert rel_l2(zn.to_numpy()[perm], want) <= 1e-13


@pytest.mark.parametrize("matrix_free", [False, True])
def test_krylov_drivers_match_the_ordinary_handle(oracle, matrix_free):
    from oracle_lib import Csr
    ns, nu, nd, bonds, U, mk = _case("hub3x3_45")
    n, ia, ja, val = B.hubbard_upper_csr(ns, nu, nd, bonds, 1.0, U)
    A = Csr(n, ia, ja, val, True)
    M, P = mk(mf=matrix_free), mk(flags=0)
    # Lanczos coefficients from the reference's start vector (they do not depend on the order of the basis)
    hs, hp = np.zeros(200), np.zeros(200)
    vs, vp = np.zeros(2 * n, dtype=np.complex128), np.zeros(2 * n, dtype=np.complex128)
    vs[:n] = oracle.vec_randomize(n, 1); vp[:n] = vs[:n]
    ms = qb.lanczos(0, 40, 100, n, M, vs, hs, "dnmcs")
    mp = qb.lanczos(0, 40, 100, n, P, vp, hp, "dnmcs")
    assert ms == mp == 40
    assert np.abs(hs[100:120] - hp[100:120]).max() <= 1e-10 and np.abs(hs[1:21] - hp[1:21]).max() <= 1e-10
    # live vectors, reference order (round-off of two summation orders is amplified along the recursion: a loose bound)
    assert rel_l2(vs[:n], vp[:n]) <= 1e-5 and rel_l2(vs[n:], vp[n:]) <= 1e-5
    assert abs(np.linalg.norm(vs[:n]) - 1) < 1e-12 and abs(np.linalg.norm(vs[n:]) - 1) < 1e-12
    # E0, ground state, E1
    out_s = qb.locate_E0_lanczos(M, nev=2, ncv=2)
    out_p = qb.locate_E0_lanczos(P, nev=2, ncv=2)
    for k in range(2):
        tol = TOL_E0 if k == 0 else 1e-8                              # E1 rests on the CG vector of E0: looser
        assert abs(out_s["eigenvals"][k] - out_p["eigenvals"][k]) <= tol * abs(out_p["eigenvals"][k])
        v = out_s["eigenvecs"][k]
        assert np.linalg.norm(oracle.spmv(A, v) - out_s["eigenvals"][k] * v) < 1e-6
    # spectral bounds and Chebyshev moments
    ws, wp = np.zeros(2 * n, dtype=np.complex128), np.zeros(2 * n, dtype=np.complex128)
    lo_s, hi_s = qb.energy_scale(n, M, ws)
    lo_p, hi_p = qb.energy_scale(n, P, wp)
    assert abs(lo_s - lo_p) <= 1e-9 * abs(lo_p) and abs(hi_s - hi_p) <= 1e-9 * abs(hi_p)
    phi = oracle.vec_randomize(n, 3)
    mu_s = qb.kpm_moments(M, phi, lo_p, hi_p, 64)
    mu_p = qb.kpm_moments(P, phi, lo_p, hi_p, 64)
    assert np.abs(mu_s - mu_p).max() <= TOL_KPM
    # thick-restart Lanczos: eigenvalues and Ritz vectors in the reference's order
    nconv, w, Uv, nprod = qb.trlan(M, 2, 8, 400)
    assert nconv >= 2 and abs(w[0] - out_p["eigenvals"][0]) <= 1e-9 * abs(w[0])
    assert np.linalg.norm(oracle.spmv(A, Uv[:, 0].copy()) - w[0] * Uv[:, 0]) < 1e-6


@pytest.mark.parametrize("tile", ["32", "256"])
def test_midsize_sector_and_tile_widths(tile):
    """Hubbard 4x3, N_up = N_dn = 6 (853,776 states): both handle kinds against the ordinary device handle, with tiles
    narrower and wider than the default (the tile width is read when the handle is created)."""
    ns, bonds = 12, B.square_bonds(4, 3)
    os.environ["QBGPU_SPECIES_TILE"] = tile
    try:
        Ms = qb.hubbard(ns, 6, 6, bonds, 1.0, 1.1, flags=SPECIES)
        Mf = qb.hubbard(ns, 6, 6, bonds, 1.0, 1.1, flags=SPECIES, matrix_free=True)
    finally:
        del os.environ["QBGPU_SPECIES_TILE"]
    P = qb.hubbard(ns, 6, 6, bonds, 1.0, 1.1)
    n = P.dim
    assert n == 853776
    x = qb.vec_randomize(n, 1, device=True)
    yp, ys, yf = qb.DeviceVector(n), qb.DeviceVector(n), qb.DeviceVector(n)
    P.MultMv(x, yp); Ms.MultMv(x, ys); Mf.MultMv(x, yf)
    ref = yp.to_numpy()
    assert rel_l2(ys.to_numpy(), ref) <= 1e-14 and rel_l2(yf.to_numpy(), ref) <= 1e-14
    L = qb.lib()
    try:                                                              # the three pass-1 variants of the matrix-free product
        for vid in (1, 2, 0):
            assert L.qbgpu_debug_set_variant(1000 + vid) == 0
            yf.zero()
            Mf.MultMv(x, yf)
            assert rel_l2(yf.to_numpy(), ref) <= 1e-14, f"pass-1 variant {vid}"
    finally:
        L.qbgpu_debug_set_variant(1000)
    e_p = qb.locate_E0_lanczos(P, nev=1, ncv=0)["eigenvals"][0]
    for M in (Ms, Mf):
        e = qb.locate_E0_lanczos(M, nev=1, ncv=0)["eigenvals"][0]
        assert abs(e - e_p) <= TOL_E0 * abs(e_p)


def test_unsupported_uses_fail_loudly():
    ns, nu, nd, bonds, U, mk = _case("hub4x2_35")
    with pytest.raises(qb.QbgpuError):
        qb.hubbard(ns, nu, nd, bonds, 1.0, U, flags=SPECIES, rows=(0, 100))           # no row shards
    with pytest.raises(qb.QbgpuError):
        qb.heisenberg(12, 6, B.chain_bonds(12), 1.0, flags=SPECIES)                   # one species only
    M = mk()
    with pytest.raises(qb.QbgpuError):
        M.to_dense()
    with pytest.raises(qb.QbgpuError):
        M.download_expanded()


@pytest.mark.parametrize("world", [2, 3])
def test_matrix_free_row_shards_and_column_parts(world):
    """Multi-GPU building blocks of the matrix-free species product on one device: every 'rank' holds a shard of whole up
    configurations, split into one column part per owner; multiplying the parts in the exchange's order (own part first, or
    rank order) reproduces the rows of the full product.  Vectors are in the internal order (shards have no permutation)."""
    import ctypes as C
    from quantum_basis_b200 import dist
    ns, nu, nd, bonds, U, mk = _case("hub3x3_45")
    full = mk(mf=True)
    n = full.dim
    Du = SB.configurations(ns, nu).size
    Dd = n // Du
    rng = np.random.default_rng(8)
    x = rng.normal(size=n) + 1j * rng.normal(size=n)
    xd, yd = qb.DeviceVector.from_numpy(x), qb.DeviceVector(n)
    L = qb.lib()
    one, zero = (C.c_double * 2)(1.0, 0.0), (C.c_double * 2)(0.0, 0.0)
    _lib.check(L.qbgpu_spmv_fused(full.handle, C.c_void_p(xd.ptr), None, C.c_void_p(yd.ptr), one, zero, zero, None))   # internal order
    want = yd.to_numpy()
    chunk = -(-Du // world) * Dd
    bounds = [min(n, q * chunk) for q in range(world)] + [n]
    for rank in range(world):
        lo, hi = bounds[rank], bounds[rank + 1]
        S = qb.hubbard(ns, nu, nd, bonds, 1.0, U, matrix_free=True, flags=SPECIES, rows=(lo, hi))
        assert not S.has_internal_order() and S.info.row_lo == lo and S.info.row_hi == hi
        ys = qb.DeviceVector.from_numpy(np.full(hi - lo, 5.0 + 1.0j))
        S.MultMv(xd, ys)                                              # the unsplit shard
        assert rel_l2(ys.to_numpy(), want[lo:hi]) <= 1e-14
        parts = dist.DeviceKernels.split(qb, S, bounds)
        for order in ([rank] + [q for q in range(world) if q != rank], list(range(world))):
            yp = qb.DeviceVector.from_numpy(np.full(hi - lo, -3.0 + 2.0j))
            for k, p in enumerate(order):
                b = (C.c_double * 2)(0.0 if k == 0 else 1.0, 0.0)
                _lib.check(L.qbgpu_zmv(parts[p].handle, one, C.c_void_p(xd.ptr), b, C.c_void_p(yp.ptr), 1))
            assert rel_l2(yp.to_numpy(), want[lo:hi]) <= 1e-14, (rank, order)
    with pytest.raises(qb.QbgpuError):
        qb.hubbard(ns, nu, nd, bonds, 1.0, U, matrix_free=True, flags=SPECIES, rows=(0, Dd + 1))   # not whole up configurations
