"""CPU test of the HOST logic above the step-level entry points (qbgpu_cg_restart / qbgpu_cg_step / qbgpu_cheb_step): the loops of
quantum_basis_b200.eigenvec_CG_stepwise and kpm_moments_stepwise run here against a stand-in for the library that does each
documented pass (include/qbgpu.h, "eigenvec_CG, one call per pass") in numpy on the oracle's matrix.  What is tested is the loop
structure the reference prescribes (src/lanczos.cc:293-332: when to restart, when to stop) and the indexing of the moments --
not the kernels: those are tests/test_gpu_zz_step_entries.py's job on the GPU.
"""
import ctypes as C

import numpy as np
import pytest

import quantum_basis_b200 as qb
from quantum_basis_b200 import csr as qcsr


def _arr(ptr, n, dtype):
    addr = ptr.value if isinstance(ptr, C.c_void_p) else int(ptr)
    nbytes = n * np.dtype(dtype).itemsize
    return np.frombuffer((C.c_char * nbytes).from_address(addr), dtype=dtype)


class FakeLib:
    """numpy stand-in for the entry points the two mirrors call; 'device' memory is host memory"""

    def __init__(self, oracle, A):
        self.oracle, self.A, self.n = oracle, A, A.dim
        self.bufs = {}
        self.calls = {"restart": 0, "step": 0, "cheb": 0}

    # memory
    def qbgpu_malloc(self, pref, nbytes):
        b = np.zeros(max(int(nbytes), 1), dtype=np.uint8)
        self.bufs[b.ctypes.data] = b
        pref._obj.value = b.ctypes.data
        return 0

    def qbgpu_free(self, p):
        self.bufs.pop(p.value, None)
        return 0

    def qbgpu_memcpy_h2d(self, d, s, nb):
        _arr(d, nb, np.uint8)[:] = _arr(s, nb, np.uint8)
        return 0

    qbgpu_memcpy_d2h = qbgpu_memcpy_h2d
    qbgpu_memcpy_d2d = qbgpu_memcpy_h2d

    def qbgpu_memset0(self, d, nb):
        _arr(d, nb, np.uint8)[:] = 0
        return 0

    def qbgpu_dznrm2(self, n, x, out):
        out._obj.value = float(np.linalg.norm(_arr(x, n, np.complex128)))
        return 0

    def qbgpu_last_error(self):
        return b"fake"

    # the documented passes
    def qbgpu_cg_restart(self, h, e, sc, v, r, p, vnorm, accu):
        n, E0 = self.n, complex(e[0], e[1])
        sc, v, r, p = _arr(sc, 8, np.float64), _arr(v, n, np.complex128), _arr(r, n, np.complex128), _arr(p, n, np.complex128)
        rn = float(np.linalg.norm(v))
        vnorm._obj.value = rn
        v /= rn
        r[:] = E0 * v - self.oracle.spmv(self.A, v)
        d = np.vdot(v, r)
        sc[1], sc[2], sc[3] = d.real, d.imag, float(np.vdot(r, r).real)
        p[:] = r
        sc[0] = np.sqrt(sc[3])
        accu._obj.value = float(sc[0])
        self.calls["restart"] += 1
        return 0

    def qbgpu_cg_step(self, h, e, sc, v, r, p, pp, accu):
        n, E0 = self.n, complex(e[0], e[1])
        sc = _arr(sc, 8, np.float64)
        v, r, p, pp = (_arr(a, n, np.complex128) for a in (v, r, p, pp))
        pp[:] = self.oracle.spmv(self.A, p) + (np.finfo(float).eps - E0) * p
        d = np.vdot(p, pp)
        sc[1], sc[2], sc[3] = d.real, d.imag, float(np.vdot(pp, pp).real)
        alpha = sc[0] ** 2 / d
        v += alpha * p
        r -= alpha * pp
        sc[4] = float(np.vdot(r, r).real)
        beta = np.sqrt(sc[4]) / sc[0]
        p[:] = r + beta * beta * p
        sc[5] = sc[0] * beta
        sc[0] = sc[5]
        accu._obj.value = float(sc[0])
        self.calls["step"] += 1
        return 0

    def qbgpu_cheb_step(self, h, lo, hi, first, cur, prev, nxt, dots):
        n = self.n
        cc, ss = 0.5 * (hi + lo), 0.5 * (hi - lo)
        cur, nxt_a = _arr(cur, n, np.complex128), _arr(nxt, n, np.complex128)
        ht = (self.oracle.spmv(self.A, cur) - cc * cur) / ss
        new = ht if first else 2.0 * ht - _arr(prev, n, np.complex128)
        nxt_a[:] = new
        if dots.value:
            d = _arr(dots, 3, np.float64)
            ip = np.vdot(cur, new)
            d[0], d[1], d[2] = ip.real, ip.imag, float(np.vdot(new, new).real)
        self.calls["cheb"] += 1
        return 0


class FakeMat:
    def __init__(self, A):
        self.dim, self.handle, self.is_complex, self.dtype = A.dim, None, True, np.dtype(np.complex128)


@pytest.fixture
def fake(oracle, monkeypatch):
    def install(A):
        L = FakeLib(oracle, A)
        monkeypatch.setattr(qcsr, "lib", lambda: L)
        return L
    return install


@pytest.mark.parametrize("name", ["heis12_full", "hubbard4x2"])
def test_cg_loop_follows_the_reference(oracle, fake, name):
    A, meta, ex = oracle.load_golden(name)
    L = fake(A)
    n, E0 = A.dim, meta["lanczos_E0"]
    M = FakeMat(A)
    vecs = [qb.DeviceVector.from_numpy(oracle.vec_randomize(n, 1))] + [qb.DeviceVector(n) for _ in range(3)]
    m, accu = qb.eigenvec_CG_stepwise(n, 1000, 0, M, E0, *vecs)
    v = vecs[0].to_numpy()
    assert accu < 2e-12
    assert abs(m - meta["cg_steps"]) <= 5                    # the reference's own run of eigenvec_CG on this matrix
    assert L.calls["restart"] >= 1 and L.calls["restart"] + L.calls["step"] == m
    assert abs(np.linalg.norm(v) - 1.0) < 1e-9
    assert np.linalg.norm(oracle.spmv(A, v) - E0 * v) < 1e-9
    assert abs(np.vdot(ex["cg_vec"], v)) > 1 - 1e-8
    # maxit cuts the loop like the reference's while (m < maxit)
    vecs2 = [qb.DeviceVector.from_numpy(oracle.vec_randomize(n, 1))] + [qb.DeviceVector(n) for _ in range(3)]
    m2, accu2 = qb.eigenvec_CG_stepwise(n, 7, 0, M, E0, *vecs2)
    assert m2 == 7 and accu2 > 2e-12


@pytest.mark.parametrize("name", ["tri4x4_k01", "hubbard4x2"])
def test_cheb_loop_indexes_the_moments_like_the_whole_loop(oracle, fake, name):
    A, meta, ex = oracle.load_golden(name)
    L = fake(A)
    phi = oracle.vec_randomize(A.dim, 3)
    lo, hi = meta["escale_lo"], meta["escale_hi"]
    dphi = qb.DeviceVector.from_numpy(phi)
    for nmom in (1, 2, 3, 7, 64):
        L.calls["cheb"] = 0
        mu = qb.kpm_moments_stepwise(FakeMat(A), dphi, lo, hi, nmom)
        assert L.calls["cheb"] == nmom // 2 + 1               # one product per two moments
        assert np.abs(mu - oracle.kpm_moments(A, phi, lo, hi, nmom)).max() <= 1e-9
