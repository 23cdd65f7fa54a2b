"""GPU tests of the step-level entry points of SURVEY section 8(b): qbgpu_cg_restart / qbgpu_cg_step (the body of the reference's
eigenvec_CG loop, src/lanczos.cc:293-332, with the loop itself kept on the host) and qbgpu_cheb_step (one fused product of the
Chebyshev recurrence).  Checked against the oracle (residual of the eigenvector, the C restatement's moments) and against the
whole-loop entry points, which are made of the same pieces.

(First hardware run: the last GPU call of round 2, profiles/r02zk_pytest_gpu_full_suite_final.log.)
"""
import numpy as np
import pytest

import quantum_basis_b200 as qb

pytestmark = pytest.mark.gpu

TOL_KPM = 1e-9


def make(A):
    return qb.csr_mat(A.dim, A.ia, A.ja, A.val, A.sym)


@pytest.mark.parametrize("name", ["heis12_full", "hubbard4x2", "tri4x4_k00"])
def test_cg_loop_on_the_host_over_the_step_entries(oracle, name, monkeypatch):
    A, meta, ex = oracle.load_golden(name)
    M = make(A)
    n = A.dim
    E0 = meta["lanczos_E0"]
    x0 = oracle.vec_randomize(n, 1)
    # the step entries work on the handle's own element type (complex here): keep the whole loop on complex vectors too
    monkeypatch.setenv("QBGPU_NO_REAL_MODE", "1")
    dv = [qb.DeviceVector.from_numpy(x0)] + [qb.DeviceVector(n) for _ in range(3)]
    for w in dv[1:]:
        w.zero()
    m1, accu1 = qb.eigenvec_CG(n, 1000, 0, M, E0, *dv)
    v1 = dv[0].to_numpy()
    ds = [qb.DeviceVector.from_numpy(x0)] + [qb.DeviceVector(n) for _ in range(3)]
    for w in ds[1:]:
        w.zero()
    m2, accu2 = qb.eigenvec_CG_stepwise(n, 1000, 0, M, E0, *ds)
    v2 = ds[0].to_numpy()
    # the reference's own criteria (src/lanczos.cc:295-317): converged residual estimate, unit norm; and the eigen-residual
    assert accu2 < 2e-12
    assert abs(np.linalg.norm(v2) - 1.0) < 1e-9
    assert np.linalg.norm(oracle.spmv(A, v2) - E0 * v2) < 1e-9
    assert abs(m2 - meta["cg_steps"]) <= 5
    # the same pieces in the same order as the whole loop
    assert abs(m1 - m2) <= 1 and abs(accu1 - accu2) < 1e-12
    assert abs(np.vdot(v1, v2)) > 1 - 1e-9
    for w in dv + ds:
        w.free()


def test_cg_step_entries_reject_bad_arguments(oracle):
    A, meta, ex = oracle.load_golden("heis12_full")
    M = make(A)
    L = qb.lib()
    assert L.qbgpu_cg_step(M.handle, None, None, None, None, None, None, None) != 0
    assert L.qbgpu_cg_restart(M.handle, None, None, None, None, None, None, None) != 0
    assert L.qbgpu_cheb_step(M.handle, 1.0, -1.0, 1, None, None, None, None) != 0


@pytest.mark.parametrize("name", ["tri4x4_k01", "hubbard4x2", "heis16_k3"])
def test_cheb_step_loop_reproduces_the_moments(oracle, name):
    A, meta, ex = oracle.load_golden(name)
    M = make(A)
    phi = oracle.vec_randomize(A.dim, 3)
    lo, hi = meta["escale_lo"], meta["escale_hi"]
    dphi = qb.DeviceVector.from_numpy(phi)
    for nmom in (1, 2, 7, 64):
        mu = qb.kpm_moments_stepwise(M, dphi, lo, hi, nmom)
        ref = oracle.kpm_moments(A, phi, lo, hi, nmom)
        assert np.abs(mu - ref).max() <= TOL_KPM
        whole = qb.kpm_moments(M, phi, lo, hi, nmom)
        assert np.abs(mu - whole).max() <= 1e-12
    dphi.free()
