"""GPU tests of resuming the device Krylov loops (k > 0: qbgpu_lanczos_resume_*, qbgpu_eigenvec_cg_* entered with m > 0;
reference src/lanczos.cc:144-148, 287-292), of the reference-format checkpoints around them (quantum_basis_b200/ckpt.py; the
file protocol itself is covered on the CPU by test_ckpt_cpu.py), and of the opt-in fp64 route of MultMv."""
import os

import numpy as np
import pytest

import lin_builders as B
import species_builders as SB
import quantum_basis_b200 as qb
from quantum_basis_b200 import _lib
from gpu_species_common import SPECIES, TOL_MV, TOL_E0, TOL_KPM, CASES, rel_l2, _case

pytestmark = pytest.mark.gpu

def test_lanczos_resume_and_checkpoints(oracle, tmp_path):
    import ctypes as C
    from quantum_basis_b200 import ckpt
    A, meta, ex = oracle.load_golden("heis12_full")
    M = qb.csr_mat(A.dim, A.ia, A.ja, A.val, A.sym)
    n, maxit = A.dim, 200
    # uninterrupted
    v0, h0 = np.zeros(2 * n, dtype=np.complex128), np.zeros(2 * maxit)
    v0[:n] = oracle.vec_randomize(n, 1)
    m0 = qb.lanczos(0, maxit - 1, maxit, n, M, v0, h0, "sr_val0")
    assert m0 == meta["lanczos_steps"]
    # cut in two with the plain entry point: same coefficients; the stop rule restarts its counters, so it may run longer
    v1, h1 = np.zeros(2 * n, dtype=np.complex128), np.zeros(2 * maxit)
    v1[:n] = oracle.vec_randomize(n, 1)
    assert qb.lanczos(0, 20, maxit, n, M, v1, h1, "sr_val0") == 20
    m1 = qb.lanczos(20, maxit - 1 - 20, maxit, n, M, v1, h1, "sr_val0")
    assert m1 >= m0 and np.abs(h1[maxit:maxit + 30] - h0[maxit:maxit + 30]).max() < 1e-11 and np.abs(h1[:30] - h0[:30]).max() < 1e-11
    # cut in pieces WITH the stop rule's memory: stops at the same step
    L = qb.lib()
    v2, h2 = np.zeros(2 * n, dtype=np.complex128), np.zeros(2 * maxit)
    v2[:n] = oracle.vec_randomize(n, 1)
    st, m, k = (C.c_double * 4)(0, 0, 0, 0), C.c_int64(0), 0
    while True:
        _lib.check(L.qbgpu_lanczos_resume_z(M.handle, k, min(13, maxit - 1 - k), maxit, C.byref(m), C.c_void_p(v2.ctypes.data),
                                            C.c_void_p(h2.ctypes.data), b"sr_val0", 0, st))
        if m.value < k + 13:
            break
        k = m.value
    assert abs(m.value - m0) <= 1                                    # same stop step (the carried Ritz value comes from QL, the loop's from bisection)
    assert abs(qb.hess_eigen(h2, maxit, m.value)[0][0] - meta["lanczos_E0"]) <= TOL_E0 * abs(meta["lanczos_E0"])
    # the reference's files around it: interrupted after two pieces, then resumed; then resumed from the REFERENCE's checkpoint
    d = str(tmp_path / ckpt.DIRNAME)
    v3, h3 = np.zeros(2 * n, dtype=np.complex128), np.zeros(2 * maxit)
    v3[:n] = oracle.vec_randomize(n, 1)
    assert ckpt.lanczos_checkpointed(M, v3, h3, "sr_val0", maxit, every=10, dirpath=d, max_chunks=2) == 20
    v4, h4 = np.zeros(2 * n, dtype=np.complex128), np.zeros(2 * maxit)          # a fresh process would start like this
    m4 = ckpt.lanczos_checkpointed(M, v4, h4, "sr_val0", maxit, every=10, dirpath=d)
    assert abs(m4 - m0) <= 1
    assert abs(qb.hess_eigen(h4, maxit, m4)[0][0] - meta["lanczos_E0"]) <= TOL_E0 * abs(meta["lanczos_E0"])
    if oracle.have_qb_ref():
        w = str(tmp_path / "ref")
        path = str(tmp_path / "A.qbcsr")
        oracle.write_qbcsr(path, A)
        oracle.run_qb_ref(["file_z", path, "--lanczos-ckpt", "sr_val0", maxit, 20], workdir=w)
        v5, h5 = np.zeros(2 * n, dtype=np.complex128), np.zeros(2 * maxit)
        m5 = ckpt.lanczos_checkpointed(M, v5, h5, "sr_val0", maxit, every=50, dirpath=os.path.join(w, ckpt.DIRNAME))
        assert abs(m5 - m0) <= 1
        assert abs(qb.hess_eigen(h5, maxit, m5)[0][0] - meta["lanczos_E0"]) <= TOL_E0 * abs(meta["lanczos_E0"])


def test_cg_resume_and_checkpoints(oracle, tmp_path):
    """eigenvec_CG entered with m > 0 (src/lanczos.cc:287-292) and the reference's CG checkpoints around it."""
    from quantum_basis_b200 import ckpt
    A, meta, ex = oracle.load_golden("heis12_full")
    M = qb.csr_mat(A.dim, A.ia, A.ja, A.val, A.sym)
    n, E0 = A.dim, meta["lanczos_E0"]
    mk = lambda: [oracle.vec_randomize(n, 1)] + [np.zeros(n, dtype=np.complex128) for _ in range(3)]   # noqa: E731
    v, r, p, pp = mk()
    m_full, accu = qb.eigenvec_CG(n, 1000, 0, M, E0, v, r, p, pp)
    assert m_full == meta["cg_steps"] and accu < 2e-12
    v2, r2, p2, pp2 = mk()
    m, _ = qb.eigenvec_CG(n, 12, 0, M, E0, v2, r2, p2, pp2)          # stops at step 12 ...
    assert m == 12
    m, accu2 = qb.eigenvec_CG(n, 1000, m, M, E0, v2, r2, p2, pp2)    # ... and continues from (v, r, p)
    assert abs(m - m_full) <= 2 and accu2 < 2e-12 and rel_l2(v2, v) < 1e-8      # (the resumed piece runs on complex vectors, the first on fp64)
    d = str(tmp_path / ckpt.DIRNAME)
    v3, r3, p3, pp3 = mk()
    assert ckpt.cg_checkpointed(M, E0, v3, r3, p3, pp3, every=10, dirpath=d, max_chunks=2)[0] == 20
    v4, r4, p4, pp4 = [np.zeros(n, dtype=np.complex128) for _ in range(4)]          # a fresh process
    m4, accu4 = ckpt.cg_checkpointed(M, E0, v4, r4, p4, pp4, every=10, dirpath=d)
    assert abs(m4 - m_full) <= 2 and accu4 < 2e-12 and rel_l2(v4, v) < 1e-8


def test_opt_in_real_mode_of_the_plain_product(oracle):
    """QBGPU_MV_REAL_MODE: MultMv on complex device vectors without imaginary parts multiplies on fp64 copies (spmv.cu:
    mv_real_mode); same numbers as the complex kernel, and vectors WITH imaginary parts take the complex kernel."""
    A, meta, ex = oracle.load_golden("hubbard4x2")
    M = qb.csr_mat(A.dim, A.ia, A.ja, A.val, A.sym)
    n = A.dim
    xr = oracle.vec_randomize(n, 1)                                   # imag == 0, like every vector of the reference's flows
    xc = xr + 1j * oracle.vec_randomize(n, 2).real
    outs = {}
    for mode in ("off", "on"):
        if mode == "on":
            os.environ["QBGPU_MV_REAL_MODE"] = "1"
        try:
            for tag, x in (("real", xr), ("cplx", xc)):
                xd, yd = qb.DeviceVector.from_numpy(x), qb.DeviceVector.from_numpy(np.full(n, 2.0 + 0.0j))
                M.MultMv(xd, yd)
                y1 = yd.to_numpy()
                M.MultMv2(xd, yd)                                     # y += H x
                outs[(mode, tag)] = (y1, yd.to_numpy())
        finally:
            os.environ.pop("QBGPU_MV_REAL_MODE", None)
    for tag in ("real", "cplx"):
        for j in range(2):
            assert rel_l2(outs[("on", tag)][j], outs[("off", tag)][j]) <= 1e-15      # (bit-identical by construction; not required)
    assert rel_l2(outs[("on", "real")][0], ex["y1"]) <= TOL_MV


def test_outer_state_machine_resumes_every_stage_and_matches_the_uninterrupted_run(oracle, tmp_path):
    """model<T>::locate_E0_lanczos with enable_ckpt (src/model.cc:1124-1316, 2519-2746): E0 / V0 / E1 / V1 interrupted after
    every piece of 10 steps and continued by calling again -- same energies and vectors as the uninterrupted device run, the
    reference's state file and eigenvector files on disk at the end."""
    import struct
    from quantum_basis_b200 import ckpt
    A, meta, ex = oracle.load_golden("heis12_full")
    M = qb.csr_mat(A.dim, A.ia, A.ja, A.val, A.sym)
    n = A.dim
    ref = qb.locate_E0_lanczos(M, nev=2, ncv=2, maxit=300)
    d = str(tmp_path / ckpt.DIRNAME)
    calls, res = 0, None
    while True:
        res = ckpt.locate_E0_lanczos_checkpointed(M, nev=2, ncv=2, maxit=300, every=10, dirpath=d, max_chunks=1)
        calls += 1
        assert calls < 200
        if res["finished"]:
            break
    assert calls > 8                                                          # it really was interrupted in every stage
    assert abs(res["eigenvals"][0] - ref["eigenvals"][0]) < 1e-10 and abs(res["eigenvals"][1] - ref["eigenvals"][1]) < 1e-8
    assert abs(res["eigenvals"][0] - meta["lanczos_E0"]) < 1e-10              # the compiled reference's E0
    assert abs(abs(np.vdot(res["eigenvecs"][0], ref["eigenvecs"][0])) - 1.0) < 1e-6
    for j in (0, 1):                                                           # both are eigenvectors (E1 may be degenerate: no overlap test)
        y = np.zeros(n, dtype=np.complex128)
        M.MultMv(np.ascontiguousarray(res["eigenvecs"][j]), y)
        assert np.linalg.norm(y - res["eigenvals"][j] * res["eigenvecs"][j]) < (1e-8 if j == 0 else 1e-5)
        assert abs(np.linalg.norm(res["eigenvecs"][j]) - 1.0) < 1e-9
    assert abs(np.vdot(res["eigenvecs"][0], res["eigenvecs"][1])) < 1e-6
    f0 = os.path.join(d, "lczs_E0_sym0_sec0.Qckpt")
    flags = struct.unpack("<????qddd", open(f0, "rb").read())
    assert flags[:5] == (True, True, True, True, 2) and abs(flags[5] - res["eigenvals"][0]) < 1e-14
    v0 = qb.vec_disk_read(os.path.join(d, "eigenvec0_sym0_sec0.dat"), n, M.dtype)
    v1 = qb.vec_disk_read(os.path.join(d, "eigenvec1_sym0_sec0.dat"), n, M.dtype)
    assert v0 is not None and v1 is not None and abs(abs(np.vdot(v0, res["eigenvecs"][0])) - 1.0) < 1e-12
    assert sorted(x for x in os.listdir(d) if not x.startswith("eigenvec")) == ["lczs_E0_sym0_sec0.Qckpt"]      # stage checkpoints cleaned
    # a finished run is returned from the disk without any device work
    again = ckpt.locate_E0_lanczos_checkpointed(M, nev=2, ncv=2, maxit=300, every=10, dirpath=d, max_chunks=0)
    assert again["finished"] and np.array_equal(again["eigenvecs"][1], v1)
