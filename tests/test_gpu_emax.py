"""GPU test of model<T>::locate_Emax_iram (src/model.cc:1370-1422) over the host-ARPACK seam and the device-resident
thick-restart Lanczos on -H."""
import os

import numpy as np
import pytest

import lin_builders as B
import species_builders as SB
import quantum_basis_b200 as qb
from quantum_basis_b200 import _lib
from gpu_species_common import SPECIES, TOL_MV, TOL_E0, TOL_KPM, CASES, rel_l2, _case

pytestmark = pytest.mark.gpu

@pytest.mark.parametrize("name", ["tri4x4_k01", "hubbard4x2"])
def test_locate_Emax_iram(oracle, name):
    """model<T>::locate_Emax_iram (src/model.cc:1370-1422): the highest eigenvalues through the host-ARPACK seam and through
    the device-resident thick-restart Lanczos on -H; Emax skips the artificial states of zero-norm representatives."""
    A, meta, ex = oracle.load_golden(name)
    M = qb.csr_mat(A.dim, A.ia, A.ja, A.val, A.sym)
    w = np.linalg.eigvalsh(M.to_dense())[::-1]
    for dev in (False, True):
        out = qb.locate_Emax_iram(M, nev=2, ncv=10, maxit=400, device_resident=dev)
        assert out["nconv"] >= 1
        assert abs(out["eigenvals"][0] - w[0]) <= 1e-9 * abs(w[0])
        want = next(e for e in w if e < 100.0)
        if out["Emax"] < 100.0:                                       # reached only when an eigenvalue below fake_pos is among the nev
            assert abs(out["Emax"] - want) <= 1e-8 * abs(want)
        v = out["eigenvecs"][0]
        assert np.linalg.norm(oracle.spmv(A, v.copy()) - out["eigenvals"][0] * v) < 1e-6
