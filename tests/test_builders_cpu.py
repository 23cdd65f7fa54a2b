"""The numpy Lin-order builders (tests/lin_builders.py) pinned against CSR matrices assembled by the reference."""
import numpy as np
import pytest

import lin_builders as B


def test_hubbard4x2_is_bit_identical_to_reference(oracle):
    A, meta, ex = oracle.load_golden("hubbard4x2")
    n, ia, ja, val = B.hubbard_upper_csr(8, 4, 4, B.square_bonds(4, 2), t=1.0, U=1.1)
    assert n == A.dim and np.array_equal(ia, A.ia) and np.array_equal(ja, A.ja)
    assert np.array_equal(val, A.val)


@pytest.mark.parametrize("args,mk", [
    (["heis_chain", 12, "sz", 0], lambda: B.heisenberg_upper_csr(12, 6, B.chain_bonds(12))),
    (["heis_chain", 15, "sz", 0.5], lambda: B.heisenberg_upper_csr(15, 7, B.chain_bonds(15))),
    (["hubbard", 3, 3, 4, 5, 1, 2.3], lambda: B.hubbard_upper_csr(9, 4, 5, B.square_bonds(3, 3), U=2.3)),
])
def test_against_live_reference(oracle, args, mk, tmp_path):
    if not oracle.have_qb_ref():
        pytest.skip("oracle/_ref/qb_ref not built (needs /root/reference)")
    f = str(tmp_path / "H.qbcsr")
    oracle.run_qb_ref(args + ["--dump", f], workdir=str(tmp_path))
    A = oracle.read_qbcsr(f)
    n, ia, ja, val = mk()
    assert n == A.dim and np.array_equal(ia, A.ia) and np.array_equal(ja, A.ja) and np.array_equal(val, A.val)
