#!/usr/bin/env python
"""oracle/make_repr_hashes.py -- TEST INFRASTRUCTURE ONLY.

Runs the UNMODIFIED reference (oracle/_ref/qb_ref, this container only) on a list of translation-symmetric sectors
(fill_Weisse_table + enumerate_basis_repr + generate_Ham_sparse_repr, src/model.cc:205-249, :275-487, :688-836) and
records dim, nnz and the SHA-256 of the raw ia / ja / val arrays of the assembled csr_mat in
tests/golden/repr_hashes.json.  tests/test_builders_cpu.py requires the numpy restatement (tests/repr_builders.py) to
reproduce every hash, i.e. to be bit-identical to the reference; the device builder is then checked against the numpy
restatement on the GPU box, where the reference is not available.
"""
import hashlib
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import oracle_lib as O  # noqa: E402

CASES = (
    [["heis_chain_k", 12, 0, k] for k in range(12)]
    + [["heis_chain_k", 14, 0, 0], ["heis_chain_k", 14, 0, 3], ["heis_chain_k", 14, 0, 7], ["heis_chain_k", 14, 1, 2]]
    + [["heis_chain_k", 16, 0, 0], ["heis_chain_k", 16, 0, 3], ["heis_chain_k", 16, 0, 8], ["heis_chain_k", 16, 2, 4]]
    + [["heis_chain_k", 20, 0, 0], ["heis_chain_k", 20, 0, 7], ["heis_chain_k", 22, 1, 5]]
    + [["tri_k", 4, 2, 0, 0, 0], ["tri_k", 4, 2, 0, 1, 1], ["tri_k", 4, 2, 0, 2, 0], ["tri_k", 2, 4, 0, 1, 1],
       ["tri_k", 2, 4, 0, 0, 2], ["tri_k", 4, 3, 0, 1, 2], ["tri_k", 4, 3, 0, 0, 0], ["tri_k", 3, 4, 0, 1, 1],
       ["tri_k", 3, 4, 0, 2, 3], ["tri_k", 6, 2, 1, 3, 1], ["tri_k", 6, 3, 0, 1, 2], ["tri_k", 4, 5, 0, 3, 2],
       ["tri_k", 4, 4, 0, 0, 0], ["tri_k", 4, 4, 0, 0, 1], ["tri_k", 4, 4, 0, 1, 2], ["tri_k", 4, 4, 1, 3, 3]]
    # single-orbital Hubbard model in momentum sectors (fermion signs of the translations): Lx Ly N_up N_dn t U m n
    + [["hubbard_k", 4, 2, 4, 4, 1.0, 1.1, 0, 0], ["hubbard_k", 4, 2, 4, 4, 1.0, 1.1, 1, 0], ["hubbard_k", 4, 2, 4, 4, 1.0, 1.1, 2, 1],
       ["hubbard_k", 4, 2, 3, 5, 1.0, 2.3, 1, 1], ["hubbard_k", 2, 4, 2, 3, 0.7, 1.9, 1, 2], ["hubbard_k", 4, 3, 2, 2, 1.0, 1.1, 1, 2],
       ["hubbard_k", 6, 2, 3, 3, 1.0, 4.0, 5, 1]]
)


def sha(a):
    """SHA-256 of the raw array; floating-point arrays with -0.0 folded into +0.0 first (x + 0.0), so that the digest
    pins every VALUE bit for bit without depending on the sign of a zero imaginary part"""
    import numpy as np
    a = np.ascontiguousarray(a)
    if a.dtype.kind == "c":
        a = (a.real + 0.0) + 1j * (a.imag + 0.0)
    elif a.dtype.kind == "f":
        a = a + 0.0
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    assert O.have_qb_ref(), "build oracle/_ref first: make -C oracle ref"
    out = []
    for args in CASES:
        wd = tempfile.mkdtemp(prefix="qbrepr_")
        csr = os.path.join(wd, "H.qbcsr")
        O.run_qb_ref(args + ["--dump", csr], threads=8, workdir=wd)
        A = O.read_qbcsr(csr)
        out.append({"qb_ref_args": args, "dim": A.dim, "nnz": A.nnz, "ia": sha(A.ia), "ja": sha(A.ja), "val": sha(A.val)})
        print(args, A.dim, A.nnz, flush=True)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "repr_hashes.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
