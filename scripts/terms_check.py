#!/usr/bin/env python
"""One-shot check of the term-coded matrix-free product (QBGPU_MATFREE_TERMS) against the walk kernel (bit-identical: same
entries in the same order) and the stored matrix (1e-13), complex and fp64 vectors, plus E0 from the fused Lanczos."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import quantum_basis_b200 as qb  # noqa: E402
import lin_builders as B  # noqa: E402

assert qb.lib().qbgpu_init(0) == 0
ok = True
cases = [("hubbard 4x3 (6,6)", lambda **k: qb.hubbard(12, 6, 6, B.square_bonds(4, 3), 1.0, 1.1, **k)),
         ("hubbard 4x2 (3,5)", lambda **k: qb.hubbard(8, 3, 5, B.square_bonds(4, 2), 1.0, 2.3, **k)),
         ("hubbard 3x3 (4,5)", lambda **k: qb.hubbard(9, 4, 5, B.square_bonds(3, 3), 0.7, 1.9, **k)),
         ("heisenberg chain 20", lambda **k: qb.heisenberg(20, 10, B.chain_bonds(20), 1.0, **k)),
         ("heisenberg chain 15 (7 down)", lambda **k: qb.heisenberg(15, 7, B.chain_bonds(15), 1.3, **k)),
         ("heisenberg square 4x4", lambda **k: qb.heisenberg(16, 8, B.square_bonds(4, 4), 1.0, **k))]
for name, mk in cases:
    for cplx in (True, False):
        Ms = mk(is_complex=cplx)
        Mw = mk(is_complex=cplx, matrix_free=True)
        Mt = mk(is_complex=cplx, matrix_free=True, flags=64)
        n = Ms.dim
        rng = np.random.default_rng(3)
        x = rng.normal(size=n) + (1j * rng.normal(size=n) if cplx else 0.0)
        x = x.astype(np.complex128 if cplx else np.float64)
        ys, yw, yt = (np.zeros_like(x) for _ in range(3))
        Ms.MultMv(x, ys); Mw.MultMv(x, yw); Mt.MultMv(x, yt)
        same = np.array_equal(yw, yt)
        rel = np.linalg.norm(yt - ys) / np.linalg.norm(ys)
        # the replay also multiplies its padding codes by zero, which could flip the sign of a zero: values must agree
        good = (same or np.abs(yw - yt).max() == 0.0) and rel < 1e-13
        ok &= good
        print(f"{name:30s} {'complex' if cplx else 'fp64   '} n={n:8d}  terms==walk bitwise: {same}  rel err vs stored: {rel:.2e}  {'ok' if good else 'FAIL'}", flush=True)
    Ms = mk(is_complex=True); Mt = mk(is_complex=True, matrix_free=True, flags=64)
    e_s = qb.locate_E0_lanczos(Ms, nev=1, ncv=0)["eigenvals"][0]
    e_t = qb.locate_E0_lanczos(Mt, nev=1, ncv=0)["eigenvals"][0]
    good = abs(e_s - e_t) < 1e-10 * max(1.0, abs(e_s))
    ok &= good
    print(f"{name:30s} E0 stored {e_s:.12f}  term-coded {e_t:.12f}  {'ok' if good else 'FAIL'}", flush=True)
print("ALL OK" if ok else "FAILURES")
sys.exit(0 if ok else 1)
