"""numpy restatement of the orbit assembler (quantum_basis_b200/csrc/orbit.cu) -- TEST INFRASTRUCTURE ONLY.

BASELINE config 4 (Heisenberg on the tilted 31-site triangular cluster in a momentum sector) cannot be produced by the
reference (src/lattice.cc:1079 asserts !q_tilted(), src/model.cc:811 divides by L[d]), so this row has no
reference oracle: "parity unpinned" against the reference, by necessity.  What stands in for it:
  * this independent restatement of the textbook convention (representative = smallest bit pattern of the orbit,
    <r'_k|H|r_k> = sum_b h_b conj(chi(g_b)) sqrt(|Stab r'|/|Stab r|)), checked ON THE CPU by a property no convention
    error survives: the spectra of all momentum sectors together are the spectrum of the full Sz sector, whose matrix
    comes from tests/lin_builders.py (bit-identical to the reference's full-basis assembler);
  * the device assembler against this restatement, entry for entry (tests/test_gpu_parity.py);
  * on untilted clusters, a sector against the reference-convention sector of sectors.cu (same spectrum).
"""
import numpy as np


def _apply(states, perm):
    out = np.zeros_like(states)
    for s, p in enumerate(perm):
        out |= ((states >> np.uint64(s)) & np.uint64(1)) << np.uint64(int(p))
    return out


def _states_with_popcount(nsites, k):
    """all bit patterns with k of nsites bits set, ascending"""
    from itertools import combinations
    if nsites <= 22:
        a = np.arange(1 << nsites, dtype=np.uint64)
        pc = np.zeros(a.size, dtype=np.int64)
        t = a.copy()
        while np.any(t):
            pc += (t & np.uint64(1)).astype(np.int64)
            t >>= np.uint64(1)
        return a[pc == k]
    return np.array(sorted(sum(1 << i for i in c) for c in combinations(range(nsites), k)), dtype=np.uint64)


def heisenberg_orbit_upper_csr(nsites, ndown, perms, chi, bonds, J=1.0):
    """(reps, stab, ia, ja, val): representatives (ascending), stabiliser sizes and the upper-triangle CSR of the sector."""
    perms = np.asarray(perms)
    chi = np.asarray(chi, dtype=np.complex128)
    ntrans = perms.shape[0]
    st = _states_with_popcount(nsites, ndown)
    imgs = np.stack([_apply(st, perms[t]) for t in range(ntrans)])
    is_min = np.all(imgs >= st[None, :], axis=0)
    fixed = imgs == st[None, :]
    trivial = np.abs(chi - 1.0) < 1e-9
    ok = is_min & ~np.any(fixed & ~trivial[:, None], axis=0)
    reps = st[ok]
    stab = fixed[:, ok].sum(axis=0).astype(np.float64)
    n = reps.size
    u = np.uint64
    diag = np.zeros(n)
    for (p, q) in bonds:
        par = ((reps >> u(p)) & u(1)) == ((reps >> u(q)) & u(1))
        diag = diag + np.where(par, 0.25 * J, -0.25 * J)
    R, C, V, O = [np.arange(n)], [np.arange(n)], [diag.astype(np.complex128)], [np.full(n, -1)]
    for t, (p, q) in enumerate(bonds):
        act = np.nonzero(((reps >> u(p)) & u(1)) != ((reps >> u(q)) & u(1)))[0]
        if act.size == 0:
            continue
        s2 = reps[act] ^ u((1 << p) | (1 << q))
        im2 = np.stack([_apply(s2, perms[g]) for g in range(ntrans)])
        g = np.argmin(im2, axis=0)                                  # first group element reaching the minimum
        best = im2[g, np.arange(act.size)]
        j = np.searchsorted(reps, best)
        j = np.minimum(j, n - 1)
        valid = (reps[j] == best) & (j >= act)
        act, j, g = act[valid], j[valid], g[valid]
        w = 0.5 * J * np.sqrt(stab[j] / stab[act])
        R.append(act); C.append(j); V.append(w * chi[g].real + 1j * (w * chi[g].imag)); O.append(np.full(act.size, t))
    R, C, V, O = (np.concatenate(z) for z in (R, C, V, O))
    order = np.lexsort((O, C, R))
    R, C, V = R[order], C[order], V[order]
    first = np.ones(R.size, dtype=bool)
    first[1:] = (R[1:] != R[:-1]) | (C[1:] != C[:-1])
    gid = np.cumsum(first) - 1
    rank = np.arange(R.size) - np.nonzero(first)[0][gid]
    acc = np.zeros(int(gid[-1]) + 1, dtype=np.complex128)
    for t in range(int(rank.max()) + 1):                            # sequential accumulation in bond order
        sel = rank == t
        acc[gid[sel]] = acc[gid[sel]] + V[sel]
    gr, gc = R[first], C[first]
    acc[gr == gc] = acc[gr == gc].real                               # the diagonal of a Hermitian matrix is real
    ia = np.zeros(n + 1, dtype=np.int64)
    np.add.at(ia, gr + 1, 1)
    return reps, stab, np.cumsum(ia), gc.astype(np.int64), acc
