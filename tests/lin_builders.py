"""numpy restatement of the reference's full-basis assembly for two model families, in Lin-table order.

TEST INFRASTRUCTURE ONLY (an oracle-side builder): it restates, in vectorised numpy,
  - the basis order: sort by (label of odd sites, label of even sites), reference src/basis.cc:1144-1190 with the labels
    of mbasis_elem::label_sub (src/basis.cc:428-450),
  - model::generate_Ham_sparse_full (src/model.cc:619-686) for the spin-1/2 Heisenberg and the single-orbital
    Fermi-Hubbard models, with the fermion sign rule of oprXphi (src/basis.cc:2717-2731),
and is pinned against CSR matrices assembled by the compiled reference (tests/golden: heis12_full, heis16_full,
hubbard4x2).  It returns the reference's upper-triangle csr_mat arrays, so the device generators of libqbgpu can be
checked at sizes where running the reference itself is too slow for a test.
"""
import numpy as np


def chain_bonds(L):
    return [(x, (x + 1) % L) for x in range(L)]


def square_bonds(Lx, Ly):
    """Bond list exactly as examples/trans_absent/latt_square/square_Fermi_Hubbard.cc:47-90 generates it for PBC:
    site = x + y*Lx (lattice::coor2site, src/lattice.cc), one +x and one +y bond per site (duplicates kept)."""
    site = lambda x, y: (x % Lx) + (y % Ly) * Lx   # noqa: E731
    b = []
    for x in range(Lx):
        for y in range(Ly):
            b.append((site(x, y), site(x + 1, y)))
            b.append((site(x, y), site(x, y + 1)))
    return b


def _split_labels(states, nsites, bps):
    """states: integer array with `bps` bits per site (site 0 lowest) -> (label_a, label_b)."""
    la = np.zeros_like(states)
    lb = np.zeros_like(states)
    mask = (1 << bps) - 1
    for s in range(nsites):
        d = (states >> (bps * s)) & mask
        if s % 2 == 0:
            la |= d << (bps * (s // 2))
        else:
            lb |= d << (bps * (s // 2))
    return la, lb


def _popcount(a):
    a = a.astype(np.uint64)
    c = np.zeros(a.shape, dtype=np.int64)
    while np.any(a):
        c += (a & np.uint64(1)).astype(np.int64)
        a >>= np.uint64(1)
    return c


def _enumerate(nsites, bps, counts):
    """All basis states with the conserved counts (None: no restriction), in Lin order."""
    if bps == 1:
        allst = np.arange(1 << nsites, dtype=np.int64)
        keep = (_popcount(allst) == counts[0]) if counts[0] is not None else np.ones(allst.size, dtype=bool)
    else:
        allst = np.arange(1 << (2 * nsites), dtype=np.int64)
        up = np.int64(int("01" * nsites, 2))
        keep = (_popcount(allst & up) == counts[0]) & (_popcount(allst & (up << 1)) == counts[1])
    st = allst[keep]
    la, lb = _split_labels(st, nsites, bps)
    order = np.lexsort((la, lb))                      # primary key lb, secondary la
    return st[order]


def basis_states(nsites, bps, counts):
    """Basis states (bit patterns, site 0 lowest, `bps` bits per site) in the reference's Lin order."""
    return _enumerate(nsites, bps, counts)


def _index_of(sorted_states_by_value, perm, targets):
    pos = np.searchsorted(sorted_states_by_value, targets)
    return perm[pos]


def _to_upper_csr(n, rows, cols, vals):
    keep = cols >= rows
    rows, cols, vals = rows[keep], cols[keep], vals[keep]
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    # accumulate duplicates in encounter order is not needed: callers pre-merge equal (row, col) pairs
    ia = np.zeros(n + 1, dtype=np.int64)
    np.add.at(ia, rows + 1, 1)
    ia = np.cumsum(ia)
    return ia, cols.astype(np.int64), vals


def heisenberg_upper_csr(nsites, ndown, bonds, J=1.0, dtype=np.complex128):
    st = _enumerate(nsites, 1, (ndown,))
    n = st.size
    byval = np.argsort(st)
    sv = st[byval]
    rows = [np.arange(n)]
    cols = [np.arange(n)]
    diag = np.zeros(n)
    offs = {}
    for (i, j) in bonds:
        di = (st >> i) & 1
        dj = (st >> j) & 1
        diag += np.where(di == dj, 0.25 * J, -0.25 * J)
        key = (min(i, j), max(i, j))
        offs[key] = offs.get(key, 0.0) + 0.5 * J
    vals = [diag]
    for (i, j), amp in offs.items():
        di = (st >> i) & 1
        dj = (st >> j) & 1
        act = np.nonzero(di != dj)[0]
        tgt = st[act] ^ ((1 << i) | (1 << j))
        rows.append(act)
        cols.append(_index_of(sv, byval, tgt))
        vals.append(np.full(act.size, amp))
    ia, ja, val = _to_upper_csr(n, np.concatenate(rows), np.concatenate(cols), np.concatenate(vals))
    return n, ia, ja, val.astype(dtype)


def hubbard_upper_csr(nsites, nup, ndn, bonds, t=1.0, U=1.1, dtype=np.complex128):
    st = _enumerate(nsites, 2, (nup, ndn))
    n = st.size
    byval = np.argsort(st)
    sv = st[byval]
    upmask = np.int64(int("01" * nsites, 2))
    ndbl = _popcount(st & (st >> 1) & upmask)
    diag = np.zeros(n)
    for k in range(1, int(ndbl.max()) + 1):               # repeated addition like the LIL accumulation
        diag = np.where(ndbl >= k, diag + U, diag)
    amps = {}
    for (i, j) in bonds:
        key = (min(i, j), max(i, j))
        amps[key] = amps.get(key, 0.0) + (-t)
    rows, cols, vals = [np.arange(n)], [np.arange(n)], [diag]

    def parity_below(s_arr, site):
        return _popcount(s_arr & np.int64((1 << (2 * site)) - 1)) & 1

    for (i, j), amp in amps.items():
        for (to, frm) in ((i, j), (j, i)):
            for sp in (0, 1):
                bf = np.int64(1 << (2 * frm + sp))
                bt = np.int64(1 << (2 * to + sp))
                act = np.nonzero(((st & bf) != 0) & ((st & bt) == 0))[0]
                s0 = st[act]
                sg = parity_below(s0, frm)
                if sp == 1:
                    sg ^= ((s0 >> (2 * frm)) & 1)            # c_dn on a doubly occupied site: local element -1
                s1 = s0 ^ bf
                sg ^= parity_below(s1, to)
                if sp == 1:
                    sg ^= ((s1 >> (2 * to)) & 1)             # c+_dn next to an up electron: local element -1
                s2 = s1 ^ bt
                rows.append(act)
                cols.append(_index_of(sv, byval, s2))
                vals.append(np.where(sg == 1, -amp, amp))
    ia, ja, val = _to_upper_csr(n, np.concatenate(rows), np.concatenate(cols), np.concatenate(vals))
    return n, ia, ja, val.astype(dtype)


def square_szq_coefficients(Lx, Ly, qm, qn):
    """coeff_r of S^z_q = sum_r coeff_r (n_up,r - n_dn,r) as examples/trans_absent/latt_square/square_Fermi_Hubbard.cc:156-160
    builds it: 0.5/sqrt(N) exp(+i 2 pi (qm x/Lx + qn y/Ly)), site r = x + y*Lx."""
    c = np.zeros(Lx * Ly, dtype=np.complex128)
    for x in range(Lx):
        for y in range(Ly):
            c[x + y * Lx] = 0.5 / np.sqrt(Lx * Ly) * np.exp(2j * np.pi * (qm * x / Lx + qn * y / Ly))
    return c


def apply_onsite_diag(nsites, bps, counts, coef0, coef1, x, precision=2e-12):
    """model::moprXvec_full (src/model.cc:1468-1538) for one-site diagonal operators in Lin order: spins (bps = 1):
    sum_r coef0[r] S^z_r; electrons (bps = 2): sum_r coef0[r] n_up,r + coef1[r] n_dn,r.  Rows with |x_j| < precision are skipped."""
    st = basis_states(nsites, bps, counts)
    w = np.zeros(st.size, dtype=np.complex128)
    for s in range(nsites):
        if bps == 1:
            w += coef0[s] * np.where((st >> s) & 1, -0.5, 0.5)
        else:
            w += coef0[s] * ((st >> (2 * s)) & 1) + coef1[s] * ((st >> (2 * s + 1)) & 1)
    return np.where(np.abs(x) >= precision, x * w, 0.0)
