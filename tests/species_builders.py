"""numpy restatement of the *species order* of the single-orbital Hubbard basis and of the two-part split of H.

TEST INFRASTRUCTURE ONLY.  The reference orders the basis by its Lin tables (label of the odd sites, label of the even
sites; src/basis.cc:1144-1190), in which every hopping term moves a state far away in the index space.  libqbgpu's
QBGPU_SPECIES_ORDER handles keep the vectors *internally* in the order

    p = rank(up configuration) * D_dn + rank(down configuration)

(configurations = occupancy words of one spin species, ranked by (odd-site bits, even-site bits), see _order_key), in which
    H = [ U * (double occupancies)  +  hops of the down electrons ]      "local" part : stays inside the block of one iu
      + [ hops of the up electrons ]                                      "cross" part : same id, another iu
and every entry factorises as (amplitude and sign from the hopping species' own configuration) x (parity of the OTHER
species' electrons on a site interval).  This file restates that factorisation -- the hop tables, the interval masks,
the permutation between the two orders -- and tests/test_species_cpu.py checks it entry for entry against
tests/lin_builders.py, which is pinned to matrices assembled by the compiled reference.  The CUDA code
(quantum_basis_b200/csrc/species.cu) builds the same tables on the host with the same formulas.

Sign rule being factorised (reference oprXphi, src/basis.cc:2717-2731; fermions ordered site-major, up before down on
a site): for c+_{t,s} c_{f,s} the parity is  [#fermions below f] + [s = dn and up-occupied f] + [#fermions below t in
the intermediate state] + [s = dn and up-occupied t].
"""
import numpy as np

import lin_builders as lb


def _order_key(words, nsites):
    """Sort key of an occupancy word: (bits of the odd sites, bits of the even sites), odd sites major.  With this key the
    species order refines the reference's Lin order block-wise: the rows of one odd-site label (a contiguous block of the
    reference's order) occupy a RECTANGLE [iu0, iu0 + Cu) x [id0, id0 + Cd) of the species order, so the permutation between
    the two orders moves whole blocks -- cache-friendly in both directions (csrc/species.cu: build_host_tables)."""
    odd = np.zeros_like(words)
    even = np.zeros_like(words)
    for s_ in range(nsites):
        bit = (words >> s_) & 1
        if s_ % 2:
            odd |= bit << (s_ // 2)
        else:
            even |= bit << (s_ // 2)
    return (odd << nsites) | even


def configurations(nsites, nel):
    """Occupancy words of one species with `nel` electrons, in the species order (see _order_key)."""
    allw = np.arange(1 << nsites, dtype=np.int64)
    w = allw[lb._popcount(allw) == nel]
    return w[np.argsort(_order_key(w, nsites), kind="stable")]


def rank_in_class(nsites):
    """rank[w] = position of w among the words with the same popcount, in the species order."""
    allw = np.arange(1 << nsites, dtype=np.int64)
    pc = lb._popcount(allw)
    key = _order_key(allw, nsites)
    rank = np.zeros(allw.size, dtype=np.int64)
    for c in range(nsites + 1):
        sel = np.nonzero(pc == c)[0]
        rank[sel[np.argsort(key[sel], kind="stable")]] = np.arange(sel.size)
    return rank


def _compress_species(states, nsites, sp):
    """Lin-builder states carry 2 bits per site (bit 2s = up, 2s+1 = dn): extract the occupancy word of species sp."""
    w = np.zeros_like(states)
    for s in range(nsites):
        w |= ((states >> (2 * s + sp)) & 1) << s
    return w


def species_perm(nsites, nup, ndn):
    """perm[r] = species-order index of the reference's (Lin-order) basis state r."""
    st = lb.basis_states(nsites, 2, (nup, ndn))
    rank = rank_in_class(nsites)
    up = _compress_species(st, nsites, 0)
    dn = _compress_species(st, nsites, 1)
    Dd = configurations(nsites, ndn).size
    return rank[up] * Dd + rank[dn]


def merged_bonds(bonds):
    """Undirected bonds with multiplicity, like merge_bonds() in builders.cu."""
    w = {}
    for (i, j) in bonds:
        key = (min(i, j), max(i, j))
        w[key] = w.get(key, 0) + 1
    return w


def hop_table(nsites, nel, bonds, t, species):
    """CSR-like hop table of one species.

    Returns (ptr, target, amp, mask): for configuration index c the hops ptr[c]:ptr[c+1], sorted by target; amp carries
    the species' own sign; mask selects the sites of the OTHER species whose electrons flip the sign:
      up hop  f -> t : sites in [min(f,t), max(f,t))      (below_f ^ below_t)
      down hop f -> t: sites in (min(f,t), max(f,t)]      (the up electron on the same site counts as "below" a down one)
    """
    conf = configurations(nsites, nel)
    rank = rank_in_class(nsites)
    wb = merged_bonds(bonds)
    rows = []
    for c, word in enumerate(conf.tolist()):
        ent = []
        for (i, j), w in wb.items():
            for (f, tt) in ((i, j), (j, i)):
                if not ((word >> f) & 1) or ((word >> tt) & 1):
                    continue
                below_f = (1 << f) - 1
                below_t = (1 << tt) - 1
                own = (bin(word & below_f).count("1") + bin(word & below_t).count("1") + (1 if f < tt else 0)) & 1
                amp = 0.0
                for _ in range(w):
                    amp += -t
                if species == 0:
                    mask = below_f ^ below_t
                else:
                    mask = ((2 << f) - 1) ^ ((2 << tt) - 1)
                new = word ^ (1 << f) ^ (1 << tt)
                ent.append((int(rank[new]), -amp if own else amp, mask))
        ent.sort(key=lambda e: e[0])
        rows.append(ent)
    ptr = np.zeros(conf.size + 1, dtype=np.int64)
    ptr[1:] = np.cumsum([len(r) for r in rows])
    target = np.array([e[0] for r in rows for e in r], dtype=np.int64)
    amp = np.array([e[1] for r in rows for e in r], dtype=np.float64)
    mask = np.array([e[2] for r in rows for e in r], dtype=np.int64)
    return ptr, target, amp, mask


def species_parts(nsites, nup, ndn, bonds, t=1.0, U=1.1):
    """(local, cross) parts of H in species order as scipy CSR matrices (full rows, both triangles)."""
    import scipy.sparse as sp
    ul = configurations(nsites, nup)
    dl = configurations(nsites, ndn)
    Du, Dd = ul.size, dl.size
    n = Du * Dd
    pu, tu, au, mu = hop_table(nsites, nup, bonds, t, 0)
    pd, td, ad, md = hop_table(nsites, ndn, bonds, t, 1)
    iu = np.repeat(np.arange(Du), Dd)
    idn = np.tile(np.arange(Dd), Du)
    # diagonal: U added once per doubly occupied site (repeated addition, like the LIL accumulation)
    ndbl = lb._popcount(ul[iu] & dl[idn])
    diag = np.zeros(n)
    for k in range(1, int(ndbl.max()) + 1 if n else 1):
        diag = np.where(ndbl >= k, diag + U, diag)
    rows, cols, vals = [np.arange(n)], [np.arange(n)], [diag]
    # local part: down hops, inside the block of one iu
    for d in range(Dd):
        for e in range(pd[d], pd[d + 1]):
            r = np.arange(Du) * Dd + d
            par = lb._popcount(ul & md[e]) & 1
            rows.append(r)
            cols.append(np.arange(Du) * Dd + td[e])
            vals.append(np.where(par == 1, -ad[e], ad[e]))
    local = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    rows, cols, vals = [], [], []
    for u in range(Du):
        for e in range(pu[u], pu[u + 1]):
            r = u * Dd + np.arange(Dd)
            par = lb._popcount(dl & mu[e]) & 1
            rows.append(r)
            cols.append(tu[e] * Dd + np.arange(Dd))
            vals.append(np.where(par == 1, -au[e], au[e]))
    if rows:
        cross = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    else:
        cross = sp.csr_matrix((n, n))
    return local, cross


def slice_order(Du, Dd, tile):
    """Traversal order of the 32-row slices of the cross part: by (tile of the slice's first down index, slice)."""
    n = Du * Dd
    ns = (n + 31) // 32
    first = np.arange(ns, dtype=np.int64) * 32
    t_of = (first % Dd) // tile
    return np.lexsort((np.arange(ns), t_of))
