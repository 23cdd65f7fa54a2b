"""numpy restatement of the reference's translation-symmetric ("repr") assembly for the spin-1/2 Heisenberg model.

TEST INFRASTRUCTURE ONLY (an oracle-side builder, like lin_builders.py).  It restates, for untilted lattices with one
site per unit cell and periodic boundaries in every direction (chain, square, triangular ...):

  - the site numbering of lattice::site2coor_old (src/lattice.cc:591-615): the first even direction ("dim_spec") is the
    fastest digit, so parent sites 2t / 2t+1 are the two sublattices that zipper_basis interleaves (src/basis.cc:946-969);
  - classify_trans_full2rep (src/basis.cc:1351-1421): sublattice representatives = first state of each orbit in integer
    order, dist2rep = first displacement (lexicographic, last index fastest) that reaches a state;
  - the representative convention of the Weisse tables (classify_Weisse_tables, src/basis.cc:1670-2101, used at
    src/model.cc:356-405 and :760-804).  Stated without the tables: in the orbit of a parent state pick the elements
    whose even-site half is the smaller of the two sublattice representatives itself; among those take the one whose
    odd-site half has the smallest dist2rep, and among equal ones the smallest parent displacement;
  - the basis order: Lin order (odd-site label, even-site label) when the Lin tables J = Ja[ia] + Jb[ib] exist, else
    plain integer order (src/model.cc:435-443, src/basis.cc:1193-1348, ALGraph::BSF_set_JaJb src/miscellaneous.cc:660-708);
  - norm_trans_repr (src/basis.cc:2104-2202): nu = orbit size if every stabiliser translation t has k.t integer, else 0;
  - model::generate_Ham_sparse_repr (src/model.cc:688-836): rows of zero norm carry only fake_pos + i/dim on the
    diagonal; otherwise coef = sqrt(nu_i/nu_j) * conj(c) * exp(2 pi i k.disp_i / L), accumulated with the LIL rules
    (src/sparse.cc:57-81: off-diagonal partial sums below 1e-14 are erased) in the order of mopr::operator+=
    (src/operators.cc:901-925): terms sorted by (lower site, higher site).

Pinned bit-for-bit (dim, ia, ja and every value) against matrices assembled by the compiled reference: the goldens
heis16_k3 and tri4x4_k00/k01/k12, and - when oracle/_ref/qb_ref is present - chains L=12 (all k), L=14, L=16 and
triangular 4x2, 2x4, 4x3, 3x4, 6x2 clusters (tests/test_builders_cpu.py).
"""
import cmath
import itertools
import math

import numpy as np


class Lattice:
    """Untilted lattice, one site per cell, PBC everywhere; numbering of the reference (dim_spec digit fastest)."""

    def __init__(self, L, spec=None):
        self.L = [int(x) for x in L]
        self.dim = len(self.L)
        if spec is None:
            even = [d for d in range(self.dim) if self.L[d] % 2 == 0]
            if not even:
                raise ValueError("no even direction: the reference cannot divide this lattice")
            spec = even[0]
        self.spec = spec
        self.order = [spec] + [d for d in range(self.dim) if d != spec]
        self.N = int(np.prod(self.L))
        self.coor = np.zeros((self.N, self.dim), dtype=np.int64)
        for s in range(self.N):
            r = s
            for d in self.order:
                self.coor[s, d] = r % self.L[d]
                r //= self.L[d]
        self.disps = list(itertools.product(*[range(x) for x in self.L]))       # index 0 most significant

    def site(self, c):
        s, mul = 0, 1
        for d in self.order:
            s += (int(c[d]) % self.L[d]) * mul
            mul *= self.L[d]
        return s

    def plan(self, disp):
        """plan[site] = site + disp (lattice::translation_plan); a transform writes old[site] to new[plan[site]]."""
        return [self.site(self.coor[s] + np.asarray(disp)) for s in range(self.N)]

    def child(self):
        Ls = list(self.L)
        Ls[self.spec] //= 2
        return Lattice(Ls, self.spec)


def chain_bonds(L):
    return [(x, (x + 1) % L) for x in range(L)]


def triangular_bonds(Lx, Ly):
    """Bonds +x, +x+y, +y of every site, numbered like the reference's lattice (oracle/ref_driver.cc build_triangular)."""
    lat = Lattice([Lx, Ly])
    b = []
    for x in range(Lx):
        for y in range(Ly):
            s = lat.site((x, y))
            b += [(s, lat.site((x + 1, y))), (s, lat.site((x + 1, y + 1))), (s, lat.site((x, y + 1)))]
    return b


def square_bonds(Lx, Ly):
    lat = Lattice([Lx, Ly])
    b = []
    for x in range(Lx):
        for y in range(Ly):
            s = lat.site((x, y))
            b += [(s, lat.site((x + 1, y))), (s, lat.site((x, y + 1)))]
    return b


def _apply_plan(x, plan):
    out = np.zeros_like(x)
    for s, p in enumerate(plan):
        out |= ((x >> np.uint64(s)) & np.uint64(1)) << np.uint64(p)
    return out


def _popcount(a):
    a = a.copy()
    c = np.zeros(a.shape, dtype=np.int64)
    while np.any(a):
        c += (a & np.uint64(1)).astype(np.int64)
        a >>= np.uint64(1)
    return c


def lin_tables_exist(ia, ib):
    """True when J = Ja[ia] + Jb[ib] has a solution for J = 0..n-1 (weighted union-find over the two label sets)."""
    parent, off, ids = [], [], {}

    def node(key):
        if key not in ids:
            ids[key] = len(parent)
            parent.append(len(parent))
            off.append(0)
        return ids[key]

    def find(x):
        path = []
        while parent[x] != x:
            path.append(x)
            x = parent[x]
        acc = 0
        for y in reversed(path):
            acc += off[y]
            off[y] = acc
            parent[y] = x
        return x

    for J, (a, b) in enumerate(zip(ia.tolist(), ib.tolist())):
        x, y = node((0, a)), node((1, b))
        rx, ry = find(x), find(y)
        px = off[x] if x != rx else 0
        py = off[y] if y != ry else 0
        if rx == ry:
            if px - py != J:
                return False
        else:
            parent[rx] = ry
            off[rx] = J + py - px
    return True


class Sector:
    """Representatives, norms and lookup of one (Sz, k) sector."""

    def __init__(self, L, ndown, k):
        par = self.par = Lattice(L)
        sub = self.sub = par.child()
        self.k = [int(x) for x in k]
        N, Ns = par.N, par.N // 2
        self.N, self.Ns, self.ndown = N, Ns, ndown
        if Ns > 16:
            raise ValueError("at most 32 sites")
        u = np.uint64
        allsub = np.arange(1 << Ns, dtype=u)
        splans = [sub.plan(d) for d in sub.disps]
        self.subT = np.stack([_apply_plan(allsub, p) for p in splans])                     # subT[j][x]
        rep = np.full(1 << Ns, -1, dtype=np.int64)
        dist = np.zeros(1 << Ns, dtype=np.int64)
        for x in range(1 << Ns):                                                             # src/basis.cc:1385-1419
            if rep[x] >= 0:
                continue
            orbit = self.subT[:, x].astype(np.int64)
            for jd, y in enumerate(orbit):
                if rep[y] < 0:
                    rep[y] = x
                    dist[y] = jd
        self.rep, self.dist = rep, dist
        # every parent translation acts on the two halves by sublattice translations, possibly exchanging them
        self.ntrans = len(par.disps)
        sp_index = {tuple(p): j for j, p in enumerate(splans)}
        self.fwd = []
        for d in par.disps:
            p = par.plan(d)
            if p[0] % 2 == 0:
                ja = sp_index[tuple(p[2 * t] // 2 for t in range(Ns))]
                jb = sp_index[tuple((p[2 * t + 1] - 1) // 2 for t in range(Ns))]
                self.fwd.append((False, ja, jb))
            else:                                             # new even half = S_ja(old odd half), new odd half = S_jb(old even half)
                ja = sp_index[tuple(p[2 * t + 1] // 2 for t in range(Ns))]
                jb = sp_index[tuple((p[2 * t] - 1) // 2 for t in range(Ns))]
                self.fwd.append((True, ja, jb))
        dindex = {d: i for i, d in enumerate(par.disps)}
        self.inv = [self.fwd[dindex[tuple((-x) % l for x, l in zip(d, par.L))]] for d in par.disps]
        self.compat = [sum(self.k[q] % par.L[q] * d[q] * (N // par.L[q]) for q in range(par.dim)) % N for d in par.disps]
        self.phase = []
        for d in par.disps:                                                                  # src/model.cc:808-814
            e = 0.0
            for q in range(par.dim):
                e += self.k[q] * d[q] / float(par.L[q])
            self.phase.append(cmath.exp(complex(0.0, 2.0 * math.pi * e)))
        self._enumerate()

    # (a, b) halves <-> bit pattern with site s at bit s
    def zip(self, a, b):
        out = np.zeros_like(a)
        for t in range(self.Ns):
            out |= ((a >> np.uint64(t)) & np.uint64(1)) << np.uint64(2 * t)
            out |= ((b >> np.uint64(t)) & np.uint64(1)) << np.uint64(2 * t + 1)
        return out

    def unzip(self, s):
        a = np.zeros_like(s)
        b = np.zeros_like(s)
        for t in range(self.Ns):
            a |= ((s >> np.uint64(2 * t)) & np.uint64(1)) << np.uint64(t)
            b |= ((s >> np.uint64(2 * t + 1)) & np.uint64(1)) << np.uint64(t)
        return a, b

    def _move(self, tr, a, b):
        swap, ja, jb = tr
        if swap:
            return self.subT[ja][b], self.subT[jb][a]
        return self.subT[ja][a], self.subT[jb][b]

    def canon(self, a, b):
        """For halves (a, b): (disp index i, a_c, b_c) with state = T_i (a_c zip b_c) and (a_c, b_c) the representative."""
        rs = np.minimum(self.rep[a], self.rep[b]).astype(np.uint64)
        big = np.iinfo(np.int64).max
        best = np.full(a.shape, big, dtype=np.int64)
        ca = np.zeros_like(a)
        cb = np.zeros_like(b)
        for i in range(self.ntrans):
            ai, bi = self._move(self.inv[i], a, b)
            key = self.dist[bi] * self.ntrans + i
            upd = (ai == rs) & (key < best)
            best = np.where(upd, key, best)
            ca = np.where(upd, ai, ca)
            cb = np.where(upd, bi, cb)
        assert np.all(best < big)
        return best % self.ntrans, ca, cb

    def _enumerate(self):
        u = np.uint64
        Ns = self.Ns
        allsub = np.arange(1 << Ns, dtype=u)
        reps = allsub[self.rep == np.arange(1 << Ns)]
        pc = _popcount(allsub)
        A, B = [], []
        for a in reps:                                          # candidates: even half a representative not above rep[odd half]
            ok = (pc == self.ndown - pc[int(a)]) & (self.rep >= int(a))
            bsel = allsub[ok]
            A.append(np.full(bsel.size, a, dtype=u))
            B.append(bsel)
        a = np.concatenate(A)
        b = np.concatenate(B)
        _, ca, cb = self.canon(a, b)
        keep = (ca == a) & (cb == b)
        a, b = a[keep], b[keep]
        order = np.lexsort((a, b))                              # Lin order: odd-site label first
        a, b = a[order], b[order]
        self.lin_order = lin_tables_exist(a, b)
        st = self.zip(a, b)
        if not self.lin_order:
            order = np.argsort(st, kind="stable")
            a, b, st = a[order], b[order], st[order]
        self.a, self.b, self.states = a, b, st
        self.n = st.size
        self._byval = np.argsort(st, kind="stable")
        self._sorted = st[self._byval]
        # norms
        cnt = np.zeros(self.n, dtype=np.int64)
        ok = np.ones(self.n, dtype=bool)
        for i in range(self.ntrans):
            ai, bi = self._move(self.fwd[i], a, b)
            fixed = (ai == a) & (bi == b)
            cnt += fixed
            if self.compat[i] != 0:
                ok &= ~fixed
        self.nu = np.where(ok, (self.ntrans // cnt).astype(np.float64), 0.0)

    def index(self, ca, cb):
        st = self.zip(ca, cb)
        pos = np.searchsorted(self._sorted, st)
        pos = np.minimum(pos, self.n - 1)
        assert np.all(self._sorted[pos] == st)
        return self._byval[pos]


def heisenberg_sector_upper_csr(L, ndown, k, bonds, J=1.0, fake_pos=100.0, sector=None):
    """Upper-triangle csr_mat<complex<double>> of the reference for the (Sz, k) sector: (sector, ia, ja, val)."""
    S = sector if sector is not None else Sector(L, ndown, k)
    n, st, nu = S.n, S.states, S.nu
    u = np.uint64
    bonds = sorted((min(p, q), max(p, q)) for (p, q) in bonds)
    live = nu > 0
    diag = np.zeros(n)
    for (p, q) in bonds:
        par = ((st >> u(p)) & u(1)) == ((st >> u(q)) & u(1))
        diag = diag + np.where(par, 0.25 * J, -0.25 * J)
    rows = np.arange(n)
    diag = np.where(live, diag, fake_pos + rows / float(n)).astype(np.complex128)
    R, C, V, O = [], [], [], []
    for t, (p, q) in enumerate(bonds):
        act = np.nonzero(live & (((st >> u(p)) & u(1)) != ((st >> u(q)) & u(1))))[0]
        if act.size == 0:
            continue
        s2 = st[act] ^ u((1 << p) | (1 << q))
        a2, b2 = S.unzip(s2)
        i, ca, cb = S.canon(a2, b2)
        j = S.index(ca, cb)
        keep = (nu[j] > 0) & (j >= act)
        act, i, j = act[keep], i[keep], j[keep]
        x = np.sqrt(nu[act] / nu[j]) * (0.5 * J)
        ph = np.asarray(S.phase)[i]
        R.append(act)
        C.append(j)
        V.append(x * ph.real + 1j * (x * ph.imag))
        O.append(np.full(act.size, t))
    R, C, V, O = (np.concatenate(z) for z in (R, C, V, O))
    order = np.lexsort((O, C, R))
    R, C, V = R[order], C[order], V[order]
    first = np.ones(R.size, dtype=bool)
    first[1:] = (R[1:] != R[:-1]) | (C[1:] != C[:-1])
    gid = np.cumsum(first) - 1
    ng = int(gid[-1]) + 1
    gr, gc = R[first], C[first]
    rank = np.arange(R.size) - np.nonzero(first)[0][gid]
    isdiag = gr == gc
    acc = np.where(isdiag, diag[gr], 0.0 + 0.0j)
    present = isdiag.copy()
    for t in range(int(rank.max()) + 1):                        # LIL accumulation, src/sparse.cc:57-81
        sel = rank == t
        g = gid[sel]
        acc[g] = acc[g] + V[sel]
        present[g] = True
        dead = np.zeros(ng, dtype=bool)
        dead[g] = (~isdiag[g]) & (np.abs(acc[g]) < 1e-14)
        acc[dead] = 0.0
        present[dead] = False
    have_diag = np.zeros(n, dtype=bool)
    have_diag[gr[isdiag]] = True
    missing = np.nonzero(~have_diag)[0]
    gr = np.concatenate([gr[present], missing])
    gc = np.concatenate([gc[present], missing])
    gv = np.concatenate([acc[present], diag[missing]])
    order = np.lexsort((gc, gr))
    gr, gc, gv = gr[order], gc[order], gv[order]
    ia = np.zeros(n + 1, dtype=np.int64)
    np.add.at(ia, gr + 1, 1)
    return S, np.cumsum(ia), gc.astype(np.int64), gv


def apply_sz(S_old, S_new, coef, x):
    """model::moprXvec_repr (src/model.cc:1716-1846) restricted to diagonal one-site terms A = sum_r coef[r] S^z_r:
    y[j] = sum_r sqrt(nu_old[j]/nu_new[j]) * x[j] * coef[r] * sz_r(state_j); rows with |x[j]|, nu_old[j] or nu_new[j]
    below lanczos_precision contribute nothing (:1752, :1760)."""
    assert S_old.n == S_new.n and np.array_equal(S_old.states, S_new.states)
    u = np.uint64
    y = np.zeros(S_old.n, dtype=np.complex128)
    ok = (np.abs(x) >= 2e-12) & (S_old.nu >= 2e-12) & (S_new.nu > 2e-12)
    t = np.zeros(S_old.n, dtype=np.complex128)
    t[ok] = np.sqrt(S_old.nu[ok] / S_new.nu[ok]) * x[ok]
    for r in range(S_old.N):
        sz = np.where((S_old.states >> u(r)) & u(1), -0.5, 0.5)
        y += t * (coef[r] * sz)
    return y


def szq_coefficients(L, q):
    """exp(-i 2 pi q x / L) / sqrt(L) per site of a chain (oracle/ref_driver.cc flow_heis_chain_szq)."""
    Q = 2.0 * 3.1415926535897932 * q / float(L)
    return np.array([cmath.exp(complex(0.0, -Q * x)) / math.sqrt(float(L)) for x in range(L)])


def apply_splus(S_old, S_new, coef, x):
    """The same branch for A = sum_r coef[r] S^+_r (acts on down spins; the new sector has one fewer).  No reference
    vector pins it directly; tests pin it to apply_sminus through <w, A v> = <A^+ w, v>."""
    return _apply_ladder(S_old, S_new, coef, x, 1)


def _apply_ladder(S_old, S_new, coef, x, digit):
    u = np.uint64
    y = np.zeros(S_new.n, dtype=np.complex128)
    ok = (np.abs(x) >= 2e-12) & (S_old.nu >= 2e-12)
    rows = np.nonzero(ok)[0]
    phase = np.conj(np.asarray(S_new.phase))
    for r in range(S_old.N):
        act = rows[((S_old.states[rows] >> u(r)) & u(1)) == digit]
        if act.size == 0:
            continue
        s2 = S_old.states[act] ^ (u(1) << u(r))
        a2, b2 = S_new.unzip(s2)
        i, ca, cb = S_new.canon(a2, b2)
        tgt = S_new.index(ca, cb)
        keep = S_new.nu[tgt] >= 2e-12
        act, i, tgt = act[keep], i[keep], tgt[keep]
        val = np.sqrt(S_old.nu[act] / S_new.nu[tgt]) * x[act] * coef[r] * phase[i]
        np.add.at(y, tgt, val)
    return y


def apply_sminus(S_old, S_new, coef, x):
    """model::moprXvec_repr (src/model.cc:1762-1834), off-diagonal branch, for A = sum_r coef[r] S^-_r on spin-1/2:
    every term lowers one up spin of the old representative; the produced state is brought to its representative in the
    NEW sector (one more down spin) and
        y[i] += sqrt(nu_old[j] / nu_new[i]) * x[j] * coef[r] * exp(-2 pi i k_new . disp_i / L).
    (The reference adds these from several threads; which order the contributions to one y[i] arrive in is not fixed,
    so agreement is to rounding, not to the bit.)"""
    u = np.uint64
    y = np.zeros(S_new.n, dtype=np.complex128)
    ok = (np.abs(x) >= 2e-12) & (S_old.nu >= 2e-12)
    rows = np.nonzero(ok)[0]
    phase = np.conj(np.asarray(S_new.phase))                     # exp(-2 pi i k_new . disp / L)
    for r in range(S_old.N):
        act = rows[((S_old.states[rows] >> u(r)) & u(1)) == 0]   # digit 0 = up: S^- acts
        if act.size == 0:
            continue
        s2 = S_old.states[act] | (u(1) << u(r))
        a2, b2 = S_new.unzip(s2)
        i, ca, cb = S_new.canon(a2, b2)
        tgt = S_new.index(ca, cb)
        keep = S_new.nu[tgt] >= 2e-12
        act, i, tgt = act[keep], i[keep], tgt[keep]
        val = np.sqrt(S_old.nu[act] / S_new.nu[tgt]) * x[act] * coef[r] * phase[i]
        np.add.at(y, tgt, val)
    return y
