"""numpy restatement of the reference's translation-symmetric assembly for the single-orbital Fermi-Hubbard model.

TEST INFRASTRUCTURE ONLY -- groundwork for a device assembler of Hubbard momentum sectors ("next").
Same representative convention as tests/repr_builders.py (the Weisse tables do not know about statistics: they compare
bit patterns), with what fermions add:
  * basis digit per site: 0 empty, 1 up, 2 down, 3 both ("electron" orbital, src/basis.cc:49-96), two bits per site;
  * a translation permutes the singly occupied sites and picks up the parity of that permutation
    (mbasis_elem::transform, src/basis.cc:593-620: bubble-sort count over the images of the odd-fermion sites);
  * norm_trans_repr (src/basis.cc:2134-2147): a stabiliser translation t must satisfy k.t + sgn_t/2 integer;
  * generate_Ham_sparse_repr (src/model.cc:815-820): the element carries (-1)^sgn of the translation that maps the
    representative onto the state the Hamiltonian term produced;
  * hopping amplitudes with the sign rule of oprXphi (src/basis.cc:2717-2731), as in tests/lin_builders.py;
  * term order of mopr::operator+= (src/operators.cc:901-925): ascending (lower site, higher site); equal keys in
    REVERSE order of insertion (a new term is inserted before the ones it compares equal to).
Pinned bit for bit against matrices assembled by the compiled reference (qb_ref hubbard_k, tests/golden/repr_hashes.json)
and against the published sector energies of examples/trans_symmetric/latt_square/square_Fermi_Hubbard.cc:112-119.
"""
import numpy as np

import repr_builders as R


def square_hops(Lx, Ly):
    """Directed hops in the order the reference example inserts them: per site, +x then +y bond, and per bond
    c+_up,i c_up,j ; c+_up,j c_up,i ; c+_dn,i c_dn,j ; c+_dn,j c_dn,i  -> (to, from, spin)."""
    lat = R.Lattice([Lx, Ly])
    hops = []
    for x in range(Lx):
        for y in range(Ly):
            i = lat.site((x, y))
            for j in (lat.site((x + 1, y)), lat.site((x, y + 1))):
                hops += [(i, j, 0), (j, i, 0), (i, j, 1), (j, i, 1)]
    return hops


def _digits_apply(x, plan):
    out = np.zeros_like(x)
    for s, p in enumerate(plan):
        out |= ((x >> np.uint64(2 * s)) & np.uint64(3)) << np.uint64(2 * int(p))
    return out


class ElectronSector(R.Sector):
    def __init__(self, L, nup, ndn, k):
        par = self.par = R.Lattice(L)
        sub = self.sub = par.child()
        self.k = [int(x) for x in k]
        N, Ns = par.N, par.N // 2
        self.N, self.Ns, self.nup, self.ndn = N, Ns, nup, ndn
        if Ns > 8:
            raise ValueError("at most 16 sites in this restatement")
        u = np.uint64
        nhalf = 1 << (2 * Ns)
        allsub = np.arange(nhalf, dtype=u)
        splans = [sub.plan(d) for d in sub.disps]
        self.subT = np.stack([_digits_apply(allsub, p) for p in splans])
        rep = np.full(nhalf, -1, dtype=np.int64)
        dist = np.zeros(nhalf, dtype=np.int64)
        for x in range(nhalf):
            if rep[x] >= 0:
                continue
            for jd, y in enumerate(self.subT[:, x].astype(np.int64)):
                if rep[y] < 0:
                    rep[y] = x
                    dist[y] = jd
        self.rep, self.dist = rep, dist
        self.ntrans = len(par.disps)
        self.plans = [par.plan(d) for d in par.disps]
        sp_index = {tuple(p): j for j, p in enumerate(splans)}
        self.fwd = []
        for p in self.plans:
            if p[0] % 2 == 0:
                ja = sp_index[tuple(p[2 * t] // 2 for t in range(Ns))]
                jb = sp_index[tuple((p[2 * t + 1] - 1) // 2 for t in range(Ns))]
                self.fwd.append((False, ja, jb))
            else:
                ja = sp_index[tuple(p[2 * t + 1] // 2 for t in range(Ns))]
                jb = sp_index[tuple((p[2 * t] - 1) // 2 for t in range(Ns))]
                self.fwd.append((True, ja, jb))
        dindex = {d: i for i, d in enumerate(par.disps)}
        self.inv = [self.fwd[dindex[tuple((-x) % l for x, l in zip(d, par.L))]] for d in par.disps]
        self.kt = [sum(self.k[q] % par.L[q] * d[q] * (N // par.L[q]) for q in range(par.dim)) for d in par.disps]
        import cmath
        import math
        self.phase = []
        for d in par.disps:
            e = 0.0
            for q in range(par.dim):
                e += self.k[q] * d[q] / float(par.L[q])
            self.phase.append(cmath.exp(complex(0.0, 2.0 * math.pi * e)))
        self._enumerate()

    # halves <-> parent pattern, two bits per site
    def zip(self, a, b):
        out = np.zeros_like(a)
        for t in range(self.Ns):
            out |= ((a >> np.uint64(2 * t)) & np.uint64(3)) << np.uint64(4 * t)
            out |= ((b >> np.uint64(2 * t)) & np.uint64(3)) << np.uint64(4 * t + 2)
        return out

    def unzip(self, s):
        a = np.zeros_like(s)
        b = np.zeros_like(s)
        for t in range(self.Ns):
            a |= ((s >> np.uint64(4 * t)) & np.uint64(3)) << np.uint64(2 * t)
            b |= ((s >> np.uint64(4 * t + 2)) & np.uint64(3)) << np.uint64(2 * t)
        return a, b

    def translation_sign(self, st, i):
        """parity of the permutation translation i induces on the singly occupied sites of the parent patterns st"""
        p = self.plans[i]
        single = [(((st >> np.uint64(2 * s)) & np.uint64(1)) ^ ((st >> np.uint64(2 * s + 1)) & np.uint64(1))).astype(np.int64) for s in range(self.N)]
        sg = np.zeros(st.shape, dtype=np.int64)
        for s1 in range(self.N):
            for s2 in range(s1 + 1, self.N):
                if p[s1] > p[s2]:
                    sg += single[s1] & single[s2]
        return sg & 1

    def _enumerate(self):
        u = np.uint64
        Ns = self.Ns
        nhalf = 1 << (2 * Ns)
        allsub = np.arange(nhalf, dtype=u)
        reps = allsub[self.rep == np.arange(nhalf)]
        mask_up = u(int("01" * Ns, 2))
        cu = R._popcount(allsub & mask_up)
        cd = R._popcount((allsub >> u(1)) & mask_up)
        A, B = [], []
        for a in reps:
            ok = (cu == self.nup - cu[int(a)]) & (cd == self.ndn - cd[int(a)]) & (self.rep >= int(a))
            bsel = allsub[ok]
            A.append(np.full(bsel.size, a, dtype=u))
            B.append(bsel)
        a = np.concatenate(A)
        b = np.concatenate(B)
        _, ca, cb = self.canon(a, b)
        keep = (ca == a) & (cb == b)
        a, b = a[keep], b[keep]
        order = np.lexsort((a, b))
        a, b = a[order], b[order]
        self.lin_order = R.lin_tables_exist(a, b)
        st = self.zip(a, b)
        if not self.lin_order:
            order = np.argsort(st, kind="stable")
            a, b, st = a[order], b[order], st[order]
        self.a, self.b, self.states = a, b, st
        self.n = st.size
        self._byval = np.argsort(st, kind="stable")
        self._sorted = st[self._byval]
        cnt = np.zeros(self.n, dtype=np.int64)
        ok = np.ones(self.n, dtype=bool)
        for i in range(self.ntrans):
            ai, bi = self._move(self.fwd[i], a, b)
            fixed = (ai == a) & (bi == b)
            cnt += fixed
            bad = (self.kt[i] + self.translation_sign(st, i) * (self.N // 2)) % self.N != 0         # src/basis.cc:2129-2147
            ok &= ~(fixed & bad)
        self.nu = np.where(ok, (self.ntrans // cnt).astype(np.float64), 0.0)


def _fermions_below(st, site):
    return R._popcount(st & np.uint64((1 << (2 * site)) - 1)) & 1


def hubbard_sector_upper_csr(L, nup, ndn, k, hops, t=1.0, U=1.1, fake_pos=100.0):
    """Upper-triangle csr_mat<complex<double>> of the reference for the (N_up, N_dn, k) sector: (sector, ia, ja, val).
    hops: directed (to, from, spin) in the reference's insertion order (square_hops)."""
    S = ElectronSector(L, nup, ndn, k)
    n, st, nu = S.n, S.states, S.nu
    u = np.uint64
    live = nu > 0
    mask_up = u(int("01" * S.N, 2))
    ndbl = R._popcount(st & (st >> u(1)) & mask_up)
    diag = np.zeros(n)
    for r in range(1, int(ndbl.max()) + 1 if n else 1):
        diag = np.where(ndbl >= r, diag + U, diag)
    rows = np.arange(n)
    diag = np.where(live, diag, fake_pos + rows / float(n)).astype(np.complex128)
    # term order: ascending (lower site, higher site); equal keys in reverse insertion order
    order = sorted(range(len(hops)), key=lambda q: (min(hops[q][0], hops[q][1]), max(hops[q][0], hops[q][1]), -q))
    Rr, Cc, Vv, Oo = [], [], [], []
    for pos, q in enumerate(order):
        to, frm, sp = hops[q]
        bf, bt = u(1 << (2 * frm + sp)), u(1 << (2 * to + sp))
        act = np.nonzero(live & ((st & bf) != 0) & ((st & bt) == 0))[0]
        if act.size == 0:
            continue
        s0 = st[act]
        sg = _fermions_below(s0, frm)
        if sp == 1:
            sg ^= ((s0 >> u(2 * frm)) & u(1)).astype(np.int64)
        s1 = s0 ^ bf
        sg ^= _fermions_below(s1, to)
        if sp == 1:
            sg ^= ((s1 >> u(2 * to)) & u(1)).astype(np.int64)
        s2 = s1 ^ bt
        amp = np.where(sg == 1, t, -t)                              # (-t) * (-1)^sg
        a2, b2 = S.unzip(s2)
        i, ca, cb = S.canon(a2, b2)
        j = S.index(ca, cb)
        keep = (nu[j] > 0) & (j >= act)
        act, i, j, amp, ca, cb = act[keep], i[keep], j[keep], amp[keep], ca[keep], cb[keep]
        if act.size == 0:
            continue
        # sign of translating the representative by disp_i onto the produced state
        tsg = np.zeros(act.size, dtype=np.int64)
        cst = S.zip(ca, cb)
        for ti in np.unique(i):
            sel = i == ti
            tsg[sel] = S.translation_sign(cst[sel], int(ti))
        x = np.sqrt(nu[act] / nu[j]) * amp
        ph = np.asarray(S.phase)[i]
        val = x * ph.real + 1j * (x * ph.imag)
        val = np.where(tsg == 1, -val, val)
        Rr.append(act); Cc.append(j); Vv.append(val); Oo.append(np.full(act.size, pos))
    Rr, Cc, Vv, Oo = (np.concatenate(z) for z in (Rr, Cc, Vv, Oo))
    o = np.lexsort((Oo, Cc, Rr))
    Rr, Cc, Vv = Rr[o], Cc[o], Vv[o]
    first = np.ones(Rr.size, dtype=bool)
    first[1:] = (Rr[1:] != Rr[:-1]) | (Cc[1:] != Cc[:-1])
    gid = np.cumsum(first) - 1
    ng = int(gid[-1]) + 1
    gr, gc = Rr[first], Cc[first]
    rank = np.arange(Rr.size) - np.nonzero(first)[0][gid]
    isdiag = gr == gc
    acc = np.where(isdiag, diag[gr], 0.0 + 0.0j)
    present = isdiag.copy()
    for tt in range(int(rank.max()) + 1):                           # LIL accumulation, src/sparse.cc:57-81
        sel = rank == tt
        g = gid[sel]
        acc[g] = acc[g] + Vv[sel]
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
    o = np.lexsort((gc, gr))
    gr, gc, gv = gr[o], gc[o], gv[o]
    ia = np.zeros(n + 1, dtype=np.int64)
    np.add.at(ia, gr + 1, 1)
    return S, np.cumsum(ia), gc.astype(np.int64), gv
