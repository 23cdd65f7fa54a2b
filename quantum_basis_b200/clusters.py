"""Finite clusters of the triangular / square lattice spanned by two superlattice vectors A0, A1 (possibly tilted).

Host-side model definition (sites, translation permutations, characters, bonds) for the orbit assembler (qbgpu_build_heisenberg_orbit, BASELINE config 4): sites are the points of
Z^2 / (Z A0 + Z A1); the translation group is that same quotient acting by addition; its characters are
chi_m(R) = exp(2 pi i m . (R M^-1)) with M = [A0; A1] and m an integer pair.  The reference's 31-site cluster
(latt_special/triangular_31site.toml:11-12) is A0 = [5, 1], A1 = [-1, 6].
"""
import cmath
from fractions import Fraction

import numpy as np


class Cluster:
    def __init__(self, A0, A1):
        self.A0, self.A1 = tuple(A0), tuple(A1)
        self.det = A0[0] * A1[1] - A0[1] * A1[0]
        assert self.det > 0
        pts = {}
        span = abs(A0[0]) + abs(A0[1]) + abs(A1[0]) + abs(A1[1]) + 1
        for x in range(-span, span + 1):
            for y in range(-span, span + 1):
                key = self.canon((x, y))
                pts.setdefault(key, None)
        self.sites = sorted(pts)                       # canonical fractional coordinates (numerators over det)
        assert len(self.sites) == self.det
        self.index = {k: i for i, k in enumerate(self.sites)}
        # one integer point per site
        self.point = [None] * self.det
        for x in range(-span, span + 1):
            for y in range(-span, span + 1):
                i = self.index[self.canon((x, y))]
                if self.point[i] is None or (abs(x) + abs(y), x, y) < (abs(self.point[i][0]) + abs(self.point[i][1]),) + self.point[i]:
                    self.point[i] = (x, y)

    def canon(self, R):
        """fractional coordinates of R in the (A0, A1) basis reduced to [0, 1)^2, as numerators over det"""
        x, y = R
        f0 = (x * self.A1[1] - y * self.A1[0]) % self.det          # R = f0/det A0 + f1/det A1
        f1 = (-x * self.A0[1] + y * self.A0[0]) % self.det
        return (f0, f1)

    def site_of(self, R):
        return self.index[self.canon(R)]

    def translations(self):
        """perms[t][s] = site reached from s by the translation that moves site 0's point to site t's point"""
        o = self.point[0]
        perms = []
        for t in range(self.det):
            d = (self.point[t][0] - o[0], self.point[t][1] - o[1])
            perms.append([self.site_of((self.point[s][0] + d[0], self.point[s][1] + d[1])) for s in range(self.det)])
        return np.array(perms, dtype=np.int32)

    def characters(self, m):
        o = self.point[0]
        chi = []
        for t in range(self.det):
            d = (self.point[t][0] - o[0], self.point[t][1] - o[1])
            f0 = Fraction(d[0] * self.A1[1] - d[1] * self.A1[0], self.det)
            f1 = Fraction(-d[0] * self.A0[1] + d[1] * self.A0[0], self.det)
            ph = (m[0] * f0 + m[1] * f1) % 1
            chi.append(cmath.exp(2j * cmath.pi * float(ph)))
        return np.array(chi)

    def distinct_momenta(self):
        """one integer pair per irrep of the translation group"""
        seen, out = set(), []
        for m0 in range(self.det):
            for m1 in range(self.det):
                key = tuple(np.round(self.characters((m0, m1)), 9))
                if key not in seen:
                    seen.add(key)
                    out.append((m0, m1))
                if len(out) == self.det:
                    return out
        return out

    def triangular_bonds(self):
        """+a1, +a1+a2, +a2 from every site (the bond set of oracle/ref_driver.cc build_triangular)"""
        b = []
        for s in range(self.det):
            x, y = self.point[s]
            for d in ((1, 0), (1, 1), (0, 1)):
                b.append((s, self.site_of((x + d[0], y + d[1]))))
        return b
