#!/usr/bin/env python
"""Offline LRU model of the vector traffic of the two-pass species-order product (CPU only; companion of simulate.py).

Same scaled problem as simulate.py -- Hubbard model on the 4x4 lattice, N_up = N_dn = 4 (3,312,400 states, 26.6 entries
per row), fully associative LRU of 32-byte sectors holding the same FRACTION of the vector as the B200's L2 holds of
BASELINE config 3's -- but the traversal of csrc/species.cu:
  pass 1  rows ascending in species order p = iu * D + id; gathers: own entry + the down hops (iu, id');  y written once;
  pass 2  rows by tiles of W down indices (tau, iu, id in tile); gathers: the up hops (iu', id);  y read and written.
Reported in units of the vector size: x misses of both passes, y misses of pass 2 (a y miss costs a read and a write),
and the total next to the 10.1 of the one-pass product in the reference's order (profiles/r01_cache_sim_orders_and_tiles.txt).

  gcc -O2 -o /tmp/lru scripts/cache_sim/lru.c && python scripts/cache_sim/species_two_pass.py /tmp/lru
"""
import itertools
import subprocess
import sys

import numpy as np

LRU = sys.argv[1] if len(sys.argv) > 1 else "/tmp/lru"
TRACE = "/tmp/qb_cache_sim_trace2.bin"
Lx = Ly = 4
ns, nup, ndn = 16, 4, 4
site = lambda x, y: (x % Lx) + (y % Ly) * Lx   # noqa: E731
bonds = sorted({(min(a, b), max(a, b)) for x in range(Lx) for y in range(Ly)
                for (a, b) in ((site(x, y), site(x + 1, y)), (site(x, y), site(x, y + 1)))})


def configs(k):
    return np.array(sorted(sum(1 << i for i in c) for c in itertools.combinations(range(ns), k)), dtype=np.int64)


def hop_lists(Cf):
    idx = {int(v): i for i, v in enumerate(Cf)}
    out = [[] for _ in range(Cf.size)]
    for (a, b) in bonds:
        for (i, j) in ((a, b), (b, a)):
            for s in np.nonzero(((Cf >> i) & 1 == 1) & ((Cf >> j) & 1 == 0))[0]:
                out[s].append(idx[int(Cf[s]) ^ ((1 << i) | (1 << j))])
    return [sorted(o) for o in out]


U, D = configs(nup), configs(ndn)
nU, nD = U.size, D.size
n = nU * nD
hu, hd = hop_lists(U), hop_lists(D)
print(f"dim {n}  up hops/config {np.mean([len(h) for h in hu]):.1f}  down hops/config {np.mean([len(h) for h in hd]):.1f}", flush=True)


def simulate(trace, cap, eps):
    with open(TRACE, "wb") as f:
        f.write(np.int64(trace.size).tobytes())
        f.write(trace.astype(np.int32).tobytes())
    return int(subprocess.run([LRU, TRACE, str(cap), str(eps)], capture_output=True, text=True).stdout.split()[1])


# pass 1 trace: for every row (iu, id) ascending: own entry, then the down-hop targets (sorted)
lens = np.array([1 + len(h) for h in hd])
per_block = np.concatenate([np.array(sorted([d] + hd[d])) for d in range(nD)])          # columns of one iu block, relative
trace1 = (np.arange(nU)[:, None] * nD + per_block[None, :]).ravel()

for eps, label in ((2, "complex"), (4, "fp64")):
    sectors = n // eps
    for frac in (0.024, 0.047, 0.10):
        cap = max(64, int(frac * sectors))
        m1 = simulate(trace1, cap, eps)
        for W in (8, 16, 32, 64):
            # pass 2 trace: tiles of W down indices; per (tau, iu): for id in tile: y access, then the up-hop targets (iu', id)
            t2 = []
            yspace = ((n + 63) // 64) * 64                                                  # y in an address space of its own
            for tau in range((nD + W - 1) // W):
                ids = np.arange(tau * W, min(nD, tau * W + W))
                for iu in range(nU):
                    cols = np.concatenate([(yspace + iu * nD + ids)[:, None], (np.array(hu[iu])[None, :] * nD + ids[:, None])], axis=1)
                    t2.append(cols.ravel())
            trace2 = np.concatenate(t2)
            is_y = trace2 >= yspace
            m2_all = simulate(trace2, cap, eps)
            m2_x = simulate(trace2[~is_y], cap, eps)
            m2_y = m2_all - m2_x
            total = (m1 + m2_x + 2 * m2_y) / sectors + 1.0                                   # + the y write of pass 1
            print(f"{label:7s} cache {frac * 100:4.1f}% of x  W={W:3d} (tile = {nU * W * 100.0 / (frac * n):5.1f}% of the cache): "
                  f"pass 1 x {m1 / sectors:.2f} | pass 2 x {m2_x / sectors:.2f}  y misses {m2_y / sectors:.2f} (rmw: x2) | "
                  f"total {total:.2f} vector sizes", flush=True)
