"""Offline check for DESIGN section 8: shared-memory bank conflicts of the per-code lookups in the term-coded replay.

Replays the code streams of sampled 32-row slices (Hubbard 4x3, N_up = N_dn = 6, walk order) and reports the mean
conflict degree of a 256-entry table lookup (distinct addresses per bank; equal addresses broadcast).  Result: 1.36 for
32-bit tables -- not the bottleneck.  python scripts/cache_sim/term_code_bank_conflicts.py
"""
# Estimate shared-memory bank conflicts of the per-code table lookups in the term-coded replay (Hubbard 4x3, N_up=N_dn=6).
import sys, numpy as np
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'tests'))
import lin_builders as B
ns = 12; bonds = B.square_bonds(4, 3)
st = B.basis_states(ns, 2, (6, 6)).astype(np.uint64)          # Lin order, 2 bits per site (up = bit 0, dn = bit 1)
n = st.size
nbr = [set() for _ in range(ns)]
for (i, j) in bonds:
    nbr[i].add(j); nbr[j].add(i)
code_of = {}
for f in range(ns):
    for t in sorted(nbr[f]):
        code_of[(f, t)] = len(code_of)
# walk order codes per row
rows = []
for r in range(0, n, max(1, n // 20000)):                      # sample slices of 32 consecutive rows
    pass
rng = np.random.default_rng(1)
starts = rng.integers(0, n // 32 - 1, size=1500) * 32
deg32 = []; deg64 = []
for s0 in starts:
    lists = []
    for r in range(s0, s0 + 32):
        s = int(st[r]); codes = []
        for sp in (0, 1):
            occ = [(s >> (2 * q + sp)) & 1 for q in range(ns)]
            for f in range(ns):
                if occ[f]:
                    for t in sorted(nbr[f]):
                        if not occ[t]: codes.append(code_of[(f, t)] * 2 + sp)
        lists.append(codes)
    maxlen = max(len(c) for c in lists)
    for k in range(maxlen):
        cs = [c[k] for c in lists if len(c) > k]
        banks = np.bincount(np.array(cs) % 32, minlength=32)
        # distinct addresses in the same bank conflict; identical addresses broadcast
        per_bank = [len(set(c for c in cs if c % 32 == b)) for b in range(32)]
        deg32.append(max(per_bank))
        per_bank64 = [len(set(c for c in cs if (2 * c) % 32 // 2 == b)) for b in range(16)]   # 64-bit words: 16 bank pairs
        deg64.append(max(per_bank64))
print("mean conflict degree, 32-bit tables:", np.mean(deg32), " 64-bit tables (per half-warp phases ignored):", np.mean(deg64), " samples", len(deg32))
