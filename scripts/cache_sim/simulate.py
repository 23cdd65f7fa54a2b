#!/usr/bin/env python
"""Offline model of the DRAM traffic of the gathered vector x in the row-ordered product (CPU only, no GPU needed).

The measured gap between the production kernel and its algorithmic roofline on BASELINE config 3 is the re-read of x
(DESIGN section 7).  This script asks whether a different internal basis order or a 2-D tiled traversal would remove
it, before anyone writes a kernel for it: it builds the single-orbital Hubbard model on the same 4x4 lattice at a
filling small enough for a trace to fit in memory (N_up = N_dn = 4: 3.3 M states, 26.6 entries per row), writes the
column trace of the product in processing order and replays it through a fully associative LRU cache of 32-byte
sectors (scripts/cache_sim/lru.c) whose size is the same FRACTION of the vector as the B200's L2 is of config 3's
vector (126 MB / 2.65 GB = 4.7 %; also half and twice that).

  gcc -O2 -o /tmp/lru scripts/cache_sim/lru.c && python scripts/cache_sim/simulate.py /tmp/lru

Result (profiles/r01_cache_sim_orders_and_tiles.txt): the reference's Lin order, a spin-species tensor order
(up-major) and plain integer order all re-read x 9-10 times; row-block x column-group tilings trade x traffic for
read-modify-write traffic on y and come out worse (11.7-18.9 against 10.1 vector sizes).  The re-read is a property of
the hopping graph (every state has ~27 neighbours spread over the whole space), not of the traversal.
"""
import itertools
import subprocess
import sys
import time

import numpy as np

LRU = sys.argv[1] if len(sys.argv) > 1 else "/tmp/lru"
TRACE = "/tmp/qb_cache_sim_trace.bin"
Lx = Ly = 4; ns = 16
nup = ndn = 4
site = lambda x, y: (x % Lx) + (y % Ly) * Lx
bonds = set()
for x in range(Lx):
    for y in range(Ly):
        for (a, b) in ((site(x, y), site(x + 1, y)), (site(x, y), site(x, y + 1))):
            bonds.add((min(a, b), max(a, b)))
bonds = sorted(bonds)
def configs(k):
    return np.array(sorted(sum(1 << i for i in c) for c in itertools.combinations(range(ns), k)), dtype=np.int64)
U, D = configs(nup), configs(ndn)
nU, nD = U.size, D.size
n = nU * nD
print("dim", n, flush=True)
# hop tables for one species: for config index c -> list of target config indices
def hop_table(C):
    idx = {int(v): i for i, v in enumerate(C)}
    src, dst = [], []
    for (a, b) in bonds:
        for (i, j) in ((a, b), (b, a)):
            m = ((C >> i) & 1 == 1) & ((C >> j) & 1 == 0)
            s = np.nonzero(m)[0]
            t = C[s] ^ ((1 << i) | (1 << j))
            src.append(s); dst.append(np.array([idx[int(v)] for v in t]))
    return np.concatenate(src), np.concatenate(dst)
hu_s, hu_d = hop_table(U); hd_s, hd_d = hop_table(D)
# full state index in "tensor" coordinates (iu, id); entries: up-hops (iu->iu', id), down-hops (iu, id->id'), diagonal
def spread(v):   # config bits -> bit 2s
    out = np.zeros_like(v)
    for s in range(ns): out |= ((v >> s) & 1) << (2 * s)
    return out
Us, Ds = spread(U), spread(D) << 1
def lin_key(iu, idn):
    st = Us[iu] | Ds[idn]
    la = np.zeros_like(st); lb = np.zeros_like(st)
    for s in range(ns):
        d = (st >> (2 * s)) & 3
        if s % 2 == 0: la |= d << (2 * (s // 2))
        else: lb |= d << (2 * (s // 2))
    return (lb << 32) | la
iu_all = np.repeat(np.arange(nU), nD); id_all = np.tile(np.arange(nD), nU)
orders = {}
t0 = time.time()
key = lin_key(iu_all, id_all)
pos = np.empty(n, dtype=np.int64); pos[np.argsort(key, kind="stable")] = np.arange(n); orders["lin"] = pos
orders["tensor_up_major"] = iu_all * nD + id_all
st = Us[iu_all] | Ds[id_all]
pos = np.empty(n, dtype=np.int64); pos[np.argsort(st, kind="stable")] = np.arange(n); orders["integer"] = pos
print("orders built", time.time() - t0, flush=True)
def trace_for(pos):
    # rows in position order; build (rowpos, colpos) for all entries then sort by (rowpos, colpos)
    R, C = [pos], [pos]                                   # diagonal
    # up hops: for every id
    r = (hu_s[:, None] * nD + np.arange(nD)[None, :]).ravel(); c = (hu_d[:, None] * nD + np.arange(nD)[None, :]).ravel()
    R.append(pos[r]); C.append(pos[c])
    r = (np.arange(nU)[:, None] * nD + hd_s[None, :]).ravel(); c = (np.arange(nU)[:, None] * nD + hd_d[None, :]).ravel()
    R.append(pos[r]); C.append(pos[c])
    R = np.concatenate(R); C = np.concatenate(C)
    o = np.lexsort((C, R))
    return R[o], C[o].astype(np.int32)
res = {}
for name, pos in orders.items():
    R, C = trace_for(pos)
    nnz = C.size
    far = np.abs(C.astype(np.int64) - R) 
    with open(TRACE, "wb") as f:
        f.write(np.int64(nnz).tobytes()); f.write(C.tobytes())
    for frac in (0.024, 0.047, 0.10):
        for eps, label in ((2, "complex"), (4, "fp64")):
            sectors_total = n // eps
            cap = max(64, int(frac * sectors_total))
            out = subprocess.run([LRU, TRACE, str(cap), str(eps)], capture_output=True, text=True).stdout.split()
            misses = int(out[1])
            res[(name, frac, label)] = misses / sectors_total
            print(f"{name:18s} cache {frac*100:4.1f}% of x  {label:7s}: x traffic = {misses / sectors_total:6.2f} x vector size   (nnz/row {nnz/n:.1f}, far>1% of n: {np.mean(far > 0.01*n)*100:.0f}%)", flush=True)

# ---- 2-D tiled traversals of the Lin-ordered matrix (x misses + read-modify-write of y)
pos = orders["lin"]
R, C = trace_for(pos)
C = C.astype(np.int64)
nnz = C.size
eps = 2
sectors_total = n // eps
def simulate(trace, cap):
    with open(TRACE, "wb") as f:
        f.write(np.int64(trace.size).tobytes()); f.write(trace.astype(np.int32).tobytes())
    out = subprocess.run([LRU, TRACE, str(cap), str(eps)], capture_output=True, text=True).stdout.split()
    return int(out[1])
for frac in (0.047,):
    cap = int(frac * sectors_total)
    base = simulate(C, cap)
    print(f"untiled: x {base / sectors_total:.2f}  + y write 1.00  => {base / sectors_total + 1:.2f} vector sizes", flush=True)
    for (nrb, ncg) in ((16, 16), (32, 32), (64, 64), (32, 8), (128, 16), (64, 16), (256, 32)):
        rb = R // ((n + nrb - 1) // nrb); cg = C // ((n + ncg - 1) // ncg)
        o = np.lexsort((C, R, cg, rb))
        Rt, Ct, cgt, rbt = R[o], C[o], cg[o], rb[o]
        first = np.ones(nnz, dtype=bool); first[1:] = (Rt[1:] != Rt[:-1]) | (cgt[1:] != cgt[:-1])
        # interleave a y access (address n + 2*row to keep y sectors distinct from x: use a separate id space) before each run
        ins = np.nonzero(first)[0]
        trace = np.insert(Ct, ins, n + Rt[ins] + (n % 2))
        total_miss = simulate(trace, cap)
        # split: simulate x-only with same order to attribute
        xm = simulate(Ct, cap)
        ym = total_miss - xm
        print(f"tiles {nrb:4d} x {ncg:3d}: x {xm / sectors_total:.2f}  y-misses {ym / sectors_total:.2f} (rmw => x2)  total ~ {(xm + 2 * ym) / sectors_total:.2f} vector sizes; (row,group) visits per row {first.sum() / n:.1f}", flush=True)
