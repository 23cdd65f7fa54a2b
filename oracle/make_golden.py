#!/usr/bin/env python
"""oracle/make_golden.py -- TEST INFRASTRUCTURE ONLY.

Generates tests/golden/*.npz by running the UNMODIFIED reference compiled under oracle/_ref (make -C oracle ref;
needs /root/reference, i.e. this container).  Each fixture holds the reference-assembled csr_mat and the
reference's own outputs on it:
  y1          = csr_mat::MultMv(vec_randomize(seed=1))                     (src/sparse.cc:291-297)
  lanczos_a/b = lanczos(0, 999, 1000, ..., "sr_val0") coefficients + steps (src/lanczos.cc:134-266)
  E0          = hess_eigen(...)[0]                                         (src/lanczos.cc:355-390)
  cg_vec      = eigenvec_CG ground-state vector + steps (small cases)      (src/lanczos.cc:281-341)
  escale      = energy_scale(extend=0.1, iters=40)                         (src/kpm.cc:45-88)
  dn_a/dn_b   = lanczos(..., "dnmcs") with maxit=60 from vec_randomize(seed=1)
and the published golden value the case is pinned by (reference file:line).
"""
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import oracle_lib as O  # noqa: E402

CASES = {
    # name: (qb_ref case args, published golden E0 or None, citation, keep cg vector?)
    "heis12_full": (["heis_chain", 12, "none"], None, "same model as src/main_test.cc:18-111 at L=12", True),
    "heis16_full": (["heis_chain", 16, "none"], -7.142296361, "src/main_test.cc:88", False),
    "heis16_k3": (["heis_chain_k", 16, 0, 3], -5.615175598,
                  "examples/trans_symmetric/latt_chain/chain_Heisenberg_spin_half.cc:102-117 (k=3)", True),
    "tri4x4_k00": (["tri_k", 4, 4, 0, 0, 0], -8.555514918,
                   "examples/trans_symmetric/latt_triangular/triangular_Heisenberg_spin_half.cc:135", True),
    "tri4x4_k01": (["tri_k", 4, 4, 0, 0, 1], -8.002263841,
                   "examples/trans_symmetric/latt_triangular/triangular_Heisenberg_spin_half.cc:136", True),
    "tri4x4_k12": (["tri_k", 4, 4, 0, 1, 2], -7.588987242,
                   "examples/trans_symmetric/latt_triangular/triangular_Heisenberg_spin_half.cc:139 (E0_list[6])", True),
    "hubbard4x2": (["hubbard", 4, 2, 4, 4, 1, 1.1], -14.07605866,
                   "examples/trans_absent/latt_square/square_Fermi_Hubbard.cc:113", True),
    "honeycomb3x2_general": (["honeycomb", 3, 2], -28.60363167,
                             "examples/trans_absent/latt_honeycomb/honeycomb_Spinless_Fermion.cc:129", True),
    "tj12": (["tj_chain", 12, 8, 0], -9.762087307, "src/main_test.cc:207-208 (ARPACK E0=E1)", False),
    # the remaining full-basis examples of the reference: three-state sites, two orbitals, a three-site unit cell, boson amplitudes
    "spin1_chain10": (["spin_one_chain", 10, 0], -14.09412995,
                      "examples/trans_absent/latt_chain/chain_Heisenberg_spin_one.cc:96", True),
    "kondo4": (["kondo_chain", 4, 4, 1, 4], -12.67762138, "examples/trans_absent/latt_chain/chain_Kondo.cc:126", True),
    "kagome2x2_heis": (["kagome_heisenberg", 2, 2, 0], -5.444875217,
                       "examples/trans_absent/latt_kagome/kagome_Heisenberg_spin_half.cc:175", True),
    "kagome2x2_tj": (["kagome_tj", 2, 2, 8, 0], -15.41931496, "examples/trans_absent/latt_kagome/kagome_tJ.cc:232", False),
    "bose3x3": (["bose_hubbard", 3, 3, 9, 2, 1, 1.1], -25.81136094,
                "examples/trans_absent/latt_square/square_Bose_Hubbard.cc:100", True),
    # and two of their momentum sectors (generate_Ham_sparse_repr: complex Hermitian csr_mat of a three-state orbital; fermion
    # signs of the translations on a three-site unit cell)
    "spin1_chain12_k1": (["spin_one_chain_k", 12, 0, 1], -15.2458356,
                         "examples/trans_symmetric/latt_chain/chain_Heisenberg_spin_one.cc:99 (E0_list[1])", True),
    "kagome2x2_tj_k10": (["kagome_tj_k", 2, 2, 8, 0, 1, 0], -14.40277723,
                         "examples/trans_symmetric/latt_kagome/kagome_tJ.cc:239 (E0_list[1])", False),
    # CPU oracle only (dim 85: too small for a meaningful device run): spinless fermions on the honeycomb lattice, momentum (1,1)
    "honeycomb3x2_k11": (["honeycomb_k", 3, 2, 1, 1], -28.27163215,
                         "examples/trans_symmetric/latt_honeycomb/honeycomb_Spinless_Fermion.cc:139 (E0_list[3])", True),
}


def main():
    assert O.have_qb_ref(), "build oracle/_ref first: make -C oracle ref"
    only = sys.argv[1:]
    for name, (args, golden, cite, keep_cg) in CASES.items():
        if only and name not in only:
            continue
        wd = tempfile.mkdtemp(prefix="qbgold_")
        csr = os.path.join(wd, "H.qbcsr"); y1 = os.path.join(wd, "y1.bin")
        res = O.run_qb_ref(args + ["--dump", csr, "--mv", 1, y1, "--lanczos", "sr_val0", 1000,
                                   "--energy-scale", 40], threads=1, workdir=wd)
        A = O.read_qbcsr(csr)
        y = np.fromfile(y1, dtype=np.complex128)
        extra = {"y1": y, "lanczos_a": np.array(res["lanczos_a"]), "lanczos_b": np.array(res["lanczos_b"])}
        meta = {"case": name, "qb_ref_args": [str(a) for a in args], "golden_E0": golden, "golden_cite": cite,
                "lanczos_steps": res["lanczos_steps"], "lanczos_E0": res["lanczos_E0"],
                "escale_lo": res["escale_lo"], "escale_hi": res["escale_hi"], "max_abs_imag": res["max_abs_imag"]}
        if golden is not None and name != "tj12":
            assert abs(res["lanczos_E0"] - golden) < 1e-8, (name, res["lanczos_E0"], golden)
        if keep_cg:
            cgv = os.path.join(wd, "cg.bin")
            r2 = O.run_qb_ref(["file_z", csr, "--cg", repr(res["lanczos_E0"]), cgv], threads=1, workdir=wd)
            extra["cg_vec"] = np.fromfile(cgv, dtype=np.complex128)
            meta["cg_steps"] = r2["cg_steps"]; meta["cg_accuracy"] = r2["cg_accuracy"]
        r3 = O.run_qb_ref(["file_z", csr, "--lanczos", "dnmcs", 60], threads=1, workdir=wd)
        extra["dn_a"] = np.array(r3["lanczos_a"]); extra["dn_b"] = np.array(r3["lanczos_b"])
        meta["dn_steps"] = r3["lanczos_steps"]
        out = os.path.join(O.GOLDEN_DIR, name + ".npz")
        np.savez_compressed(out, dim=A.dim, ia=A.ia, ja=A.ja, val=A.val, sym=A.sym, meta=json.dumps(meta), **extra)
        print(f"{name}: dim={A.dim} nnz={A.nnz} sym={A.sym} steps={res['lanczos_steps']} E0={res['lanczos_E0']:.12f} "
              f"golden={golden} -> {os.path.getsize(out)/1e3:.0f} kB")


DYNAMIC_CASES = {
    # name: (L, Sz, k0, q, maxit) -- qb_ref heis_chain_szq: E0 and phi0 of sector k0, then A = sum_x exp(-i 2 pi q x/L)/sqrt(L) S^z_x
    # applied by model::moprXvec_repr into sector k0 - q and model::measure_repr_dynamic's Lanczos coefficients
    # (src/model.cc:1716-1846, 1897-1912)
    "heis16_szq3": (16, 0, 0, 3, 60),
    "heis16_szq8": (16, 0, 0, 8, 60),
    "heis12_szq1": (12, 0, 0, 1, 40),
    # the same flow with S^-_q (qb_ref heis_chain_smq): the target sector has one more down spin
    "heis16_smq3": (16, 0, 0, 3, 60, "heis_chain_smq"),
    "heis12_smq5": (12, 0, 0, 5, 40, "heis_chain_smq"),
}


def make_dynamic():
    only = sys.argv[2:]
    for name, spec in DYNAMIC_CASES.items():
        if only and name not in only:
            continue
        L, sz, k0, q, maxit = spec[:5]
        flow = spec[5] if len(spec) > 5 else "heis_chain_szq"
        wd = tempfile.mkdtemp(prefix="qbdyn_")
        pre = os.path.join(wd, "v")
        res = O.run_qb_ref([flow, L, sz, k0, q, maxit, "--dump-vecs", pre], threads=4, workdir=wd)
        phi0 = np.fromfile(pre + "_phi0.bin", dtype=np.complex128)
        aphi = np.fromfile(pre + "_Aphi0.bin", dtype=np.complex128)
        meta = {"case": name, "flow": flow, "L": L, "Sz": sz, "k0": k0, "q": q, "maxit": maxit, "E0": res["E0"], "dyn_norm": res["dyn_norm"],
                "dyn_steps": res["dyn_steps"]}
        out = os.path.join(O.GOLDEN_DIR, name + ".npz")
        np.savez_compressed(out, phi0=phi0, Aphi0=aphi, dyn_a=np.array(res["dyn_a"]), dyn_b=np.array(res["dyn_b"]), meta=json.dumps(meta))
        print(f"{name}: dim={phi0.size} E0={res['E0']:.12f} norm={res['dyn_norm']:.12f} steps={res['dyn_steps']} -> {os.path.getsize(out)/1e3:.0f} kB")


FULL_DYNAMIC_CASES = {
    # name: (Lx, Ly, nup, ndn, t, U, qm, qn, maxit) -- qb_ref hubbard_full_szq: the dynamic part of
    # examples/trans_absent/latt_square/square_Fermi_Hubbard.cc in the full basis (moprXvec_full + measure_full_dynamic,
    # src/model.cc:1468-1538, 1697-1712) with S^z_q = sum_r 0.5/sqrt(N) exp(i q.r) (n_up,r - n_dn,r)
    "hubbard4x2_szq10": (4, 2, 4, 4, 1.0, 1.1, 1, 0, 60),
    "hubbard4x2_szq21": (4, 2, 4, 4, 1.0, 1.1, 2, 1, 60),
}


def make_full_dynamic():
    only = sys.argv[2:]
    for name, (Lx, Ly, nup, ndn, t, U, qm, qn, maxit) in FULL_DYNAMIC_CASES.items():
        if only and name not in only:
            continue
        wd = tempfile.mkdtemp(prefix="qbfdyn_")
        pre = os.path.join(wd, "v")
        res = O.run_qb_ref(["hubbard_full_szq", Lx, Ly, nup, ndn, t, U, qm, qn, maxit, "--dump-vecs", pre], threads=4, workdir=wd)
        phi0 = np.fromfile(pre + "_phi0.bin", dtype=np.complex128)
        aphi = np.fromfile(pre + "_Aphi0.bin", dtype=np.complex128)
        meta = {"case": name, "flow": "hubbard_full_szq", "Lx": Lx, "Ly": Ly, "nup": nup, "ndn": ndn, "t": t, "U": U, "qm": qm, "qn": qn,
                "maxit": maxit, "E0": res["E0"], "dyn_norm": res["dyn_norm"], "dyn_steps": res["dyn_steps"]}
        out = os.path.join(O.GOLDEN_DIR, name + ".npz")
        np.savez_compressed(out, phi0=phi0, Aphi0=aphi, dyn_a=np.array(res["dyn_a"]), dyn_b=np.array(res["dyn_b"]), meta=json.dumps(meta))
        print(f"{name}: dim={phi0.size} E0={res['E0']:.12f} norm={res['dyn_norm']:.12f} steps={res['dyn_steps']} -> {os.path.getsize(out)/1e3:.0f} kB")


KPM_CASES = {
    # name: (golden matrix the moments are computed on, seed of phi, number of moments, energy_scale iterations)
    "kpm_heis16_k3": ("heis16_k3", 3, 64, 40),
    "kpm_tri4x4_k01": ("tri4x4_k01", 3, 64, 40),
    "kpm_hubbard4x2": ("hubbard4x2", 5, 96, 40),
}


def make_kpm():
    """Chebyshev moments on the reference's own csr_mat::MultMv and energy_scale (qb_ref --kpm): pins the KPM path, for
    which the reference has no routine of its own."""
    only = sys.argv[2:]
    for name, (mat, seed, nmom, iters) in KPM_CASES.items():
        if only and name not in only:
            continue
        A, meta0, _ = O.load_golden(mat)
        wd = tempfile.mkdtemp(prefix="qbkpm_")
        path = os.path.join(wd, "A.qbcsr")
        O.write_qbcsr(path, A)
        res = O.run_qb_ref(["file_z", path, "--kpm", seed, nmom, iters], threads=1, workdir=wd)
        meta = {"case": name, "matrix": mat, "seed": seed, "nmom": nmom, "iters": iters, "lo": res["kpm_lo"], "hi": res["kpm_hi"]}
        out = os.path.join(O.GOLDEN_DIR, name + ".npz")
        np.savez_compressed(out, moments=np.array(res["kpm_moments"]), meta=json.dumps(meta))
        print(f"{name}: lo={res['kpm_lo']:.12f} hi={res['kpm_hi']:.12f} mu[:4]={res['kpm_moments'][:4]} -> {os.path.getsize(out)/1e3:.1f} kB")


def make_digest():
    """Hubbard 4x3 (N_up = N_dn = 6; dim 853,776 -- the largest instance of BASELINE config 3's model the reference assembles
    in seconds): too large to commit whole, so a digest of the REFERENCE's own run: 4096 sampled entries of
    y = csr_mat::MultMv(vec_randomize(1)), |y|, the Lanczos coefficients, step count and E0."""
    wd = tempfile.mkdtemp(prefix="qbdig_")
    y1 = os.path.join(wd, "y1.bin")
    res = O.run_qb_ref(["hubbard", 4, 3, 6, 6, 1.0, 1.1, "--mv", 1, y1, "--lanczos", "sr_val0", 1000], threads=8, workdir=wd)
    y = np.fromfile(y1, dtype=np.complex128)
    rng = np.random.default_rng(43)
    idx = np.sort(rng.choice(y.size, size=4096, replace=False)).astype(np.int64)
    meta = {"case": "hubbard4x3_digest", "qb_ref_args": ["hubbard", 4, 3, 6, 6, 1.0, 1.1], "dim": int(res["dim"]), "nnz": int(res["nnz"]),
            "lanczos_steps": res["lanczos_steps"], "lanczos_E0": res["lanczos_E0"], "y_norm": float(np.linalg.norm(y)),
            "max_abs_imag": res["max_abs_imag"]}
    out = os.path.join(O.GOLDEN_DIR, "hubbard4x3_digest.npz")
    np.savez_compressed(out, idx=idx, y_at_idx=y[idx], lanczos_a=np.array(res["lanczos_a"]), lanczos_b=np.array(res["lanczos_b"]), meta=json.dumps(meta))
    print(f"hubbard4x3_digest: dim={res['dim']} nnz={res['nnz']} steps={res['lanczos_steps']} E0={res['lanczos_E0']:.12f} -> {os.path.getsize(out)/1e3:.0f} kB")


if __name__ == "__main__":
    if sys.argv[1:2] == ["kpm"]:
        make_kpm()
    elif sys.argv[1:2] == ["digest"]:
        make_digest()
    elif sys.argv[1:2] == ["dynamic"]:
        make_dynamic()
    elif sys.argv[1:2] == ["full_dynamic"]:
        make_full_dynamic()
    else:
        main()
