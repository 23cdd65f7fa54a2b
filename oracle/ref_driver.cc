/* oracle/ref_driver.cc -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A command-line driver over the UNMODIFIED reference library compiled from /root/reference/src
 * (see oracle/Makefile).  It builds the reference's own models through the reference's public API
 * (the way examples/ and src/main_test.cc do), and
 *   - dumps the reference-assembled csr_mat (dim, nnz, sym, ia, ja, val) to a .qbcsr file,
 *   - runs the reference's locate_E0_lanczos / lanczos / eigenvec_CG / energy_scale on it,
 *   - times the reference's csr_mat<T>::MultMv and lanczos step (CPU baseline),
 *   - or loads a .qbcsr produced elsewhere (e.g. by the product's builders) into a reference csr_mat
 *     and runs the same reference routines on it.
 * Results are written as one JSON object to --out.  The reference's own chatter goes to stdout.
 *
 * .qbcsr layout (little endian): char magic[8]="QBCSR1\0\0"; int64 dim; int64 nnz; int32 sym;
 * int32 is_complex; int64 ia[dim+1]; int64 ja[nnz]; (double | double[2]) val[nnz].
 */
#include <cassert>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>
#include <algorithm>
#include <sstream>
#include <unistd.h>
#include <omp.h>
#include "qbasis.h"

using cplx = std::complex<double>;
using Model = qbasis::model<cplx>;
using Opr = qbasis::opr<cplx>;
using Mopr = qbasis::mopr<cplx>;

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

/* ----------------------------------------------------------------- qbcsr i/o */
template <typename T>
static void dump_csr(const qbasis::csr_mat<T> &A, const std::string &path) {
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) { perror(path.c_str()); exit(2); }
    char magic[8] = {'Q','B','C','S','R','1',0,0};
    int64_t dim = A.dim, nnz = A.nnz;
    int32_t sym = A.sym ? 1 : 0, isc = sizeof(T) == sizeof(cplx) ? 1 : 0;
    fwrite(magic, 1, 8, f); fwrite(&dim, 8, 1, f); fwrite(&nnz, 8, 1, f); fwrite(&sym, 4, 1, f); fwrite(&isc, 4, 1, f);
    fwrite(A.ia, 8, dim + 1, f); fwrite(A.ja, 8, nnz, f); fwrite(A.val, sizeof(T), nnz, f);
    fclose(f);
}

template <typename T>
static void load_csr(qbasis::csr_mat<T> &A, const std::string &path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { perror(path.c_str()); exit(2); }
    char magic[8]; int64_t dim, nnz; int32_t sym, isc;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "QBCSR1", 6) != 0) { fprintf(stderr, "bad magic\n"); exit(2); }
    if (fread(&dim, 8, 1, f) != 1 || fread(&nnz, 8, 1, f) != 1 || fread(&sym, 4, 1, f) != 1 || fread(&isc, 4, 1, f) != 1) exit(2);
    if ((isc != 0) != (sizeof(T) == sizeof(cplx))) { fprintf(stderr, "scalar type mismatch\n"); exit(2); }
    A.dim = dim; A.nnz = nnz; A.sym = sym != 0;
    A.ia = new MKL_INT[dim + 1]; A.ja = new MKL_INT[nnz]; A.val = new T[nnz];
    if (fread(A.ia, 8, dim + 1, f) != (size_t)(dim + 1) || fread(A.ja, 8, nnz, f) != (size_t)nnz ||
        fread(A.val, sizeof(T), nnz, f) != (size_t)nnz) { fprintf(stderr, "short file\n"); exit(2); }
    fclose(f);
    /* the same handle creation csr_mat's own constructors perform (reference src/sparse.cc:129,258) */
    sparse_status_t st;
    if constexpr (sizeof(T) == sizeof(cplx)) st = mkl_sparse_z_create_csr(&A.handle, SPARSE_INDEX_BASE_ZERO, dim, dim, A.ia, A.ia + 1, A.ja, (MKL_Complex16 *)A.val);
    else                                     st = mkl_sparse_d_create_csr(&A.handle, SPARSE_INDEX_BASE_ZERO, dim, dim, A.ia, A.ia + 1, A.ja, (double *)A.val);
    if (st != SPARSE_STATUS_SUCCESS) { fprintf(stderr, "create_handle failed\n"); exit(2); }
}

static void dump_vec(const std::string &path, const void *p, size_t bytes) {
    FILE *f = fopen(path.c_str(), "wb"); if (!f) { perror(path.c_str()); exit(2); }
    fwrite(p, 1, bytes, f); fclose(f);
}

/* ------------------------------------------------------------ local matrices */
static std::vector<std::vector<cplx>> mat2(cplx a00, cplx a01, cplx a10, cplx a11) { return {{a00, a01}, {a10, a11}}; }

struct Built {
    std::unique_ptr<Model> model;
    qbasis::which_sym sym = qbasis::which_sym::full;
    std::string name;
};

static void add_heisenberg_bond(Model &M, uint32_t i, uint32_t j, double J) {
    auto Sp = mat2(0, 1, 0, 0), Sm = mat2(0, 0, 1, 0);
    std::vector<cplx> Sz{0.5, -0.5};
    Opr Spi(i, 0, false, Sp), Smi(i, 0, false, Sm), Szi(i, 0, false, Sz);
    Opr Spj(j, 0, false, Sp), Smj(j, 0, false, Sm), Szj(j, 0, false, Sz);
    M.add_Ham(cplx(0.5 * J, 0.0) * (Spi * Smj + Smi * Spj));
    M.add_Ham(cplx(J, 0.0) * (Szi * Szj));
}

static Mopr total_sz(uint32_t nsites) {
    std::vector<cplx> Sz{0.5, -0.5};
    Mopr S;
    for (uint32_t s = 0; s < nsites; s++) S += Opr(s, 0, false, Sz);
    return S;
}

/* Heisenberg chain, PBC.  mode "none": no conserved quantity (main_test.cc:18-111); "sz": Sz_total = szval. */
static Built build_heis_chain(int L, bool use_sz, double szval, int k /* -1: no translation */) {
    Built b; b.name = "heis_chain";
    qbasis::lattice latt("chain", {static_cast<uint32_t>(L)}, {"pbc"});
    b.model = std::make_unique<Model>(latt);
    Model &M = *b.model;
    M.add_orbital(latt.Nsites, "spin-1/2");
    for (int x = 0; x < L; x++) {
        uint32_t si, sj; std::vector<int> work(latt.dim);
        latt.coor2site({x}, 0, si, work); latt.coor2site({x + 1}, 0, sj, work);
        add_heisenberg_bond(M, si, sj, 1.0);
    }
    if (k < 0) {
        if (use_sz) M.enumerate_basis_full({total_sz(latt.Nsites)}, {szval}); else M.enumerate_basis_full({}, {});
        M.generate_Ham_sparse_full();
        b.sym = qbasis::which_sym::full;
    } else {
        M.fill_Weisse_table();
        if (use_sz) M.enumerate_basis_repr({k}, {total_sz(latt.Nsites)}, {szval}); else M.enumerate_basis_repr({k}, {}, {});
        M.generate_Ham_sparse_repr();
        b.sym = qbasis::which_sym::repr;
    }
    return b;
}

/* Heisenberg on the triangular lattice, PBC, Sz_total = szval, momentum (m,n) or full basis (m<0). */
static Built build_triangular(int Lx, int Ly, double szval, int m, int n) {
    Built b; b.name = "triangular";
    qbasis::lattice latt("triangular", {static_cast<uint32_t>(Lx), static_cast<uint32_t>(Ly)}, {"pbc", "pbc"});
    b.model = std::make_unique<Model>(latt);
    Model &M = *b.model;
    M.add_orbital(latt.Nsites, "spin-1/2");
    for (int x = 0; x < Lx; x++) for (int y = 0; y < Ly; y++) {
        uint32_t si, sj; std::vector<int> work(latt.dim);
        latt.coor2site({x, y}, 0, si, work);
        latt.coor2site({x + 1, y}, 0, sj, work);     add_heisenberg_bond(M, si, sj, 1.0);
        latt.coor2site({x + 1, y + 1}, 0, sj, work); add_heisenberg_bond(M, si, sj, 1.0);
        latt.coor2site({x, y + 1}, 0, sj, work);     add_heisenberg_bond(M, si, sj, 1.0);
    }
    if (m < 0) {
        M.enumerate_basis_full({total_sz(latt.Nsites)}, {szval});
        M.generate_Ham_sparse_full();
        b.sym = qbasis::which_sym::full;
    } else {
        M.fill_Weisse_table();
        M.enumerate_basis_repr({m, n}, {total_sz(latt.Nsites)}, {szval});
        M.generate_Ham_sparse_repr();
        b.sym = qbasis::which_sym::repr;
    }
    return b;
}

/* Fermi-Hubbard on the square lattice, PBC (examples/trans_absent/latt_square/square_Fermi_Hubbard.cc). */
static Built build_hubbard(int Lx, int Ly, double nup, double ndn, double t, double U, int km = -1, int kn = -1) {
    Built b; b.name = "hubbard";
    qbasis::lattice latt("square", {static_cast<uint32_t>(Lx), static_cast<uint32_t>(Ly)}, {"pbc", "pbc"});
    b.model = std::make_unique<Model>(latt);
    Model &M = *b.model;
    auto cu = std::vector<std::vector<cplx>>(4, std::vector<cplx>(4, 0.0));
    auto cd = cu;
    cu[0][1] = 1.0; cu[2][3] = 1.0; cd[0][2] = 1.0; cd[1][3] = -1.0;
    M.add_orbital(latt.Nsites, "electron");
    Mopr Nup, Ndn;
    auto hop = [&](uint32_t i, uint32_t j) {
        Opr cui(i, 0, true, cu), cdi(i, 0, true, cd), cuj(j, 0, true, cu), cdj(j, 0, true, cd);
        auto cuid = cui; cuid.dagger(); auto cdid = cdi; cdid.dagger();
        auto cujd = cuj; cujd.dagger(); auto cdjd = cdj; cdjd.dagger();
        M.add_Ham(cplx(-t, 0.0) * (cuid * cuj)); M.add_Ham(cplx(-t, 0.0) * (cujd * cui));
        M.add_Ham(cplx(-t, 0.0) * (cdid * cdj)); M.add_Ham(cplx(-t, 0.0) * (cdjd * cdi));
    };
    for (int x = 0; x < Lx; x++) for (int y = 0; y < Ly; y++) {
        uint32_t si, sj; std::vector<int> work(latt.dim);
        latt.coor2site({x, y}, 0, si, work);
        Opr cui(si, 0, true, cu), cdi(si, 0, true, cd);
        auto cuid = cui; cuid.dagger(); auto cdid = cdi; cdid.dagger();
        auto nu = cuid * cui; auto nd = cdid * cdi;
        latt.coor2site({x + 1, y}, 0, sj, work); hop(si, sj);
        latt.coor2site({x, y + 1}, 0, sj, work); hop(si, sj);
        M.add_Ham(cplx(U, 0.0) * (nu * nd));
        Nup += nu; Ndn += nd;
    }
    if (km < 0) {
        M.enumerate_basis_full({Nup, Ndn}, {nup, ndn});
        M.generate_Ham_sparse_full();
    } else {                 /* momentum sector (examples/trans_symmetric/latt_square/square_Fermi_Hubbard.cc:95-104) */
        M.fill_Weisse_table();
        M.enumerate_basis_repr({km, kn}, {Nup, Ndn}, {nup, ndn});
        M.generate_Ham_sparse_repr();
        b.sym = qbasis::which_sym::repr;
    }
    return b;
}

/* t-J chain (src/main_test.cc:113-211). */
static Built build_tj_chain(int L, double ntot, double sz) {
    Built b; b.name = "tj_chain";
    qbasis::lattice latt("chain", {static_cast<uint32_t>(L)}, {"pbc"});
    b.model = std::make_unique<Model>(latt);
    Model &M = *b.model;
    auto cu = std::vector<std::vector<cplx>>(3, std::vector<cplx>(3, 0.0));
    auto cd = cu; cu[0][1] = 1.0; cd[0][2] = 1.0;
    M.add_orbital(latt.Nsites, "tJ");
    Mopr Sz_tot, N_tot;
    const double t = 1.0, J = 1.0;
    for (int m = 0; m < L; m++) {
        uint32_t si, sj; std::vector<int> work(latt.dim);
        latt.coor2site({m}, 0, si, work); latt.coor2site({m + 1}, 0, sj, work);
        Opr cui(si, 0, true, cu), cdi(si, 0, true, cd), cuj(sj, 0, true, cu), cdj(sj, 0, true, cd);
        auto cuid = cui; cuid.dagger(); auto cdid = cdi; cdid.dagger();
        auto cujd = cuj; cujd.dagger(); auto cdjd = cdj; cdjd.dagger();
        auto Spi = cuid * cdi; auto Smi = cdid * cui; auto Spj = cujd * cdj; auto Smj = cdjd * cuj;
        auto Szi = cplx(0.5, 0.0) * (cuid * cui - cdid * cdi); auto Szj = cplx(0.5, 0.0) * (cujd * cuj - cdjd * cdj);
        auto Ni = (cuid * cui + cdid * cdi); auto Nj = (cujd * cuj + cdjd * cdj);
        M.add_Ham(cplx(-t, 0.0) * (cuid * cuj)); M.add_Ham(cplx(-t, 0.0) * (cujd * cui));
        M.add_Ham(cplx(-t, 0.0) * (cdid * cdj)); M.add_Ham(cplx(-t, 0.0) * (cdjd * cdi));
        M.add_Ham(cplx(0.5 * J, 0.0) * (Spi * Smj + Smi * Spj));
        M.add_Ham(cplx(J, 0.0) * (Szi * Szj));
        M.add_Ham(cplx(-0.25 * J, 0.0) * (Ni * Nj));
        Sz_tot += Szi; N_tot += Ni;
    }
    M.enumerate_basis_full({Sz_tot, N_tot}, {sz, ntot});
    M.generate_Ham_sparse_full();
    return b;
}

/* Spinless fermions on the honeycomb lattice, stored as a GENERAL (full, non-symmetric-storage) CSR
 * (examples/trans_absent/latt_honeycomb/honeycomb_Spinless_Fermion.cc). */
static Built build_honeycomb(int Lx, int Ly, int km = -1, int kn = -1 /* >= 0: momentum sector (trans_symmetric/latt_honeycomb) */) {
    Built b; b.name = "honeycomb";
    const double t = 1.0, V1 = 4.0;
    qbasis::lattice latt("honeycomb", {static_cast<uint32_t>(Lx), static_cast<uint32_t>(Ly)}, {"pbc", "pbc"});
    b.model = std::make_unique<Model>(latt);
    Model &M = *b.model;
    auto c = std::vector<std::vector<cplx>>(2, std::vector<cplx>(2, 0.0)); c[0][1] = 1.0;
    M.add_orbital(latt.Nsites, "spinless-fermion");
    Mopr Nf;
    for (int x = 0; x < Lx; x++) for (int y = 0; y < Ly; y++) {
        uint32_t si, sj; std::vector<int> work(latt.dim);
        latt.coor2site({x, y}, 0, si, work);
        Opr ci(si, 0, true, c); auto cid = ci; cid.dagger(); auto ni = cid * ci;
        const int nb[3][2] = {{x, y}, {x - 1, y}, {x - 1, y - 1}};
        for (auto &r : nb) {
            latt.coor2site({r[0], r[1]}, 1, sj, work);
            Opr cj(sj, 0, true, c); auto cjd = cj; cjd.dagger(); auto nj = cjd * cj;
            M.add_Ham(cplx(-t, 0.0) * (cid * cj)); M.add_Ham(cplx(-t, 0.0) * (cjd * ci));
            M.add_Ham(cplx(V1, 0.0) * (ni * nj)); M.add_Ham(cplx(-0.5 * V1, 0.0) * (ni + nj));
        }
        latt.coor2site({x, y}, 1, sj, work);
        Opr cj(sj, 0, true, c); auto cjd = cj; cjd.dagger();
        Nf += (ni + cjd * cj);
    }
    if (km < 0) { M.enumerate_basis_full({Nf}, {double(Lx * Ly - 2)}); M.generate_Ham_sparse_full(0, false); }
    else { M.fill_Weisse_table(); M.enumerate_basis_repr({km, kn}, {Nf}, {double(Lx * Ly - 2)}); M.generate_Ham_sparse_repr(); b.sym = qbasis::which_sym::repr; }
    return b;
}

/* ------------------------------------------------------------------ actions */

/* Config-5 style flow of the reference, end to end (examples/trans_symmetric/latt_chain/chain_Heisenberg_spin_one_excitation.cc:
 * 95-128 with S^z instead of S^-): ground state of the sector (Sz, k0), then for momentum transfer q the sector k0 - q is
 * enumerated as sector 1, A = sum_x exp(-i 2 pi q x / L) / sqrt(L) S^z_x is applied to phi0 (model::moprXvec_repr,
 * src/model.cc:1716-1846) and model::measure_repr_dynamic (src/model.cc:1897-1912) returns the Lanczos coefficients. */
/* ---- the remaining full-basis examples of the reference (examples/trans_absent): the published E0 of each pins the oracle
 * on one more kind of matrix (three-state sites, two orbitals, a three-site unit cell, boson amplitudes sqrt(n)) ---- */

/* spin-1 Heisenberg chain, PBC, Sz_total = szval (chain_Heisenberg_spin_one.cc: L = 10, Sz = 0 -> -14.09412995) */
static Built build_spin_one_chain(int L, double szval, int k = -1 /* >= 0: momentum sector (trans_symmetric/.../chain_Heisenberg_spin_one.cc) */) {
    Built b; b.name = "spin_one_chain";
    qbasis::lattice latt("chain", {static_cast<uint32_t>(L)}, {"pbc"});
    b.model = std::make_unique<Model>(latt);
    Model &M = *b.model;
    const double h = 1.0 / sqrt(2.0);
    std::vector<std::vector<cplx>> Sx(3, std::vector<cplx>(3, 0.0)), Sy = Sx;
    Sx[0][1] = Sx[1][0] = Sx[1][2] = Sx[2][1] = h;
    Sy[0][1] = Sy[1][2] = cplx(0.0, -h); Sy[1][0] = Sy[2][1] = cplx(0.0, h);
    std::vector<cplx> Sz{1.0, 0.0, -1.0};
    M.add_orbital(latt.Nsites, "spin-1");
    Mopr Sz_tot;
    for (int x = 0; x < L; x++) {
        uint32_t si, sj; std::vector<int> work(latt.dim);
        latt.coor2site({x}, 0, si, work); latt.coor2site({x + 1}, 0, sj, work);
        Opr Sxi(si, 0, false, Sx), Syi(si, 0, false, Sy), Szi(si, 0, false, Sz);
        Opr Sxj(sj, 0, false, Sx), Syj(sj, 0, false, Sy), Szj(sj, 0, false, Sz);
        M.add_Ham(cplx(1.0, 0.0) * (Sxi * Sxj + Syi * Syj));
        M.add_Ham(cplx(1.0, 0.0) * (Szi * Szj));
        Sz_tot += Szi;
    }
    if (k < 0) { M.enumerate_basis_full({Sz_tot}, {szval}); M.generate_Ham_sparse_full(); }
    else { M.fill_Weisse_table(); M.enumerate_basis_repr({k}, {Sz_tot}, {szval}); M.generate_Ham_sparse_repr(); b.sym = qbasis::which_sym::repr; }
    return b;
}

/* Kondo chain, PBC: electrons (orbital 0) coupled to local spins 1/2 (orbital 1), N_elec = nelec
 * (chain_Kondo.cc: L = 4, t = 1, J_Kondo = 4, N_elec = L -> -12.67762138) */
static Built build_kondo_chain(int L, double nelec, double t, double JK) {
    Built b; b.name = "kondo_chain";
    qbasis::lattice latt("chain", {static_cast<uint32_t>(L)}, {"pbc"});
    b.model = std::make_unique<Model>(latt);
    Model &M = *b.model;
    auto cu = std::vector<std::vector<cplx>>(4, std::vector<cplx>(4, 0.0));
    auto cd = cu;
    cu[0][1] = 1.0; cu[2][3] = 1.0; cd[0][2] = 1.0; cd[1][3] = -1.0;
    auto Sp = mat2(0, 1, 0, 0), Sm = mat2(0, 0, 1, 0);
    std::vector<cplx> Sz{0.5, -0.5};
    M.add_orbital(latt.Nsites, "electron");
    M.add_orbital(latt.Nsites, "spin-1/2");
    Mopr N_tot;
    for (int x = 0; x < L; x++) {
        uint32_t si, sj; std::vector<int> work(latt.dim);
        latt.coor2site({x}, 0, si, work); latt.coor2site({x + 1}, 0, sj, work);
        Opr cui(si, 0, true, cu), cdi(si, 0, true, cd), cuj(sj, 0, true, cu), cdj(sj, 0, true, cd);
        auto cuid = cui; cuid.dagger(); auto cdid = cdi; cdid.dagger();
        auto cujd = cuj; cujd.dagger(); auto cdjd = cdj; cdjd.dagger();
        M.add_Ham(cplx(-t, 0.0) * (cuid * cuj)); M.add_Ham(cplx(-t, 0.0) * (cujd * cui));
        M.add_Ham(cplx(-t, 0.0) * (cdid * cdj)); M.add_Ham(cplx(-t, 0.0) * (cdjd * cdi));
        auto spi = cuid * cdi; auto smi = cdid * cui;                    // the electron's spin on site i
        auto szi = cplx(0.5, 0.0) * (cuid * cui - cdid * cdi);
        Opr Spi(si, 1, false, Sp), Smi(si, 1, false, Sm), Szi(si, 1, false, Sz);      // the local spin on site i
        M.add_Ham(cplx(0.5 * JK, 0.0) * (Spi * smi + Smi * spi));
        M.add_Ham(cplx(JK, 0.0) * (Szi * szi));
        N_tot += (cuid * cui + cdid * cdi);
    }
    M.enumerate_basis_full({N_tot}, {nelec});
    M.generate_Ham_sparse_full();
    return b;
}

/* the six bonds of one kagome unit cell (m, n), as (sublattice, cell) pairs -- the bonds of the reference's two kagome examples */
struct KagomeBond { int sa, sb, dm, dn; };       // site (m, n; sa) -- site (m + dm, n + dn; sb)
static const KagomeBond kKagomeBonds[6] = {{0, 2, 1, 0}, {0, 2, 0, 0}, {1, 0, 0, 1}, {1, 0, 0, 0}, {2, 1, -1, -1}, {2, 1, 0, 0}};

/* spin-1/2 Heisenberg on the kagome lattice, PBC (kagome_Heisenberg_spin_half.cc: 2 x 2, Sz = 0 -> -5.444875217) */
static Built build_kagome_heisenberg(int Lx, int Ly, double szval) {
    Built b; b.name = "kagome_heisenberg";
    qbasis::lattice latt("kagome", {static_cast<uint32_t>(Lx), static_cast<uint32_t>(Ly)}, {"pbc", "pbc"});
    b.model = std::make_unique<Model>(latt);
    Model &M = *b.model;
    M.add_orbital(latt.Nsites, "spin-1/2");
    for (int m = 0; m < Lx; m++) for (int n = 0; n < Ly; n++)
        for (const auto &kb : kKagomeBonds) {
            uint32_t si, sj; std::vector<int> work(latt.dim);
            latt.coor2site({m, n}, kb.sa, si, work); latt.coor2site({m + kb.dm, n + kb.dn}, kb.sb, sj, work);
            add_heisenberg_bond(M, si, sj, 1.0);
        }
    M.enumerate_basis_full({total_sz(latt.Nsites)}, {szval});
    M.generate_Ham_sparse_full();
    return b;
}

/* t-J model on the kagome lattice, PBC (kagome_tJ.cc: 2 x 2, N = 8, Sz = 0, t = J = 1 -> -15.41931496) */
static Built build_kagome_tj(int Lx, int Ly, double ntot, double szval, int km = -1, int kn = -1 /* >= 0: momentum sector (trans_symmetric/latt_kagome/kagome_tJ.cc) */) {
    Built b; b.name = "kagome_tj";
    qbasis::lattice latt("kagome", {static_cast<uint32_t>(Lx), static_cast<uint32_t>(Ly)}, {"pbc", "pbc"});
    b.model = std::make_unique<Model>(latt);
    Model &M = *b.model;
    auto cu = std::vector<std::vector<cplx>>(3, std::vector<cplx>(3, 0.0));
    auto cd = cu; cu[0][1] = 1.0; cd[0][2] = 1.0;
    M.add_orbital(latt.Nsites, "tJ");
    const double t = 1.0, J = 1.0;
    struct Site { Opr cu, cd, cud, cdd; Mopr sp, sm, sz, n; };
    auto site_ops = [&](uint32_t s) {
        Opr c_u(s, 0, true, cu), c_d(s, 0, true, cd);
        auto c_ud = c_u; c_ud.dagger(); auto c_dd = c_d; c_dd.dagger();
        return Site{c_u, c_d, c_ud, c_dd, Mopr(c_ud * c_d), Mopr(c_dd * c_u), cplx(0.5, 0.0) * (c_ud * c_u - c_dd * c_d), c_ud * c_u + c_dd * c_d};
    };
    Mopr Sz_tot, N_tot;
    for (int m = 0; m < Lx; m++) for (int n = 0; n < Ly; n++) {
        std::vector<int> work(latt.dim);
        for (const auto &kb : kKagomeBonds) {
            uint32_t si, sj;
            latt.coor2site({m, n}, kb.sa, si, work); latt.coor2site({m + kb.dm, n + kb.dn}, kb.sb, sj, work);
            Site a = site_ops(si), c = site_ops(sj);
            M.add_Ham(cplx(-t, 0.0) * (a.cud * c.cu)); M.add_Ham(cplx(-t, 0.0) * (c.cud * a.cu));
            M.add_Ham(cplx(-t, 0.0) * (a.cdd * c.cd)); M.add_Ham(cplx(-t, 0.0) * (c.cdd * a.cd));
            M.add_Ham(cplx(0.5 * J, 0.0) * (a.sp * c.sm + a.sm * c.sp));
            M.add_Ham(cplx(J, 0.0) * (a.sz * c.sz));
            M.add_Ham(cplx(-0.25 * J, 0.0) * (a.n * c.n));
        }
        for (int sub = 0; sub < 3; sub++) {
            uint32_t s; latt.coor2site({m, n}, sub, s, work);
            Site a = site_ops(s);
            Sz_tot += a.sz; N_tot += a.n;
        }
    }
    if (km < 0) { M.enumerate_basis_full({Sz_tot, N_tot}, {szval, ntot}); M.generate_Ham_sparse_full(); }
    else { M.fill_Weisse_table(); M.enumerate_basis_repr({km, kn}, {Sz_tot, N_tot}, {szval, ntot}); M.generate_Ham_sparse_repr(); b.sym = qbasis::which_sym::repr; }
    return b;
}

/* Bose-Hubbard model on the square lattice, PBC, at most nmax bosons per site
 * (square_Bose_Hubbard.cc: 3 x 3, t = 1, U = 1.1, N = 9, Nmax = 2 -> -25.81136094) */
static Built build_bose_hubbard(int Lx, int Ly, double ntot, int nmax, double t, double U) {
    Built b; b.name = "bose_hubbard";
    qbasis::lattice latt("square", {static_cast<uint32_t>(Lx), static_cast<uint32_t>(Ly)}, {"pbc", "pbc"});
    b.model = std::make_unique<Model>(latt);
    Model &M = *b.model;
    qbasis::extra_info limit; limit.Nmax = static_cast<uint8_t>(nmax);
    auto bm = std::vector<std::vector<cplx>>(nmax + 1, std::vector<cplx>(nmax + 1, 0.0));
    for (int d = 0; d < nmax; d++) bm[d][d + 1] = cplx(sqrt(double(d + 1)), 0.0);
    M.add_orbital(latt.Nsites, "boson", limit);
    Mopr N_tot;
    for (int x = 0; x < Lx; x++) for (int y = 0; y < Ly; y++) {
        uint32_t si; std::vector<int> work(latt.dim);
        latt.coor2site({x, y}, 0, si, work);
        Opr bi(si, 0, false, bm); auto bid = bi; bid.dagger(); auto ni = bid * bi;
        const int nb[2][2] = {{x + 1, y}, {x, y + 1}};
        for (auto &r : nb) {
            uint32_t sj; latt.coor2site({r[0], r[1]}, 0, sj, work);
            Opr bj(sj, 0, false, bm); auto bjd = bj; bjd.dagger();
            M.add_Ham(cplx(-t, 0.0) * (bid * bj)); M.add_Ham(cplx(-t, 0.0) * (bjd * bi));
        }
        M.add_Ham(cplx(0.5 * U, 0.0) * (ni * ni - ni));
        N_tot += ni;
    }
    M.enumerate_basis_full({N_tot}, {ntot});
    M.generate_Ham_sparse_full();
    return b;
}

static void flow_heis_chain_szq(int L, double szval, int k0, int q, MKL_INT maxit, const std::string &vec_prefix, struct Json &js, char kind = 'z');

struct Json {
    std::ostringstream o; bool first = true;
    Json() { o << std::setprecision(17) << "{"; }
    void key(const std::string &k) { if (!first) o << ", "; first = false; o << "\"" << k << "\": "; }
    void num(const std::string &k, double v) { key(k); o << v; }
    void integer(const std::string &k, long long v) { key(k); o << v; }
    void str(const std::string &k, const std::string &v) { key(k); o << "\"" << v << "\""; }
    void arr(const std::string &k, const double *p, long long n) { key(k); o << "["; for (long long i = 0; i < n; i++) o << (i ? ", " : "") << p[i]; o << "]"; }
    std::string done() { o << "}"; return o.str(); }
};


static void flow_heis_chain_szq(int L, double szval, int k0, int q, MKL_INT maxit, const std::string &vec_prefix, Json &js, char kind)
{
    qbasis::lattice latt("chain", {static_cast<uint32_t>(L)}, {"pbc"});
    Model M(latt);
    M.add_orbital(latt.Nsites, "spin-1/2");
    for (int x = 0; x < L; x++) {
        uint32_t si, sj; std::vector<int> work(latt.dim);
        latt.coor2site({x}, 0, si, work); latt.coor2site({x + 1}, 0, sj, work);
        add_heisenberg_bond(M, si, sj, 1.0);
    }
    M.fill_Weisse_table();
    M.enumerate_basis_repr({k0}, {total_sz(latt.Nsites)}, {szval});
    M.generate_Ham_sparse_repr();
    M.locate_E0_lanczos(qbasis::which_sym::repr, 1, 1);
    js.num("E0", M.eigenvals_repr[0]);
    js.integer("dim0", M.dim_repr[0]);
    std::vector<cplx> Sz{0.5, -0.5};
    Mopr Szq;
    const double Q = 2.0 * 3.1415926535897932 * q / static_cast<double>(L);
    for (int x = 0; x < L; x++) {
        uint32_t si; std::vector<int> work(latt.dim);
        latt.coor2site({x}, 0, si, work);
        if (kind == 'm') Szq += (std::exp(cplx(0.0, -Q * x)) / std::sqrt(static_cast<double>(L))) * Opr(si, 0, false, mat2(0, 0, 1, 0));   /* S^-_x */
        else             Szq += (std::exp(cplx(0.0, -Q * x)) / std::sqrt(static_cast<double>(L))) * Opr(si, 0, false, Sz);
    }
    /* S^- lowers Sz by one (examples/trans_symmetric/latt_chain/chain_Heisenberg_spin_one_excitation.cc:121) */
    M.enumerate_basis_repr({k0 - q}, {total_sz(latt.Nsites)}, {kind == 'm' ? szval - 1.0 : szval}, 1);
    M.generate_Ham_sparse_repr(1);
    M.switch_sec_mat(1);
    js.integer("dim1", M.dim_repr[1]);
    if (!vec_prefix.empty()) {
        std::vector<cplx> vnew(M.dim_repr[1]);
        M.moprXvec_repr(Szq, 0, 1, static_cast<MKL_INT>(0), vnew.data());
        dump_vec(vec_prefix + "_phi0.bin", M.eigenvecs_repr.data(), sizeof(cplx) * M.dim_repr[0]);
        dump_vec(vec_prefix + "_Aphi0.bin", vnew.data(), sizeof(cplx) * M.dim_repr[1]);
    }
    std::vector<double> hess(2 * maxit, 0.0);
    MKL_INT m = 0; double norm = 0.0;
    M.measure_repr_dynamic(Szq, 0, 1, maxit, m, norm, hess.data());
    js.num("dyn_norm", norm); js.integer("dyn_steps", m);
    js.arr("dyn_b", hess.data(), m); js.arr("dyn_a", hess.data() + maxit, m);
}

/* The dynamic part of examples/trans_absent/latt_square/square_Fermi_Hubbard.cc (:147-186) in the FULL basis: E0 and phi0 by
   locate_E0_lanczos, then S^z_q phi0 (moprXvec_full, src/model.cc:1468-1538) with
   S^z_q = sum_r 0.5/sqrt(N) exp(i q.r) (n_up,r - n_dn,r), and measure_full_dynamic's dnmcs coefficients (:1697-1712). */
static void flow_hubbard_full_szq(int Lx, int Ly, double nup, double ndn, double t, double U, int qm, int qn, MKL_INT maxit,
                                  const std::string &vec_prefix, Json &js)
{
    Built b = build_hubbard(Lx, Ly, nup, ndn, t, U);
    Model &M = *b.model;
    M.locate_E0_lanczos(qbasis::which_sym::full, 1, 1);
    js.num("E0", M.eigenvals_full[0]);
    js.integer("dim0", M.dim_full[0]);
    qbasis::lattice latt("square", {static_cast<uint32_t>(Lx), static_cast<uint32_t>(Ly)}, {"pbc", "pbc"});
    auto cu = std::vector<std::vector<cplx>>(4, std::vector<cplx>(4, 0.0));
    auto cd = cu;
    cu[0][1] = 1.0; cu[2][3] = 1.0; cd[0][2] = 1.0; cd[1][3] = -1.0;
    Mopr Szq;
    const double PI = 3.1415926535897932;
    for (int x = 0; x < Lx; x++) for (int y = 0; y < Ly; y++) {
        const double qdotr = 2.0 * PI * (qm * x / static_cast<double>(Lx) + qn * y / static_cast<double>(Ly));
        auto coeff = 0.5 / sqrt(static_cast<double>(Lx * Ly)) * std::exp(cplx{0.0, qdotr});
        uint32_t si; std::vector<int> work(latt.dim);
        latt.coor2site({x, y}, 0, si, work);
        Opr cui(si, 0, true, cu), cdi(si, 0, true, cd);
        auto cuid = cui; cuid.dagger(); auto cdid = cdi; cdid.dagger();
        Szq += coeff * (cuid * cui - cdid * cdi);
    }
    if (!vec_prefix.empty()) {
        std::vector<cplx> vnew(M.dim_full[0]);
        M.moprXvec_full(Szq, 0, 0, static_cast<MKL_INT>(0), vnew.data());
        dump_vec(vec_prefix + "_phi0.bin", M.eigenvecs_full.data(), sizeof(cplx) * M.dim_full[0]);
        dump_vec(vec_prefix + "_Aphi0.bin", vnew.data(), sizeof(cplx) * M.dim_full[0]);
    }
    std::vector<double> hess(2 * maxit, 0.0);
    MKL_INT m = 0; double norm = 0.0;
    M.measure_full_dynamic(Szq, 0, 0, maxit, m, norm, hess.data());
    js.num("dyn_norm", norm); js.integer("dyn_steps", m);
    js.arr("dyn_b", hess.data(), m); js.arr("dyn_a", hess.data() + maxit, m);
}

/* ------------------------------------------------------------------------------------------------ hubbard_direct
 * The reference's csr_mat<complex<double>> of the square-lattice Hubbard model filled DIRECTLY -- ia/ja/val in the layout of
 * src/sparse.cc:202-233 (upper triangle, every diagonal stored, columns ascending), rows in the Lin order of
 * src/basis.cc:1144-1190 -- without the LIL intermediate, whose forward_list nodes need > 140 GB for BASELINE config 3
 * (SURVEY F6).  Only the ASSEMBLY is restated (same statement as tests/lin_builders.py: hubbard_upper_csr; `--check` compares
 * it bit for bit with what the reference's own generate_Ham_sparse_full assembles, tests/test_oracle.py runs that on 4x3);
 * the matrix then lives in the reference's own class and every action (--time-mv: csr_mat::MultMv, --lanczos, ...) is the
 * reference's code.  This is what lets `bench.py --impl reference` time the reference on config 3 itself. */
static inline int popc32(uint32_t v) { return __builtin_popcount(v); }

static void fill_hubbard_direct(qbasis::csr_mat<cplx> &H, int Lx, int Ly, int nup, int ndn, double t, double U)
{
    const int ns = Lx * Ly, nA = (ns + 1) / 2, nB = ns / 2;
    if (2 * nA > 24) { fprintf(stderr, "hubbard_direct: at most 24 sites\n"); exit(2); }
    /* merged bonds with multiplicity, in the order the example adds them (x outer, y inner; +x then +y) */
    struct Bd { int i, j, w; };
    std::vector<Bd> bonds;
    auto site = [&](int x, int y) { return ((x % Lx + Lx) % Lx) + ((y % Ly + Ly) % Ly) * Lx; };
    auto add = [&](int i, int j) { for (auto &b : bonds) if ((b.i == i && b.j == j) || (b.i == j && b.j == i)) { b.w++; return; } bonds.push_back({i, j, 1}); };
    for (int x = 0; x < Lx; x++) for (int y = 0; y < Ly; y++) { add(site(x, y), site(x + 1, y)); add(site(x, y), site(x, y + 1)); }
    /* Lin tables: a-label = digits (up + 2 dn) of the even sites, b-label = of the odd sites; classes by (n_up, n_dn) */
    const uint32_t sizeA = 1u << (2 * nA), sizeB = 1u << (2 * nB);
    const int nc = nA + 1;
    auto counts = [](uint32_t lab, int &c0, int &c1) { c0 = popc32(lab & 0x55555555u); c1 = popc32(lab & 0xAAAAAAAAu); };
    std::vector<int32_t> csize(nc * nc, 0), rankA(sizeA), class_off(nc * nc, 0);
    for (uint32_t a = 0; a < sizeA; a++) { int c0, c1; counts(a, c0, c1); rankA[a] = csize[c0 * nc + c1]++; }
    { int32_t acc = 0; for (int k = 0; k < nc * nc; k++) { class_off[k] = acc; acc += csize[k]; } }
    std::vector<uint32_t> alist(sizeA);
    for (uint32_t a = 0; a < sizeA; a++) { int c0, c1; counts(a, c0, c1); alist[class_off[c0 * nc + c1] + rankA[a]] = a; }
    std::vector<int64_t> Jb((size_t)sizeB + 1, 0);
    int64_t run = 0;
    for (uint32_t b = 0; b < sizeB; b++) {
        Jb[b] = run;
        int c0, c1; counts(b, c0, c1);
        const int n0 = nup - c0, n1 = ndn - c1;
        if (n0 >= 0 && n0 <= nA && n1 >= 0 && n1 <= nA) run += csize[n0 * nc + n1];
    }
    Jb[sizeB] = run;
    const int64_t n = run;
    auto parity_below = [](uint32_t la, uint32_t lb, int s) {
        const int ka = (s + 1) >> 1, kb = s >> 1;
        const uint32_t ma = ka >= 16 ? 0xFFFFFFFFu : ((1u << (2 * ka)) - 1u), mb = kb >= 16 ? 0xFFFFFFFFu : ((1u << (2 * kb)) - 1u);
        return (popc32(la & ma) + popc32(lb & mb)) & 1;
    };
    /* one row: upper-triangle entries (col >= row), unsorted; returns their number */
    auto row_upper = [&](int64_t r, uint32_t la, uint32_t lb, int64_t *cols, double *vals) {
        int cnt = 0;
        double diag = 0.0;
        const int ndbl = popc32(la & (la >> 1) & 0x55555555u) + popc32(lb & (lb >> 1) & 0x55555555u);
        for (int k = 0; k < ndbl; k++) diag += U;
        cols[cnt] = r; vals[cnt] = diag; cnt++;
        for (const auto &b : bonds) {
            double amp = 0.0;
            for (int k = 0; k < b.w; k++) amp += -t;
            for (int dir = 0; dir < 2; dir++) {
                const int to = dir ? b.j : b.i, from = dir ? b.i : b.j;
                for (int sp = 0; sp < 2; sp++) {
                    const uint32_t bf = 1u << (2 * (from >> 1) + sp), bt = 1u << (2 * (to >> 1) + sp);
                    const uint32_t lf = (from & 1) ? lb : la, lt = (to & 1) ? lb : la;
                    if (!(lf & bf) || (lt & bt)) continue;
                    int sg = parity_below(la, lb, from);
                    if (sp == 1 && (lf & (1u << (2 * (from >> 1))))) sg ^= 1;
                    uint32_t na = la, nb = lb;
                    if (from & 1) nb ^= bf; else na ^= bf;
                    sg ^= parity_below(na, nb, to);
                    const uint32_t lt2 = (to & 1) ? nb : na;
                    if (sp == 1 && (lt2 & (1u << (2 * (to >> 1))))) sg ^= 1;
                    if (to & 1) nb ^= bt; else na ^= bt;
                    const int64_t c = Jb[nb] + rankA[na];
                    if (c > r) { cols[cnt] = c; vals[cnt] = sg ? -amp : amp; cnt++; }
                }
            }
        }
        return cnt;
    };
    auto state_of = [&](int64_t r, uint32_t lb_hint, uint32_t &la) {
        int c0, c1; counts(lb_hint, c0, c1);
        la = alist[class_off[(nup - c0) * nc + (ndn - c1)] + (int32_t)(r - Jb[lb_hint])];
    };
    H.dim = n; H.sym = true;
    H.ia = new MKL_INT[n + 1];
    const int maxrow = 1 + 4 * (int)bonds.size();
    /* pass 1: row lengths (parallel over b-labels: the rows of one label are contiguous) */
    #pragma omp parallel
    {
        std::vector<int64_t> cols(maxrow); std::vector<double> vals(maxrow);
        #pragma omp for schedule(dynamic, 64)
        for (int64_t b = 0; b < (int64_t)sizeB; b++)
            for (int64_t r = Jb[b]; r < Jb[b + 1]; r++) { uint32_t la; state_of(r, (uint32_t)b, la); H.ia[r + 1] = row_upper(r, la, (uint32_t)b, cols.data(), vals.data()); }
    }
    H.ia[0] = 0;
    for (int64_t r = 0; r < n; r++) H.ia[r + 1] += H.ia[r];
    H.nnz = H.ia[n];
    H.ja = new MKL_INT[H.nnz];
    H.val = new cplx[H.nnz];
    /* pass 2: fill, columns ascending inside a row */
    #pragma omp parallel
    {
        std::vector<int64_t> cols(maxrow); std::vector<double> vals(maxrow); std::vector<int> idx(maxrow);
        #pragma omp for schedule(dynamic, 64)
        for (int64_t b = 0; b < (int64_t)sizeB; b++)
            for (int64_t r = Jb[b]; r < Jb[b + 1]; r++) {
                uint32_t la; state_of(r, (uint32_t)b, la);
                const int cnt = row_upper(r, la, (uint32_t)b, cols.data(), vals.data());
                for (int k = 0; k < cnt; k++) idx[k] = k;
                std::sort(idx.begin(), idx.begin() + cnt, [&](int x, int y) { return cols[x] < cols[y]; });
                MKL_INT at = H.ia[r];
                for (int k = 0; k < cnt; k++) { H.ja[at] = cols[idx[k]]; H.val[at] = cplx(vals[idx[k]], 0.0); at++; }
            }
    }
    /* the handle the reference's constructor creates (src/sparse.cc:258) */
    if (mkl_sparse_z_create_csr(&H.handle, SPARSE_INDEX_BASE_ZERO, H.dim, H.dim, H.ia, H.ia + 1, H.ja, H.val) != SPARSE_STATUS_SUCCESS) { fprintf(stderr, "create_csr failed\n"); exit(2); }
}

template <typename T>
static void run_actions(qbasis::csr_mat<T> &H, int argc, char **argv, int argi, Json &js)
{
    const MKL_INT n = H.dim;
    js.integer("dim", n); js.integer("nnz", H.nnz); js.integer("sym", H.sym ? 1 : 0);
    js.integer("is_complex", sizeof(T) == sizeof(cplx));
    double maximag = 0.0;
    if constexpr (sizeof(T) == sizeof(cplx)) for (MKL_INT p = 0; p < H.nnz; p++) maximag = std::max(maximag, std::abs(H.val[p].imag()));
    js.num("max_abs_imag", maximag);
    for (int a = argi; a < argc; a++) {
        std::string opt = argv[a];
        if (opt == "--dump" && a + 1 < argc) {
            dump_csr(H, argv[++a]);
        } else if (opt == "--vec-write" && a + 2 < argc) {
            /* vec_randomize(seed) written by the reference's vec_disk_write (src/miscellaneous.cc:437-468) */
            uint32_t seed = (uint32_t)atoi(argv[++a]);
            std::vector<T> x(n);
            qbasis::vec_randomize(n, x.data(), seed);
            qbasis::vec_disk_write(argv[++a], n, x.data());
        } else if (opt == "--mv" && a + 2 < argc) {
            /* y = H x for x = vec_randomize(seed); dump y (reference csr_mat::MultMv, src/sparse.cc:291-297) */
            uint32_t seed = (uint32_t)atoi(argv[++a]);
            std::vector<T> x(n), y(n);
            qbasis::vec_randomize(n, x.data(), seed);
            H.MultMv(x.data(), y.data());
            dump_vec(argv[++a], y.data(), sizeof(T) * n);
        } else if (opt == "--time-mv" && a + 2 < argc) {
            /* CPU baseline: the reference's csr_mat::MultMv (src/sparse.cc:291-297) timed REPS times after WARM warm-ups */
            int reps = atoi(argv[++a]), warm = atoi(argv[++a]);
            std::vector<T> x(n), y(n);
            qbasis::vec_randomize(n, x.data(), 1);
            for (int w = 0; w < warm; w++) H.MultMv(x.data(), y.data());
            std::vector<double> ts;
            double tall0 = now_s();
            for (int r = 0; r < reps; r++) { double t0 = now_s(); H.MultMv(x.data(), y.data()); ts.push_back(now_s() - t0); }
            double tall = now_s() - tall0;
            std::sort(ts.begin(), ts.end());
            js.num("mv_median_s", ts[ts.size() / 2]); js.num("mv_min_s", ts[0]); js.num("mv_total_s", tall); js.integer("mv_reps", reps);
            js.integer("threads", qb_shim_get_spmv_threads());
        } else if (opt == "--lanczos" && a + 2 < argc) {
            /* reference lanczos(0, maxit-1, maxit, m, dim, H, v, hess, purpose) from vec_randomize(seed=1),
               exactly as model::locate_E0_lanczos starts it (src/model.cc:1158-1186) */
            std::string purpose = argv[++a];
            MKL_INT maxit = atoll(argv[++a]);
            std::vector<double> hess(2 * maxit, 0.0), ritz, s;
            std::vector<T> v(2 * n);
            qbasis::vec_randomize(n, v.data(), 1);
            MKL_INT m = 0;
            double t0 = now_s();
            qbasis::lanczos(static_cast<MKL_INT>(0), maxit - 1, maxit, m, n, H, v.data(), hess.data(), purpose);
            double dt = now_s() - t0;
            qbasis::hess_eigen(hess.data(), maxit, m, "sr", ritz, s);
            js.integer("lanczos_steps", m); js.num("lanczos_E0", ritz[0]); js.num("lanczos_seconds", dt);
            if (m > 1) js.num("lanczos_E1_ritz", ritz[1]);
            js.arr("lanczos_b", hess.data(), m + 1); js.arr("lanczos_a", hess.data() + maxit, m);
        } else if (opt == "--lanczos-ckpt" && a + 3 < argc) {
            /* the reference's lanczos with its own checkpoints enabled (src/ckpt.cc; out_Qckpt/ under the work directory):
               called as lanczos(0, NP, maxit, ...) it either starts from vec_randomize(seed=1) and leaves the checkpoint of
               step NP behind, or -- when out_Qckpt/ already holds a checkpoint, whoever wrote it -- resumes from there */
            std::string purpose = argv[++a];
            MKL_INT maxit = atoll(argv[++a]), np = atoll(argv[++a]);
            std::vector<double> hess(2 * maxit, 0.0), ritz, s;
            std::vector<T> v(3 * n);
            qbasis::vec_randomize(n, v.data(), 1);
            MKL_INT m = 0;
            qbasis::enable_ckpt = true;
            qbasis::lanczos(static_cast<MKL_INT>(0), np, maxit, m, n, H, v.data(), hess.data(), purpose);
            qbasis::enable_ckpt = false;
            qbasis::hess_eigen(hess.data(), maxit, m, "sr", ritz, s);
            js.integer("ckpt_steps", m); js.num("ckpt_E0", ritz[0]);
            js.arr("ckpt_b", hess.data(), m + 1); js.arr("ckpt_a", hess.data() + maxit, m);
        } else if (opt == "--cg" && a + 2 < argc) {
            /* reference eigenvec_CG as called from locate_E0_lanczos (src/model.cc:1209-1218) */
            double E0 = atof(argv[++a]);
            MKL_INT maxit = 1000, m = 0; double accu = 0.0;
            std::vector<T> v(4 * n);
            qbasis::vec_randomize(n, v.data() + 2 * n, 1);
            double t0 = now_s();
            qbasis::eigenvec_CG(n, maxit, m, H, static_cast<T>(E0), accu, v.data() + 2 * n, v.data(), v.data() + n, v.data() + 3 * n);
            js.integer("cg_steps", m); js.num("cg_accuracy", accu); js.num("cg_seconds", now_s() - t0);
            dump_vec(argv[++a], v.data() + 2 * n, sizeof(T) * n);
        } else if (opt == "--cg-ckpt" && a + 3 < argc) {
            /* the reference's eigenvec_CG with its checkpoints enabled (src/ckpt.cc:345-480): runs until step MAXIT (leaving
               CG_{V,R,P}<MAXIT>.dat in out_Qckpt/) or to convergence, starting from vec_randomize(seed=1) or from whatever CG
               checkpoint out_Qckpt/ holds */
            double E0 = atof(argv[++a]);
            MKL_INT maxit = atoll(argv[++a]), m = 0; double accu = 0.0;
            std::vector<T> v(4 * n);
            qbasis::vec_randomize(n, v.data() + 2 * n, 1);
            qbasis::enable_ckpt = true;
            qbasis::eigenvec_CG(n, maxit, m, H, static_cast<T>(E0), accu, v.data() + 2 * n, v.data(), v.data() + n, v.data() + 3 * n);
            qbasis::enable_ckpt = false;
            js.integer("cgck_steps", m); js.num("cgck_accuracy", accu);
            dump_vec(argv[++a], v.data() + 2 * n, sizeof(T) * n);
        } else if (opt == "--kpm" && a + 3 < argc) {
            /* Chebyshev moments mu_k = Re <phi| T_k(Ht) |phi>, Ht = (H - c)/s, on the REFERENCE's own pieces: [lo, hi] from the
               reference's energy_scale (src/kpm.cc:45-88; extend 0.1, ITERS iterations), every product the reference's
               csr_mat::MultMv (src/sparse.cc:291-297), dots accumulated in long double; phi = vec_randomize(SEED).
               The reference itself has no Chebyshev recurrence: this is the plain three-term recurrence
               T_0 = phi, T_1 = Ht phi, T_{k+1} = 2 Ht T_k - T_{k-1}, written here so that the moments the GPU path produces
               are pinned to numbers computed with the reference's product. */
            uint32_t seed = (uint32_t)atoi(argv[++a]);
            MKL_INT nmom = atoll(argv[++a]), iters = atoll(argv[++a]);
            std::vector<T> w(2 * n); double lo, hi;
            qbasis::energy_scale(n, H, w.data(), lo, hi, 0.1, iters);
            const double cc = 0.5 * (hi + lo), ss = 0.5 * (hi - lo);
            std::vector<T> phi(n), t0(n), t1(n), t2(n);
            qbasis::vec_randomize(n, phi.data(), seed);
            std::vector<double> mu(nmom, 0.0);
            auto redot = [&](const std::vector<T> &u, const std::vector<T> &v) {      /* Re <u, v> */
                long double acc = 0.0L;
                for (MKL_INT j = 0; j < n; j++) acc += (long double)std::real(std::conj(cplx(u[j])) * cplx(v[j]));
                return (double)acc;
            };
            t0 = phi;
            mu[0] = redot(phi, t0);
            H.MultMv(t0.data(), t1.data());
            for (MKL_INT j = 0; j < n; j++) t1[j] = (t1[j] - cc * t0[j]) / ss;
            if (nmom > 1) mu[1] = redot(phi, t1);
            for (MKL_INT k = 2; k < nmom; k++) {
                H.MultMv(t1.data(), t2.data());
                for (MKL_INT j = 0; j < n; j++) t2[j] = 2.0 * (t2[j] - cc * t1[j]) / ss - t0[j];
                mu[k] = redot(phi, t2);
                t0.swap(t1); t1.swap(t2);
            }
            js.num("kpm_lo", lo); js.num("kpm_hi", hi); js.arr("kpm_moments", mu.data(), nmom);
        } else if (opt == "--energy-scale" && a + 1 < argc) {
            /* reference energy_scale (src/kpm.cc:45-88); start vector = vec_randomize default seed */
            MKL_INT iters = atoll(argv[++a]);
            std::vector<T> v(2 * n); double lo, hi;
            qbasis::energy_scale(n, H, v.data(), lo, hi, 0.1, iters);
            js.num("escale_lo", lo); js.num("escale_hi", hi);
        } else {
            fprintf(stderr, "unknown option %s\n", opt.c_str()); exit(2);
        }
    }
}

static void usage() {
    fprintf(stderr,
        "usage: qb_ref [--threads T] [--workdir D] --out results.json <case> <args...> [actions]\n"
        " cases: heis_chain L none|sz SZ | heis_chain_k L SZ K | tri Lx Ly SZ | tri_k Lx Ly SZ M N |\n"
        "        hubbard_direct Lx Ly NUP NDN T U [--check]  (csr_mat filled without the LIL intermediate; --check: compare with the reference's assembly) |\n"
        "        hubbard Lx Ly NUP NDN T U | hubbard_k Lx Ly NUP NDN T U M N | tj_chain L N SZ | honeycomb Lx Ly | honeycomb_k Lx Ly M N | file_z F.qbcsr | file_d F.qbcsr |\n"
        "        spin_one_chain L SZ | spin_one_chain_k L SZ K | kondo_chain L NELEC T JK | kagome_heisenberg Lx Ly SZ | kagome_tj Lx Ly N SZ | kagome_tj_k Lx Ly N SZ M N | bose_hubbard Lx Ly N NMAX T U |\n"
        "        heis_chain_szq L SZ K0 Q MAXIT [--dump-vecs PREFIX]  (E0 in sector K0, then S^z_Q phi0 and its dnmcs Lanczos in K0-Q)\n"
        "        heis_chain_smq ...                                    (same with S^-_Q: the target sector has Sz - 1)\n"
        "        hubbard_full_szq Lx Ly NUP NDN T U QM QN MAXIT [--dump-vecs PREFIX]  (full basis: E0, S^z_q phi0, measure_full_dynamic)\n"
        " model actions (before matrix actions): --locate-E0 NEV NCV (reference model::locate_E0_lanczos)\n"
        " matrix actions: --dump F | --vec-write SEED F | --mv SEED F | --time-mv REPS WARM | --lanczos PURPOSE MAXIT | --lanczos-ckpt PURPOSE MAXIT NP | --cg E0 F | --cg-ckpt E0 MAXIT F | --energy-scale ITERS | --kpm SEED NMOM ITERS\n");
    exit(2);
}

int main(int argc, char **argv)
{
    int a = 1, threads = 1;
    std::string out = "", workdir = "/tmp/qb_ref_work";
    while (a < argc && argv[a][0] == '-') {
        std::string o = argv[a];
        if (o == "--threads" && a + 1 < argc) threads = atoi(argv[++a]);
        else if (o == "--out" && a + 1 < argc) out = argv[++a];
        else if (o == "--workdir" && a + 1 < argc) workdir = argv[++a];
        else usage();
        a++;
    }
    if (a >= argc || out.empty()) usage();
    /* absolute-ise paths before chdir: the reference appends log_Lanczos_*.txt / log_CG.txt to cwd */
    auto absolutise = [](std::string p) { if (!p.empty() && p[0] != '/') { char cwd[4096]; if (getcwd(cwd, sizeof cwd)) p = std::string(cwd) + "/" + p; } return p; };
    out = absolutise(out);
    std::vector<std::string> av(argv, argv + argc);
    for (int i = a; i < argc; i++) {
        const std::string &s = av[i];
        if ((s == "--dump" || s == "file_z" || s == "file_d") && i + 1 < argc) av[i + 1] = absolutise(av[i + 1]);
        if ((s == "--mv" || s == "--cg" || s == "--vec-write") && i + 2 < argc) av[i + 2] = absolutise(av[i + 2]);
        if (s == "--cg-ckpt" && i + 3 < argc) av[i + 3] = absolutise(av[i + 3]);
        if (s == "--dump-vecs" && i + 1 < argc) av[i + 1] = absolutise(av[i + 1]);
    }
    std::vector<char *> cargv;
    for (auto &s : av) cargv.push_back(const_cast<char *>(s.c_str()));
    argv = cargv.data();
    std::string cmd = "mkdir -p " + workdir; if (system(cmd.c_str()) != 0) return 2;
    if (chdir(workdir.c_str()) != 0) { perror("chdir"); return 2; }
    omp_set_num_threads(threads);
    qb_shim_set_spmv_threads(threads);
    std::cout << std::setprecision(14);

    Json js;
    std::string c = argv[a++];
    js.str("case", c);
    double t0 = now_s();
    if (c == "heis_chain_szq" || c == "heis_chain_smq") {
        if (a + 5 > argc) usage();
        int L = atoi(argv[a]); double sz = atof(argv[a + 1]); int k0 = atoi(argv[a + 2]), q = atoi(argv[a + 3]); MKL_INT maxit = atoll(argv[a + 4]); a += 5;
        std::string prefix;
        if (a + 1 < argc && std::string(argv[a]) == "--dump-vecs") { prefix = argv[a + 1]; a += 2; }
        flow_heis_chain_szq(L, sz, k0, q, maxit, prefix, js, c == "heis_chain_smq" ? 'm' : 'z');
    } else if (c == "hubbard_full_szq") {
        if (a + 9 > argc) usage();
        int Lx = atoi(argv[a]), Ly = atoi(argv[a + 1]); double nu = atof(argv[a + 2]), nd = atof(argv[a + 3]), t = atof(argv[a + 4]), U = atof(argv[a + 5]);
        int qm = atoi(argv[a + 6]), qn = atoi(argv[a + 7]); MKL_INT maxit = atoll(argv[a + 8]); a += 9;
        std::string prefix;
        if (a + 1 < argc && std::string(argv[a]) == "--dump-vecs") { prefix = argv[a + 1]; a += 2; }
        flow_hubbard_full_szq(Lx, Ly, nu, nd, t, U, qm, qn, maxit, prefix, js);
    } else if (c == "hubbard_direct") {
        if (a + 6 > argc) usage();
        int Lx = atoi(argv[a]), Ly = atoi(argv[a + 1]), nu = atoi(argv[a + 2]), nd = atoi(argv[a + 3]); double t = atof(argv[a + 4]), U = atof(argv[a + 5]); a += 6;
        qbasis::csr_mat<cplx> H;
        fill_hubbard_direct(H, Lx, Ly, nu, nd, t, U);
        js.num("build_seconds", now_s() - t0);
        if (a < argc && std::string(argv[a]) == "--check") {       /* bit for bit against the reference's own assembly */
            a++;
            Built b = build_hubbard(Lx, Ly, nu, nd, t, U);
            auto &R = b.model->HamMat_csr_full[0];
            bool same = R.dim == H.dim && R.nnz == H.nnz && R.sym == H.sym;
            /* ia, ja exactly; values by ==, i.e. identical up to the sign of a zero imaginary part (the reference's
               operator products leave -0.0 there on entries with a negative amplitude; no product can tell the difference) */
            if (same) same = memcmp(R.ia, H.ia, sizeof(MKL_INT) * (H.dim + 1)) == 0 && memcmp(R.ja, H.ja, sizeof(MKL_INT) * H.nnz) == 0;
            if (same) for (MKL_INT k = 0; k < H.nnz && same; k++) same = R.val[k] == H.val[k];
            js.integer("direct_equals_reference_assembly", same ? 1 : 0);
            if (!same && R.dim == H.dim && R.nnz == H.nnz) {             /* say where (diagnostics) */
                long long d_ia = 0, d_ja = 0, d_val = 0, d_val_num = 0;
                for (MKL_INT i = 0; i <= H.dim; i++) d_ia += R.ia[i] != H.ia[i];
                for (MKL_INT k = 0; k < H.nnz; k++) { d_ja += R.ja[k] != H.ja[k]; d_val += memcmp(&R.val[k], &H.val[k], sizeof(cplx)) != 0; d_val_num += R.val[k] != H.val[k]; }
                js.integer("direct_diff_ia", d_ia); js.integer("direct_diff_ja", d_ja); js.integer("direct_diff_val_bits", d_val); js.integer("direct_diff_val_numeric", d_val_num);
                js.integer("direct_sym_ref", R.sym ? 1 : 0);
            }
        }
        run_actions(H, argc, argv, a, js);
    } else if (c == "file_z" || c == "file_d") {
        if (a >= argc) usage();
        std::string path = argv[a++];
        if (c == "file_z") { qbasis::csr_mat<cplx> H; load_csr(H, path); run_actions(H, argc, argv, a, js); }
        else               { qbasis::csr_mat<double> H; load_csr(H, path); run_actions(H, argc, argv, a, js); }
    } else {
        Built b;
        auto need = [&](int k) { if (a + k > argc) usage(); };
        if (c == "heis_chain") { need(2); int L = atoi(argv[a++]); std::string mode = argv[a++]; double sz = 0; if (mode == "sz") { need(1); sz = atof(argv[a++]); } b = build_heis_chain(L, mode == "sz", sz, -1); }
        else if (c == "heis_chain_k") { need(3); int L = atoi(argv[a++]); double sz = atof(argv[a++]); int k = atoi(argv[a++]); b = build_heis_chain(L, true, sz, k); }
        else if (c == "tri") { need(3); int Lx = atoi(argv[a++]), Ly = atoi(argv[a++]); double sz = atof(argv[a++]); b = build_triangular(Lx, Ly, sz, -1, -1); }
        else if (c == "tri_k") { need(5); int Lx = atoi(argv[a++]), Ly = atoi(argv[a++]); double sz = atof(argv[a++]); int m = atoi(argv[a++]), n = atoi(argv[a++]); b = build_triangular(Lx, Ly, sz, m, n); }
        else if (c == "hubbard") { need(6); int Lx = atoi(argv[a++]), Ly = atoi(argv[a++]); double nu = atof(argv[a++]), nd = atof(argv[a++]), t = atof(argv[a++]), U = atof(argv[a++]); b = build_hubbard(Lx, Ly, nu, nd, t, U); }
        else if (c == "hubbard_k") { need(8); int Lx = atoi(argv[a++]), Ly = atoi(argv[a++]); double nu = atof(argv[a++]), nd = atof(argv[a++]), t = atof(argv[a++]), U = atof(argv[a++]); int km = atoi(argv[a++]), kn = atoi(argv[a++]); b = build_hubbard(Lx, Ly, nu, nd, t, U, km, kn); }
        else if (c == "tj_chain") { need(3); int L = atoi(argv[a++]); double N = atof(argv[a++]), sz = atof(argv[a++]); b = build_tj_chain(L, N, sz); }
        else if (c == "honeycomb") { need(2); int Lx = atoi(argv[a++]), Ly = atoi(argv[a++]); b = build_honeycomb(Lx, Ly); }
        else if (c == "honeycomb_k") { need(4); int Lx = atoi(argv[a++]), Ly = atoi(argv[a++]); int m = atoi(argv[a++]), n = atoi(argv[a++]); b = build_honeycomb(Lx, Ly, m, n); }
        else if (c == "spin_one_chain") { need(2); int L = atoi(argv[a++]); double sz = atof(argv[a++]); b = build_spin_one_chain(L, sz); }
        else if (c == "spin_one_chain_k") { need(3); int L = atoi(argv[a++]); double sz = atof(argv[a++]); int k = atoi(argv[a++]); b = build_spin_one_chain(L, sz, k); }
        else if (c == "kagome_tj_k") { need(6); int Lx = atoi(argv[a++]), Ly = atoi(argv[a++]); double N = atof(argv[a++]), sz = atof(argv[a++]); int m = atoi(argv[a++]), n = atoi(argv[a++]); b = build_kagome_tj(Lx, Ly, N, sz, m, n); }
        else if (c == "kondo_chain") { need(4); int L = atoi(argv[a++]); double N = atof(argv[a++]), t = atof(argv[a++]), JK = atof(argv[a++]); b = build_kondo_chain(L, N, t, JK); }
        else if (c == "kagome_heisenberg") { need(3); int Lx = atoi(argv[a++]), Ly = atoi(argv[a++]); double sz = atof(argv[a++]); b = build_kagome_heisenberg(Lx, Ly, sz); }
        else if (c == "kagome_tj") { need(4); int Lx = atoi(argv[a++]), Ly = atoi(argv[a++]); double N = atof(argv[a++]), sz = atof(argv[a++]); b = build_kagome_tj(Lx, Ly, N, sz); }
        else if (c == "bose_hubbard") { need(6); int Lx = atoi(argv[a++]), Ly = atoi(argv[a++]); double N = atof(argv[a++]); int nmax = atoi(argv[a++]); double t = atof(argv[a++]), U = atof(argv[a++]); b = build_bose_hubbard(Lx, Ly, N, nmax, t, U); }
        else usage();
        js.num("build_seconds", now_s() - t0);
        Model &M = *b.model;
        while (a < argc && std::string(argv[a]) == "--locate-E0") {
            if (a + 2 >= argc) usage();
            MKL_INT nev = atoll(argv[a + 1]), ncv = atoll(argv[a + 2]); a += 3;
            double t1 = now_s();
            M.locate_E0_lanczos(b.sym, nev, ncv);
            js.num("locate_seconds", now_s() - t1);
            auto &ev = b.sym == qbasis::which_sym::full ? M.eigenvals_full : M.eigenvals_repr;
            js.arr("locate_eigenvals", ev.data(), (long long)ev.size());
        }
        auto &H = b.sym == qbasis::which_sym::full ? M.HamMat_csr_full[0] : M.HamMat_csr_repr[0];
        run_actions(H, argc, argv, a, js);
    }
    std::ofstream fo(out); fo << js.done() << std::endl; fo.close();
    std::cout << std::endl << "QBREF done -> " << out << std::endl;
    return 0;
}
