// examples/square_fermi_hubbard.cc -- the reference's
// examples/trans_absent/latt_square/square_Fermi_Hubbard.cc on the GPU through the C++ adaptor: the Lx x Ly Fermi-Hubbard
// model (periodic, t = 1, U = 1.1) at fixed N_up, N_dn, generated in HBM in the reference's basis order and sign
// convention, E0 by the fused device Lanczos with the reference's argument list -- through the ordinary handle, through
// the species-order handle (QBGPU_SPECIES_ORDER: two-pass product, same calling convention) and through its matrix-free
// kind (nothing stored).  Defaults 4 2 4 4 reproduce the reference's assert (square_Fermi_Hubbard.cc:112); 4 4 8 8 is
// BASELINE config 3.
//
//   g++ -std=c++17 -O2 -I include examples/square_fermi_hubbard.cc -L quantum_basis_b200 -lqbgpu \
//       -Wl,-rpath,$PWD/quantum_basis_b200 -o square_fermi_hubbard && ./square_fermi_hubbard 4 2 4 4
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "qbgpu_csr_mat.hpp"

using cplx = std::complex<double>;

static double E0_by_lanczos(const qbgpu::csr_mat<cplx> &H, int64_t &steps)
{
    const int64_t n = H.dimension(), maxit = 1000;
    std::vector<cplx> v(2 * n);
    void *d = nullptr;                                                   // vec_randomize(dim, v, seed = 1), generated on the device
    qbgpu::check(qbgpu_malloc(&d, sizeof(cplx) * n), "qbgpu_malloc");
    qbgpu::check(qbgpu_vec_randomize_z(n, d, 1), "qbgpu_vec_randomize_z");
    qbgpu::check(qbgpu_memcpy_d2h(v.data(), d, sizeof(cplx) * n), "qbgpu_memcpy_d2h");
    qbgpu::check(qbgpu_free(d), "qbgpu_free");
    std::vector<double> hess(2 * maxit, 0.0);
    H.lanczos(0, maxit - 1, maxit, steps, v.data(), hess.data(), "sr_val0");   // lanczos(0, maxit-1, maxit, m, dim, H, v, hess, "sr_val0")
    std::vector<double> ritz(steps);
    qbgpu::check(qbgpu_hess_eigen(hess.data(), maxit, steps, ritz.data(), nullptr), "hess_eigen");
    return ritz[0];
}

int main(int argc, char **argv)
{
    const int Lx = argc > 1 ? std::atoi(argv[1]) : 4, Ly = argc > 2 ? std::atoi(argv[2]) : 2;
    const int nup = argc > 3 ? std::atoi(argv[3]) : 4, ndn = argc > 4 ? std::atoi(argv[4]) : 4;
    const double t = 1.0, U = 1.1;
    std::vector<int32_t> bonds;                                          // one +x and one +y bond per site, like the reference's loops (:47-90)
    auto site = [&](int x, int y) { return ((x % Lx) + Lx) % Lx + (((y % Ly) + Ly) % Ly) * Lx; };
    for (int x = 0; x < Lx; x++)
        for (int y = 0; y < Ly; y++) {
            bonds.push_back(site(x, y)); bonds.push_back(site(x + 1, y));
            bonds.push_back(site(x, y)); bonds.push_back(site(x, y + 1));
        }
    const struct { const char *name; int matrix_free, flags; } kinds[3] = {
        {"stored, reference order  ", 0, 0}, {"stored, species order     ", 0, QBGPU_SPECIES_ORDER}, {"matrix-free, species order", 1, QBGPU_SPECIES_ORDER}};
    int bad = 0;
    try {
        double first = 0.0;
        for (int k = 0; k < 3; k++) {
            qbgpu_matrix_t h = nullptr;
            const auto t0 = std::chrono::steady_clock::now();
            if (kinds[k].matrix_free)
                qbgpu::check(qbgpu_create_matfree_hubbard(&h, Lx * Ly, nup, ndn, (int)bonds.size() / 2, bonds.data(), t, U, 1, kinds[k].flags, 0, -1), "qbgpu_create_matfree_hubbard");
            else
                qbgpu::check(qbgpu_build_hubbard(&h, Lx * Ly, nup, ndn, (int)bonds.size() / 2, bonds.data(), t, U, 1, kinds[k].flags, 0, -1), "qbgpu_build_hubbard");
            auto H = qbgpu::csr_mat<cplx>::adopt(h);
            const double t_build = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            int64_t steps = 0;
            const auto t1 = std::chrono::steady_clock::now();
            const double E0 = E0_by_lanczos(H, steps);
            const double t_lan = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
            std::printf("%s  dim = %lld  build %.2f s  Lanczos %lld steps in %.2f s  E0 = %.10f\n", kinds[k].name, (long long)H.dimension(),
                        t_build, (long long)steps, t_lan, E0);
            if (k == 0) first = E0;
            if (std::abs(E0 - first) > 1e-10 * std::abs(first)) bad++;
            if (Lx == 4 && Ly == 2 && nup == 4 && ndn == 4 && std::abs(E0 + 14.07605866) > 1e-8) bad++;   // square_Fermi_Hubbard.cc:112
        }
        // The momentum-resolved part (examples/trans_symmetric/latt_square/square_Fermi_Hubbard.cc): every (m, n) sector built on
        // the device with the reference's representatives, norms and matrix elements, E0 per sector.  Their minimum is the E0
        // above; for 4 x 2 at (4, 4) the list is the reference's E0_list (:112-119, index = Ly m + n).
        if (Lx * Ly <= 16 && Lx % 2 == 0) {                            // (the site numbering above is the sector code's when x is the even direction)
            const double E0_list_4x2[8] = {-14.07605866, -10.50470669, -12.16861094, -12.19847764, -10.54300366, -14.03137587, -12.16861094, -12.19847764};
            std::vector<int32_t> hops;                                   // per bond: c+_up,i c_up,j ; c+_up,j c_up,i ; c+_dn,i c_dn,j ; c+_dn,j c_dn,i
            // the sector code numbers sites with the first even direction fastest (src/lattice.cc:591-615); for Lx even that is x
            for (size_t b = 0; b + 1 < bonds.size(); b += 2)
                for (int sp = 0; sp < 2; sp++) {
                    hops.push_back(bonds[b]); hops.push_back(bonds[b + 1]); hops.push_back(sp);
                    hops.push_back(bonds[b + 1]); hops.push_back(bonds[b]); hops.push_back(sp);
                }
            double lowest = 1e300;
            int64_t live = 0;
            const auto t2 = std::chrono::steady_clock::now();
            for (int m = 0; m < Lx; m++)
                for (int n = 0; n < Ly; n++) {
                    qbgpu::electron_sector sec({Lx, Ly}, nup, ndn, {m, n});
                    live += sec.dim() - sec.info.zero_norm;
                    auto H = sec.hubbard(hops, t, U);
                    int64_t steps = 0;
                    const double E = E0_by_lanczos(H, steps);
                    std::printf("  k = (%d,%d)  dim = %lld (%lld with zero norm)  E0 = %.10f\n", m, n, (long long)sec.dim(), (long long)sec.info.zero_norm, E);
                    if (E < lowest) lowest = E;
                    if (Lx == 4 && Ly == 2 && nup == 4 && ndn == 4 && std::abs(E - E0_list_4x2[Ly * m + n]) > 1e-8) bad++;
                }
            std::printf("momentum sectors: %lld live representatives in all sectors, lowest E0 = %.10f, %.2f s\n", (long long)live, lowest,
                        std::chrono::duration<double>(std::chrono::steady_clock::now() - t2).count());
            if (std::abs(lowest - first) > 1e-8) bad++;
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 2;
    }
    return bad ? 1 : 0;
}
