// examples/chain_heisenberg_momentum.cc -- the reference's examples/trans_symmetric/latt_chain/chain_Heisenberg_spin_half.cc
// on the GPU through the C++ adaptor: every momentum sector of the L-site spin-1/2 Heisenberg chain at Sz = 0, assembled in
// HBM in the reference's representative convention, E0 by the fused device Lanczos with the reference's argument list.
//
//   g++ -std=c++17 -O2 -I include examples/chain_heisenberg_momentum.cc -L quantum_basis_b200 -lqbgpu \
//       -Wl,-rpath,$PWD/quantum_basis_b200 -o chain_heisenberg_momentum && ./chain_heisenberg_momentum 16
//
// For L = 16 the reference's own asserts (chain_Heisenberg_spin_half.cc:102-117) are checked.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "qbgpu_csr_mat.hpp"

using cplx = std::complex<double>;

int main(int argc, char **argv)
{
    const int L = argc > 1 ? std::atoi(argv[1]) : 16;
    std::vector<int32_t> bonds;
    for (int x = 0; x < L; x++) { bonds.push_back(x); bonds.push_back((x + 1) % L); }
    const double golden16[16] = {-7.142296361, -6.523407057, -5.990986863, -5.615175598, -5.451965668, -5.525353087, -5.823231143, -6.298652725,
                                 -6.872106678, -6.298652725, -5.823231143, -5.525353087, -5.451965668, -5.615175598, -5.990986863, -6.523407057};
    int bad = 0;
    try {
        for (int k = 0; k < L; k++) {
            qbgpu::sector sec({L}, L / 2, {k});                              // enumerate_basis_repr({k}, {Sz_total}, {0})
            auto H = sec.heisenberg(bonds, 1.0);                             // generate_Ham_sparse_repr()
            const int64_t n = H.dimension(), maxit = 1000;
            std::vector<cplx> v(2 * n);
            void *d = nullptr;                                               // vec_randomize(dim, v, seed = 1), generated on the device
            qbgpu::check(qbgpu_malloc(&d, sizeof(cplx) * n), "qbgpu_malloc");
            qbgpu::check(qbgpu_vec_randomize_z(n, d, 1), "qbgpu_vec_randomize_z");
            qbgpu::check(qbgpu_memcpy_d2h(v.data(), d, sizeof(cplx) * n), "qbgpu_memcpy_d2h");
            qbgpu::check(qbgpu_free(d), "qbgpu_free");
            std::vector<double> hess(2 * maxit, 0.0);
            int64_t m = 0;
            H.lanczos(0, maxit - 1, maxit, m, v.data(), hess.data(), "sr_val0");          // lanczos(0, maxit-1, maxit, m, dim, H, v, hess, "sr_val0")
            std::vector<double> ritz(m);
            qbgpu::check(qbgpu_hess_eigen(hess.data(), maxit, m, ritz.data(), nullptr), "hess_eigen");
            std::printf("k = %2d  dim = %8lld  zero-norm = %5lld  steps = %3lld  E0 = %.9f\n", k, (long long)n, (long long)sec.info.zero_norm, (long long)m, ritz[0]);
            if (L == 16 && std::abs(ritz[0] - golden16[k]) > 1e-8) bad++;
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 2;
    }
    return bad ? 1 : 0;
}
