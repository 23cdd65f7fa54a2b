/* oracle/ab_driver.cc -- TEST INFRASTRUCTURE ONLY.
 *
 * A/B drop-in check: the reference's OWN templates lanczos<T,MAT> and eigenvec_CG<T,MAT> (compiled here from
 * /root/reference/src/lanczos.cc, unmodified, by including that translation unit) instantiated over
 * MAT = qbgpu::csr_mat<complex<double>> (include/qbgpu_csr_mat.hpp), i.e. the reference's Krylov loops on the host
 * with every H*v served by the GPU through the MultMv2 seam -- next to the same templates over the reference's
 * csr_mat (CPU) and next to the fused device loop.  Usage: qb_ab <file.qbcsr> <out.json>
 */
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <unistd.h>
#include "qbgpu_csr_mat.hpp"
#include "lanczos.cc"            /* the reference's translation unit: brings the template definitions */

using cplx = std::complex<double>;
namespace qbasis {
template void lanczos(MKL_INT k, MKL_INT np, const MKL_INT &maxit, MKL_INT &m, const MKL_INT &dim,
                      const qbgpu::csr_mat<cplx> &mat, cplx v[], double hessenberg[], const std::string &purpose);
template void eigenvec_CG(const MKL_INT &dim, const MKL_INT &maxit, MKL_INT &m, const qbgpu::csr_mat<cplx> &mat,
                          const cplx &E0, double &accu, cplx v[], cplx r[], cplx p[], cplx pp[]);
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: qb_ab file.qbcsr out.json\n"); return 2; }
    char cwd[4096]; if (!getcwd(cwd, sizeof cwd)) return 2;
    std::string in = argv[1], out = argv[2];
    if (in[0] != '/') in = std::string(cwd) + "/" + in;
    if (out[0] != '/') out = std::string(cwd) + "/" + out;
    if (system("mkdir -p /tmp/qb_ab_work") != 0 || chdir("/tmp/qb_ab_work") != 0) return 2;
    FILE *f = fopen(in.c_str(), "rb"); if (!f) { perror(in.c_str()); return 2; }
    char magic[8]; int64_t dim, nnz; int32_t sym, isc;
    if (fread(magic, 1, 8, f) != 8 || fread(&dim, 8, 1, f) != 1 || fread(&nnz, 8, 1, f) != 1 || fread(&sym, 4, 1, f) != 1 || fread(&isc, 4, 1, f) != 1 || !isc) return 2;
    qbasis::csr_mat<cplx> H;
    H.dim = dim; H.nnz = nnz; H.sym = sym != 0;
    H.ia = new MKL_INT[dim + 1]; H.ja = new MKL_INT[nnz]; H.val = new cplx[nnz];
    if (fread(H.ia, 8, dim + 1, f) != (size_t)(dim + 1) || fread(H.ja, 8, nnz, f) != (size_t)nnz || fread(H.val, 16, nnz, f) != (size_t)nnz) return 2;
    fclose(f);
    if (mkl_sparse_z_create_csr(&H.handle, SPARSE_INDEX_BASE_ZERO, dim, dim, H.ia, H.ia + 1, H.ja, H.val) != SPARSE_STATUS_SUCCESS) return 2;

    qbgpu::csr_mat<cplx> G(H);                           /* upload where the reference creates its handle */
    const MKL_INT maxit = 1000;
    std::ofstream js(out);
    js << std::setprecision(17) << "{\"dim\": " << dim << ", \"nnz\": " << nnz;
    auto run = [&](const char *tag, auto &&call) {
        std::vector<double> hess(2 * maxit, 0.0), ritz, s;
        std::vector<cplx> v(2 * dim);
        qbasis::vec_randomize(dim, v.data(), 1);
        MKL_INT m = 0;
        double t0 = now_s();
        call(m, v.data(), hess.data());
        double dt = now_s() - t0;
        qbasis::hess_eigen(hess.data(), maxit, m, "sr", ritz, s);
        js << ", \"" << tag << "_steps\": " << m << ", \"" << tag << "_E0\": " << ritz[0] << ", \"" << tag << "_seconds\": " << dt;
        return ritz[0];
    };
    /* A: reference lanczos over the reference csr_mat (CPU) */
    run("ref_cpu", [&](MKL_INT &m, cplx *v, double *h) { qbasis::lanczos(static_cast<MKL_INT>(0), maxit - 1, maxit, m, dim, H, v, h, "sr_val0"); });
    /* B: the SAME reference template over the GPU adaptor: host loop, GPU MultMv2 */
    double E0 = run("ref_loop_gpu_mv", [&](MKL_INT &m, cplx *v, double *h) { qbasis::lanczos(static_cast<MKL_INT>(0), maxit - 1, maxit, m, dim, G, v, h, "sr_val0"); });
    /* C: the fused device loop behind the same argument list */
    run("fused_gpu", [&](MKL_INT &m, cplx *v, double *h) { int64_t mm = 0; G.lanczos(0, maxit - 1, maxit, mm, v, h, "sr_val0"); m = mm; });
    /* reference eigenvec_CG over the GPU adaptor */
    {
        std::vector<cplx> v(4 * dim);
        qbasis::vec_randomize(dim, v.data() + 2 * dim, 1);
        MKL_INT m = 0; double accu = 0.0;
        qbasis::eigenvec_CG(dim, maxit, m, G, cplx(E0), accu, v.data() + 2 * dim, v.data(), v.data() + dim, v.data() + 3 * dim);
        std::vector<cplx> r(dim, 0.0);
        H.MultMv(v.data() + 2 * dim, r.data());
        double res = 0.0; for (MKL_INT i = 0; i < dim; i++) res += std::norm(r[i] - E0 * v[2 * dim + i]);
        js << ", \"cg_ref_loop_gpu_mv_steps\": " << m << ", \"cg_accu\": " << accu << ", \"cg_residual\": " << std::sqrt(res);
    }
    js << "}" << std::endl;
    std::cout << std::endl << "QBAB done" << std::endl;
    return 0;
}
