/* oracle/shim/arpack-ng/arpack.hpp -- TEST INFRASTRUCTURE ONLY.
 * ARPACK-NG 3.9.0 (Fortran) is not in this image and there is no gfortran, so the four entry points the
 * reference calls (src/lanczos.cc:424,432,473,482) are declared with the real C++ interface's argument
 * lists and throw when reached.  Nothing on the Lanczos/CG/H*v oracle path calls them. */
#ifndef QB_ORACLE_SHIM_ARPACK_HPP
#define QB_ORACLE_SHIM_ARPACK_HPP
#include <complex>
#include <stdexcept>
typedef long long a_int;
namespace arpack {
enum class which : int { largest_algebraic, smallest_algebraic, largest_magnitude, smallest_magnitude,
                         largest_real, smallest_real, largest_imaginary, smallest_imaginary, both_ends };
enum class bmat : int { identity, generalized };
enum class howmny : int { ritz_vectors, schur_vectors, ritz_specified };
[[noreturn]] inline void unavailable(const char *f) { throw std::runtime_error(std::string("ARPACK not available in the oracle shim: ") + f); }
inline void saupd(a_int &, bmat, a_int, which, a_int, double, double *, a_int, double *, a_int, a_int *, a_int *,
                  double *, double *, a_int, a_int &) { unavailable("saupd"); }
inline void seupd(a_int, howmny, a_int *, double *, double *, a_int, double, bmat, a_int, which, a_int, double,
                  double *, a_int, double *, a_int, a_int *, a_int *, double *, double *, a_int, a_int &) { unavailable("seupd"); }
inline void naupd(a_int &, bmat, a_int, which, a_int, double, std::complex<double> *, a_int, std::complex<double> *,
                  a_int, a_int *, a_int *, std::complex<double> *, std::complex<double> *, a_int, double *, a_int &) { unavailable("naupd"); }
inline void neupd(a_int, howmny, a_int *, std::complex<double> *, std::complex<double> *, a_int, std::complex<double>,
                  std::complex<double> *, bmat, a_int, which, a_int, double, std::complex<double> *, a_int,
                  std::complex<double> *, a_int, a_int *, a_int *, std::complex<double> *, std::complex<double> *,
                  a_int, double *, a_int &) { unavailable("neupd"); }
}
#endif
