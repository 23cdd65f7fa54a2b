"""ctypes bindings for the CPU oracle (oracle/_build/libqb_oracle.so) and .qbcsr / golden-fixture helpers.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
The product package (quantum_basis_b200/) must never import this module.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "libqb_oracle.so")
QB_REF = os.path.join(ORACLE_DIR, "_ref", "qb_ref")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

_lib = None
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
c128p = np.ctypeslib.ndpointer(dtype=np.complex128, flags="C_CONTIGUOUS")


def build():
    """Compile the plain-C oracle (and nothing else)."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "liboracle"], check=True)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    L.qbo_vec_randomize_d.argtypes = [C.c_int64, f64p, C.c_uint32]
    L.qbo_vec_randomize_z.argtypes = [C.c_int64, c128p, C.c_uint32]
    L.qbo_spmv_d.argtypes = [C.c_int64, i64p, i64p, f64p, C.c_int, f64p, f64p, C.c_int, C.c_int]
    L.qbo_spmv_z.argtypes = [C.c_int64, i64p, i64p, c128p, C.c_int, c128p, c128p, C.c_int, C.c_int]
    L.qbo_spmv_z_ld.argtypes = [C.c_int64, i64p, i64p, c128p, C.c_int, c128p, c128p]
    L.qbo_expand_upper_z.argtypes = [C.c_int64, i64p, i64p, c128p, i64p, C.c_void_p, C.c_void_p]
    L.qbo_expand_upper_z.restype = C.c_int64
    L.qbo_hess_eigen.argtypes = [f64p, C.c_int64, C.c_int64, f64p, C.c_void_p]
    L.qbo_lanczos_d.argtypes = [C.c_int64, i64p, i64p, f64p, C.c_int, f64p, f64p, C.c_int64, C.c_char_p, C.c_int]
    L.qbo_lanczos_d.restype = C.c_int64
    L.qbo_lanczos_z.argtypes = [C.c_int64, i64p, i64p, c128p, C.c_int, c128p, f64p, C.c_int64, C.c_char_p, C.c_int]
    L.qbo_lanczos_z.restype = C.c_int64
    L.qbo_eigenvec_cg_d.argtypes = [C.c_int64, i64p, i64p, f64p, C.c_int, C.c_double, C.c_int64,
                                    C.POINTER(C.c_double), f64p, f64p, f64p, f64p, C.c_int]
    L.qbo_eigenvec_cg_d.restype = C.c_int64
    L.qbo_eigenvec_cg_zp.argtypes = [C.c_int64, i64p, i64p, c128p, C.c_int, C.POINTER(C.c_double), C.c_int64,
                                     C.POINTER(C.c_double), c128p, c128p, c128p, c128p, C.c_int]
    L.qbo_eigenvec_cg_zp.restype = C.c_int64
    L.qbo_energy_scale_z.argtypes = [C.c_int64, i64p, i64p, c128p, C.c_int, c128p, C.POINTER(C.c_double),
                                     C.POINTER(C.c_double), C.c_double, C.c_int64, C.c_int]
    L.qbo_kpm_moments_z.argtypes = [C.c_int64, i64p, i64p, c128p, C.c_int, c128p, C.c_double, C.c_double,
                                    C.c_int64, f64p, C.c_int]
    _lib = L
    return L


class Csr:
    """Host CSR in the reference's on-wire form (csr_mat<T>: dim, nnz, sym, ia, ja, val; qbasis.h:976-1021)."""

    def __init__(self, dim, ia, ja, val, sym):
        self.dim = int(dim)
        self.ia = np.ascontiguousarray(ia, dtype=np.int64)
        self.ja = np.ascontiguousarray(ja, dtype=np.int64)
        self.val = np.ascontiguousarray(val)
        self.sym = bool(sym)
        assert self.val.dtype in (np.float64, np.complex128)

    @property
    def nnz(self):
        return int(self.ia[-1])

    @property
    def is_complex(self):
        return self.val.dtype == np.complex128

    def astype(self, dt):
        return Csr(self.dim, self.ia, self.ja, self.val.astype(dt), self.sym)

    def to_scipy_full(self):
        import scipy.sparse as sp
        A = sp.csr_matrix((self.val, self.ja, self.ia), shape=(self.dim, self.dim))
        if self.sym:
            U = sp.triu(A, k=1)
            A = A + U.conj().T
        return A.tocsr()


def read_qbcsr(path):
    with open(path, "rb") as f:
        magic = f.read(8)
        assert magic[:6] == b"QBCSR1", magic
        dim, nnz = np.frombuffer(f.read(16), dtype=np.int64)
        sym, isc = np.frombuffer(f.read(8), dtype=np.int32)
        ia = np.frombuffer(f.read(8 * (dim + 1)), dtype=np.int64)
        ja = np.frombuffer(f.read(8 * nnz), dtype=np.int64)
        val = np.frombuffer(f.read((16 if isc else 8) * nnz), dtype=np.complex128 if isc else np.float64)
    return Csr(dim, ia.copy(), ja.copy(), val.copy(), sym)


def write_qbcsr(path, A):
    with open(path, "wb") as f:
        f.write(b"QBCSR1\0\0")
        np.array([A.dim, A.nnz], dtype=np.int64).tofile(f)
        np.array([1 if A.sym else 0, 1 if A.is_complex else 0], dtype=np.int32).tofile(f)
        A.ia.tofile(f)
        A.ja.tofile(f)
        A.val.tofile(f)


def load_golden(name):
    """Load tests/golden/<name>.npz -> (Csr, dict of recorded reference outputs)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    A = Csr(int(z["dim"]), z["ia"], z["ja"], z["val"], bool(z["sym"]))
    meta = json.loads(str(z["meta"]))
    extra = {k: z[k] for k in z.files if k not in ("dim", "ia", "ja", "val", "sym", "meta")}
    return A, meta, extra


# ------------------------------------------------------------------ oracle entry points (numpy in/out)
def vec_randomize(n, seed=1, dtype=np.complex128):
    x = np.zeros(n, dtype=dtype)
    (lib().qbo_vec_randomize_z if dtype == np.complex128 else lib().qbo_vec_randomize_d)(n, x, seed)
    return x


def spmv(A, x, y=None, nthreads=1):
    """y = H x (MultMv) when y is None, else y += H x (MultMv2) in place."""
    acc = y is not None
    if y is None:
        y = np.zeros(A.dim, dtype=A.val.dtype)
    x = np.ascontiguousarray(x, dtype=A.val.dtype)
    f = lib().qbo_spmv_z if A.is_complex else lib().qbo_spmv_d
    f(A.dim, A.ia, A.ja, A.val, int(A.sym), x, y, int(acc), nthreads)
    return y


def spmv_ld(A, x):
    y = np.zeros(A.dim, dtype=np.complex128)
    Az = A if A.is_complex else A.astype(np.complex128)
    lib().qbo_spmv_z_ld(A.dim, Az.ia, Az.ja, Az.val, int(A.sym), np.ascontiguousarray(x, dtype=np.complex128), y)
    return y


def expand_upper(A):
    Az = A if A.is_complex else A.astype(np.complex128)
    ia_full = np.zeros(A.dim + 1, dtype=np.int64)
    nnz = lib().qbo_expand_upper_z(A.dim, Az.ia, Az.ja, Az.val, ia_full, None, None)
    ja_full = np.zeros(nnz, dtype=np.int64)
    val_full = np.zeros(nnz, dtype=np.complex128)
    lib().qbo_expand_upper_z(A.dim, Az.ia, Az.ja, Az.val, ia_full, ja_full.ctypes.data, val_full.ctypes.data)
    return ia_full, ja_full, val_full


def hess_eigen(hess, maxit, m, vectors=True):
    ritz = np.zeros(m)
    s = np.zeros(m * m) if vectors else None
    lib().qbo_hess_eigen(np.ascontiguousarray(hess, dtype=np.float64), maxit, m, ritz,
                         s.ctypes.data if vectors else None)
    return ritz, (s.reshape(m, m).T if vectors else None)   # s[:, j] = j-th eigenvector


def lanczos(A, v0, maxit=1000, purpose="sr_val0", phi0=None, nthreads=1):
    """Reference lanczos(0, maxit-1, maxit, ...) -> (m, a[0:m], b[0:m+1], final v buffer)."""
    n = A.dim
    nv = 3 if phi0 is not None else 2
    v = np.zeros(nv * n, dtype=A.val.dtype)
    v[:n] = v0
    if phi0 is not None:
        v[2 * n:] = phi0
    hess = np.zeros(2 * maxit)
    f = lib().qbo_lanczos_z if A.is_complex else lib().qbo_lanczos_d
    m = f(n, A.ia, A.ja, A.val, int(A.sym), v, hess, maxit, purpose.encode(), nthreads)
    assert m >= 0
    return m, hess[maxit:maxit + m].copy(), hess[:m + 1].copy(), v


def eigenvec_cg(A, E0, v0, maxit=1000, nthreads=1):
    n = A.dim
    v = np.array(v0, dtype=A.val.dtype, copy=True)
    r = np.zeros(n, dtype=A.val.dtype); p = np.zeros_like(r); pp = np.zeros_like(r)
    accu = C.c_double(0.0)
    if A.is_complex:
        e = (C.c_double * 2)(float(np.real(E0)), float(np.imag(E0)))
        m = lib().qbo_eigenvec_cg_zp(n, A.ia, A.ja, A.val, int(A.sym), e, maxit, C.byref(accu), v, r, p, pp, nthreads)
    else:
        m = lib().qbo_eigenvec_cg_d(n, A.ia, A.ja, A.val, int(A.sym), float(E0), maxit, C.byref(accu), v, r, p, pp, nthreads)
    return m, accu.value, v


def energy_scale(A, v0, extend=0.1, iters=40, nthreads=1):
    Az = A if A.is_complex else A.astype(np.complex128)
    v = np.zeros(2 * A.dim, dtype=np.complex128); v[:A.dim] = v0
    lo = C.c_double(); hi = C.c_double()
    lib().qbo_energy_scale_z(A.dim, Az.ia, Az.ja, Az.val, int(A.sym), v, C.byref(lo), C.byref(hi), extend, iters, nthreads)
    return lo.value, hi.value


def kpm_moments(A, phi, lo, hi, nmom, nthreads=1):
    Az = A if A.is_complex else A.astype(np.complex128)
    mu = np.zeros(nmom)
    lib().qbo_kpm_moments_z(A.dim, Az.ia, Az.ja, Az.val, int(A.sym), np.ascontiguousarray(phi, dtype=np.complex128),
                            lo, hi, nmom, mu, nthreads)
    return mu


def have_qb_ref():
    return os.path.exists(QB_REF)


def run_qb_ref(args, threads=1, workdir=None, timeout=1800):
    """Run the compiled UNMODIFIED reference (oracle/_ref/qb_ref) and return its JSON result."""
    import tempfile
    workdir = workdir or tempfile.mkdtemp(prefix="qbref_")
    out = os.path.join(workdir, "result.json")
    cmd = [QB_REF, "--threads", str(threads), "--workdir", workdir, "--out", out] + [str(a) for a in args]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, timeout=timeout)
    with open(out) as f:
        return json.load(f)
