"""Checkpoint / restart of the Lanczos loop in the reference's on-disk format (src/ckpt.cc).

The reference writes, after every Lanczos step of a "sr_val*" run with checkpoints enabled, into ``out_Qckpt/``:
    HessenbergA.dat   a[0..m-1]        HessenbergB.dat   b[0..m]          (vec_disk_write format: int64 n, raw, CRC-32)
    lanczosV<m-1>.dat, lanczosV<m>.dat the two live vectors               lanczosY0.dat  phi0 (every purpose but "sr_val0")
    lczs_mlns.dat     int32 cnt_accuE0, double accuracy, theta0_prev, theta1_prev   (the stop rule's memory)
behind a two-phase commit: ``lczs_updt.Qckpt1`` (the step being written) -> ``*.new`` files -> ``lczs_updt.Qckpt2`` (new
data complete) -> old files removed, new ones renamed -> both markers removed (ckpt_lanczos_update, src/ckpt.cc:177-292).
On start-up an interrupted update is rolled forward (Qckpt2 present) or rewound by one step (ckpt_lanczos_init, :13-167).

This module reads and writes exactly those files, so a run of the device loop can be resumed by the reference on a CPU and
vice versa, and drives `qbgpu_lanczos_resume_*` in pieces of `every` steps (a device-resident loop should not copy two
vectors to disk per step: SURVEY section 5).  Purposes: "sr_val0", "sr_val1"; for "dnmcs" the reference saves nothing
(its update has no branch for it), and neither does this.  The conjugate-gradient loop's checkpoints (CG_{V,R,P}<m>.dat,
src/ckpt.cc:345-520) are handled the same way around `qbgpu_eigenvec_cg_*` entered with m > 0.  Host-side only: no device
code here.
"""
import os
import struct

import numpy as np

from ._lib import QbgpuError
from . import csr as _csr

DIRNAME = "out_Qckpt"
_NEWABLE = ("HessenbergA.dat", "HessenbergB.dat", "lanczosY0.dat", "lanczosY1.dat", "lczs_mlns.dat")


def _p(d, name):
    return os.path.join(d, name)


def _rm(path):
    try:
        os.remove(path)
    except FileNotFoundError:
        pass


def _read_marker(path):
    with open(path, "rb") as f:
        return struct.unpack("<q", f.read(8))[0]


def _vec_len(path):
    """Element count in the header of a vec_disk_write file (0 when the file is missing or too short)."""
    try:
        with open(path, "rb") as f:
            b = f.read(8)
        return struct.unpack("<q", b)[0] if len(b) == 8 else 0
    except FileNotFoundError:
        return 0


def _write_marker(path, m):
    with open(path, "wb") as f:
        f.write(struct.pack("<q", int(m)))


class _Interrupted(Exception):
    """Raised by the test hook of lanczos_store to model a crash in the middle of an update."""


def lanczos_store(dirpath, m, maxit, dim, v, hessenberg, purpose, stop_state, _crash_after=None):
    """ckpt_lanczos_update (src/ckpt.cc:177-292) for the "val" purposes.  v: host array holding the reference's columns
    (v[(m%2)*dim:] = v_m, v[((m-1)%2)*dim:] = v_{m-1}, phi0 at 2*dim for "sr_val1"); hessenberg[2*maxit];
    stop_state = (cnt_accuE0, accuracy, theta0_prev, theta1_prev).  (_crash_after: "new_files" | "commit" stops there, for
    the recovery tests.)"""
    if "val" not in purpose:
        return                                                            # the reference has no branch for "dnmcs"
    os.makedirs(dirpath, exist_ok=True)
    hess = np.asarray(hessenberg, dtype=np.float64)
    v = np.asarray(v)
    _rm(_p(dirpath, "lczs_updt.Qckpt1"))
    _rm(_p(dirpath, "lczs_updt.Qckpt2"))
    _write_marker(_p(dirpath, "lczs_updt.Qckpt1"), m)
    _csr.vec_disk_write(_p(dirpath, "HessenbergA.dat.new"), hess[maxit:maxit + m])
    _csr.vec_disk_write(_p(dirpath, "HessenbergB.dat.new"), hess[:m + 1])
    if m > 0 and not os.path.exists(_p(dirpath, f"lanczosV{m - 1}.dat")):
        _csr.vec_disk_write(_p(dirpath, f"lanczosV{m - 1}.dat"), v[((m - 1) % 2) * dim:((m - 1) % 2 + 1) * dim])
    _csr.vec_disk_write(_p(dirpath, f"lanczosV{m}.dat"), v[(m % 2) * dim:(m % 2 + 1) * dim])
    if "val0" not in purpose:
        _csr.vec_disk_write(_p(dirpath, "lanczosY0.dat.new"), v[2 * dim:3 * dim])
    with open(_p(dirpath, "lczs_mlns.dat.new"), "wb") as f:
        f.write(struct.pack("<i", int(stop_state[0])))
        f.write(struct.pack("<ddd", float(stop_state[1]), float(stop_state[2]), float(stop_state[3])))
    if _crash_after == "new_files":
        raise _Interrupted()
    _write_marker(_p(dirpath, "lczs_updt.Qckpt2"), m)                     # from here on the new data are the valid ones
    if _crash_after == "commit":
        raise _Interrupted()
    _rm(_p(dirpath, "HessenbergA.dat"))
    _rm(_p(dirpath, "HessenbergB.dat"))
    for name in os.listdir(dirpath):                                      # vectors older than v_{m-1}
        if name.startswith("lanczosV") and name.endswith(".dat") and name[8:-4].isdigit() and int(name[8:-4]) < m - 1:
            _rm(_p(dirpath, name))
    _rm(_p(dirpath, "lanczosY0.dat"))
    _rm(_p(dirpath, "lanczosY1.dat"))
    _rm(_p(dirpath, "lczs_mlns.dat"))
    os.replace(_p(dirpath, "HessenbergA.dat.new"), _p(dirpath, "HessenbergA.dat"))
    os.replace(_p(dirpath, "HessenbergB.dat.new"), _p(dirpath, "HessenbergB.dat"))
    if "val0" not in purpose:
        os.replace(_p(dirpath, "lanczosY0.dat.new"), _p(dirpath, "lanczosY0.dat"))
    os.replace(_p(dirpath, "lczs_mlns.dat.new"), _p(dirpath, "lczs_mlns.dat"))
    _rm(_p(dirpath, "lczs_updt.Qckpt1"))
    _rm(_p(dirpath, "lczs_updt.Qckpt2"))


def _recover(dirpath, maxit):
    """The start-up logic of ckpt_lanczos_init (src/ckpt.cc:37-106): finish or undo an interrupted update; returns k."""
    q1, q2 = _p(dirpath, "lczs_updt.Qckpt1"), _p(dirpath, "lczs_updt.Qckpt2")
    if os.path.exists(q1) and os.path.getsize(q1) == 8:
        k = _read_marker(q1)
        if os.path.exists(q2):                                            # new data complete: roll forward
            if not os.path.exists(_p(dirpath, f"lanczosV{k}.dat")):
                raise QbgpuError(f"checkpoint: lanczosV{k}.dat is missing although the update of step {k} was committed")
            for name in _NEWABLE:
                if os.path.exists(_p(dirpath, name + ".new")):
                    _rm(_p(dirpath, name))
                    os.replace(_p(dirpath, name + ".new"), _p(dirpath, name))
            for kk in range(0, k - 1):
                _rm(_p(dirpath, f"lanczosV{kk}.dat"))
            _rm(q1)
            _rm(q2)
        else:                                                             # rewind
            # The reference checkpoints every step, so it rewinds ONE step (k - 1).  Written every N steps, the last
            # committed step is N steps back: it is the one the old HessenbergB.dat describes (b[0..k] = k + 1 numbers).
            k -= 1
            if not (os.path.exists(_p(dirpath, f"lanczosV{k}.dat")) and (k == 0 or os.path.exists(_p(dirpath, f"lanczosV{k - 1}.dat")))
                    and _vec_len(_p(dirpath, "HessenbergB.dat")) == k + 1):
                k = max(_vec_len(_p(dirpath, "HessenbergB.dat")) - 1, 0)
                if k > 0 and not (os.path.exists(_p(dirpath, f"lanczosV{k}.dat")) and os.path.exists(_p(dirpath, f"lanczosV{k - 1}.dat"))):
                    raise QbgpuError(f"checkpoint: cannot rewind to step {k}: its vectors are missing")
            for name in _NEWABLE:
                _rm(_p(dirpath, name + ".new"))
            for name in os.listdir(dirpath):
                if name.startswith("lanczosV") and name.endswith(".dat") and name[8:-4].isdigit() and int(name[8:-4]) > k:
                    _rm(_p(dirpath, name))
            _rm(q1)
        return k
    k_bgn = 0
    while k_bgn < maxit and not os.path.exists(_p(dirpath, f"lanczosV{k_bgn}.dat")):
        k_bgn += 1
    if k_bgn == maxit:
        return 0
    k = k_bgn
    while os.path.exists(_p(dirpath, f"lanczosV{k + 1}.dat")):
        k += 1
    return k


def lanczos_load(dirpath, maxit, dim, dtype, purpose):
    """ckpt_lanczos_init (src/ckpt.cc:13-167) for the "val" purposes: returns (k, v, hessenberg, stop_state) -- k = 0 and
    zeroed arrays when there is nothing to resume from.  v has 2 columns ("sr_val0") or 3."""
    dt = np.dtype(dtype)
    ncol = 2 if "val0" in purpose else 3
    v = np.zeros(ncol * dim, dtype=dt)
    hess = np.zeros(2 * maxit)
    state = [0, 0.0, 0.0, 0.0]
    if "val" not in purpose or not os.path.isdir(dirpath):
        return 0, v, hess, state
    k = _recover(dirpath, maxit)
    if k <= 0:
        return 0, v, hess, state

    def need(name, n, dtp):
        a = _csr.vec_disk_read(_p(dirpath, name), n, dtp)
        if a is None:
            raise QbgpuError(f"checkpoint: {name} is missing or damaged (size, length or CRC-32)")
        return a
    v[((k - 1) % 2) * dim:((k - 1) % 2 + 1) * dim] = need(f"lanczosV{k - 1}.dat", dim, dt)
    v[(k % 2) * dim:(k % 2 + 1) * dim] = need(f"lanczosV{k}.dat", dim, dt)
    if "val0" not in purpose:
        v[2 * dim:3 * dim] = need("lanczosY0.dat", dim, dt)
    try:
        with open(_p(dirpath, "lczs_mlns.dat"), "rb") as f:
            b = f.read()
    except OSError as e:
        raise QbgpuError(f"checkpoint: lczs_mlns.dat is missing or unreadable ({e})")
    if len(b) != 28:
        raise QbgpuError("checkpoint: lczs_mlns.dat has the wrong size")
    state = [struct.unpack("<i", b[:4])[0], *struct.unpack("<ddd", b[4:])]
    hess[maxit:maxit + k] = need("HessenbergA.dat", k, np.float64)
    hess[:k + 1] = need("HessenbergB.dat", k + 1, np.float64)
    return k, v, hess, state


def lanczos_clean(dirpath):
    """ckpt_lanczos_clean (src/ckpt.cc:305-345): remove the Lanczos files of a finished run."""
    if not os.path.isdir(dirpath):
        return
    for name in os.listdir(dirpath):
        if name in _NEWABLE or (name.startswith("lanczosV") and name.endswith(".dat")) or name.startswith("lczs_updt.Qckpt") \
                or (name.endswith(".new") and name[:-4] in _NEWABLE):
            _rm(_p(dirpath, name))


# ------------------------------------------------------------------------------------------------------------ CG
def cg_store(dirpath, m, v, r, p):
    """ckpt_CG_update (src/ckpt.cc:430-476): CG_{V,R,P}<m>.dat behind the markers CG_updt.Qckpt1/2; the files of the
    previous step are removed once the new ones are complete."""
    os.makedirs(dirpath, exist_ok=True)
    q1, q2 = _p(dirpath, "CG_updt.Qckpt1"), _p(dirpath, "CG_updt.Qckpt2")
    _rm(q1)
    _rm(q2)
    _write_marker(q1, m)
    for tag, a in (("V", v), ("R", r), ("P", p)):
        _csr.vec_disk_write(_p(dirpath, f"CG_{tag}{m}.dat"), a)
    _write_marker(q2, m)
    for name in os.listdir(dirpath):                                      # every older step (the reference: step m-1)
        if name[:4] in ("CG_V", "CG_R", "CG_P") and name.endswith(".dat") and name[4:-4].isdigit() and int(name[4:-4]) < m:
            _rm(_p(dirpath, name))
    _rm(q1)
    _rm(q2)


def cg_load(dirpath, maxit, dim, dtype):
    """ckpt_CG_init (src/ckpt.cc:345-427): returns (m, v, r, p); m = 0 (and None vectors) when there is nothing to resume."""
    if not os.path.isdir(dirpath):
        return 0, None, None, None
    q1, q2 = _p(dirpath, "CG_updt.Qckpt1"), _p(dirpath, "CG_updt.Qckpt2")
    have = lambda k: all(os.path.exists(_p(dirpath, f"CG_{t}{k}.dat")) for t in "VRP")   # noqa: E731
    steps = sorted({int(nm[4:-4]) for nm in os.listdir(dirpath) if nm[:4] == "CG_V" and nm.endswith(".dat") and nm[4:-4].isdigit()})
    if os.path.exists(q1) and os.path.getsize(q1) == 8:
        m = _read_marker(q1)
        if os.path.exists(q2):                                            # new data complete: drop the older steps
            if not have(m):
                raise QbgpuError(f"checkpoint: the CG files of step {m} are missing although its update was committed")
            for k in steps:
                if k < m:
                    for t in "VRP":
                        _rm(_p(dirpath, f"CG_{t}{k}.dat"))
            _rm(q1)
            _rm(q2)
        else:                                                             # drop the half-written step, fall back to the last complete one
            for t in "VRP":
                _rm(_p(dirpath, f"CG_{t}{m}.dat"))
            _rm(q1)
            older = [k for k in steps if k < m and have(k)]
            m = older[-1] if older else 0
    else:
        _rm(q1)
        _rm(q2)
        m = next((k for k in range(maxit) if os.path.exists(_p(dirpath, f"CG_V{k}.dat"))), 0)
    if m <= 0:
        return 0, None, None, None
    out = []
    for t in "VRP":
        a = _csr.vec_disk_read(_p(dirpath, f"CG_{t}{m}.dat"), dim, dtype)
        if a is None:
            raise QbgpuError(f"checkpoint: CG_{t}{m}.dat is missing or damaged (size, length or CRC-32)")
        out.append(a)
    return (m, *out)


def cg_clean(dirpath):
    """ckpt_CG_clean (src/ckpt.cc:482-520)."""
    if not os.path.isdir(dirpath):
        return
    for name in os.listdir(dirpath):
        if (name[:4] in ("CG_V", "CG_R", "CG_P") and name.endswith(".dat")) or name.startswith("CG_updt.Qckpt"):
            _rm(_p(dirpath, name))


def cg_checkpointed(mat, E0, v, r, p, pp, maxit=1000, every=50, dirpath=DIRNAME, max_chunks=None, return_done=False):
    """eigenvec_CG(dim, maxit, m = 0, ...) with the reference's CG checkpoints every `every` steps; resumes from `dirpath`
    when it holds one.  v, r, p, pp: host arrays of length dim (v = the initial guess unless resuming).  Returns (m, accu)."""
    dim = mat.dim
    m, cv, cr, cp = cg_load(dirpath, maxit, dim, mat.dtype)
    if m > 0:
        v[:], r[:], p[:] = cv, cr, cp
    accu, chunks, done = 0.0, 0, False
    while m < maxit:
        stop = min(maxit, m + every)
        m_new, accu = _csr.eigenvec_CG(dim, stop, m, mat, E0, v, r, p, pp)
        if accu < _csr.lanczos_precision and m_new == stop and stop < maxit:
            # converged exactly at the chunk boundary: the loop left at m == maxit WITHOUT the reference's closing pass
            # (src/lanczos.cc:295-317: check |v|, renormalise and restart if | |v| - 1 | > 2e-12), which the uninterrupted
            # loop would still run -- give it one more step of room so that it runs here too
            m_new, accu = _csr.eigenvec_CG(dim, stop + 1, m_new, mat, E0, v, r, p, pp)
        done = m_new < stop or accu < _csr.lanczos_precision
        if m_new > m:
            cg_store(dirpath, m_new, v, r, p)
        m = m_new
        chunks += 1
        if done or (max_chunks is not None and chunks >= max_chunks):
            break
    if return_done:
        return m, accu, bool(done) if chunks else False
    return m, accu


def stop_state_from_coefficients(hessenberg, maxit, m, precision=2e-12):
    """The stop rule's memory after step m, replayed from (a, b) alone (src/lanczos.cc:228-248): what lczs_mlns.dat would
    hold.  Host-side helper for writing a checkpoint from a run that did not track it (O(m^2) per call)."""
    from scipy.linalg import eigh_tridiagonal
    hess = np.asarray(hessenberg, dtype=np.float64)
    cnt, accuracy, t0, t1 = 0, 0.0, 0.0, 0.0
    for j in range(2, m + 1):
        w, s = eigh_tridiagonal(hess[maxit:maxit + j], hess[1:j])
        if j > 3:
            accuracy = abs(hess[j] * s[j - 1, 0])
            cnt = cnt + 1 if abs((w[0] - t0) / w[0]) < precision else 0
            if cnt > 15 and accuracy < precision:
                break                                                     # the reference stops before updating theta*_prev
        t0, t1 = w[0], w[1]
    return [cnt, accuracy, t0, t1]


def lanczos_checkpointed(mat, v, hessenberg, purpose, maxit, every=50, dirpath=DIRNAME, max_chunks=None, return_done=False):
    """lanczos(0, maxit-1, maxit, ...) with the reference's checkpoints, written every `every` steps: resumes from
    `dirpath` when it holds a checkpoint (its vectors and coefficients replace the caller's), otherwise starts from v[0:dim].
    v: host array with 2 ("sr_val0") or 3 columns; hessenberg[2*maxit].  Returns the final step count m.
    `max_chunks` stops after that many pieces, leaving the checkpoint on disk (an interruption, for tests)."""
    import ctypes as C
    from ._lib import check, lib, QBGPU_HOST
    if purpose not in ("sr_val0", "sr_val1"):
        raise QbgpuError("checkpointed Lanczos: purpose must be sr_val0 or sr_val1")
    dim = mat.dim
    k, v_ck, h_ck, state = lanczos_load(dirpath, maxit, dim, mat.dtype, purpose)
    if k > 0:
        v[:v_ck.size] = v_ck
        hessenberg[:] = h_ck
    st = (C.c_double * 4)(*[float(x) for x in state])
    f = lib().qbgpu_lanczos_resume_z if mat.is_complex else lib().qbgpu_lanczos_resume_d
    m = C.c_int64(k)
    chunks, done, finished = 0, False, False
    while True:
        np_ = min(every, maxit - 1 - k)
        if np_ <= 0:
            finished = True                                            # maxit - 1 steps: the reference's loop ends here too
            break
        check(f(mat.handle, k, np_, maxit, C.byref(m), C.c_void_p(v.ctypes.data), C.c_void_p(hessenberg.ctypes.data),
                purpose.encode(), QBGPU_HOST, st))
        done = m.value < k + np_ or (int(st[0]) > 15 and st[1] < _csr.lanczos_precision)
        if m.value > k:
            lanczos_store(dirpath, m.value, maxit, dim, v, hessenberg, purpose, list(st))
        k = m.value
        chunks += 1
        if done or (max_chunks is not None and chunks >= max_chunks):
            break
    if return_done:
        return k, finished or done
    return k


# ------------------------------------------------------------------------------------------- the outer state machine
# model<T>::locate_E0_lanczos with enable_ckpt (src/model.cc:1124-1316) remembers WHICH of its four stages are done -- E0 (simple
# Lanczos), V0 (CG), E1 (Lanczos orthogonal to phi0), V1 (CG) -- in out_Qckpt/lczs_E0_sym<s>_sec<n>[_K<k...>].Qckpt: four bools,
# MKL_INT nconv, double E0, E1, gap = 36 bytes (ckpt_lczsE0_init / _updt, src/model.cc:2522-2749), updated through <name>1 (new
# content) and <name>2 (new content complete) so that a crash at any point leaves either the old or the new state; the finished
# eigenvectors go to eigenvec{0,1}_sym<s>_sec<n>[_K...].dat in the vec_disk_write format.
_E0_STATE = struct.Struct("<????qddd")


def _e0_names(dirpath, sym, sec, momentum):
    tag = f"_sym{int(sym)}_sec{int(sec)}"
    if int(sym) == 1 and momentum is not None:
        tag += "_K" + "".join(str(int(k)) for k in momentum)
    return (_p(dirpath, "lczs_E0" + tag + ".Qckpt"), _p(dirpath, "eigenvec0" + tag + ".dat"), _p(dirpath, "eigenvec1" + tag + ".dat"))


def e0_state_init(dirpath=DIRNAME, sym=0, sec=0, momentum=None):
    """ckpt_lczsE0_init (src/model.cc:2519-2657) without the vector loads: create the state file when there is none (or a
    truncated one), roll an interrupted update forward when its new content was complete, then read the state.
    Returns dict(E0_done, V0_done, E1_done, V1_done, nconv, E0, E1, gap)."""
    if os.path.exists(dirpath) and not os.path.isdir(dirpath):
        os.remove(dirpath)
    os.makedirs(dirpath, exist_ok=True)
    f0, _, _ = _e0_names(dirpath, sym, sec, momentum)
    f1, f2 = f0 + "1", f0 + "2"
    size = _E0_STATE.size
    fresh = dict(E0_done=False, V0_done=False, E1_done=False, V1_done=False, nconv=0, E0=0.0, E1=0.0, gap=0.0)
    if (not os.path.exists(f0) and not os.path.exists(f1)) or (os.path.exists(f0) and os.path.getsize(f0) != size):
        with open(f0, "wb") as f:
            f.write(_E0_STATE.pack(False, False, False, False, 0, 0.0, 0.0, 0.0))
        return fresh
    updating = os.path.exists(f1) and os.path.getsize(f1) == size
    if updating and os.path.exists(f2):                    # new data finished writing: take it, drop the stage's own checkpoints
        _rm(f0)
        with open(f1, "rb") as a, open(f0, "wb") as b:
            b.write(a.read())
        lanczos_clean(dirpath)
        cg_clean(dirpath)
    _rm(f1)
    _rm(f2)
    with open(f0, "rb") as f:
        e0d, v0d, e1d, v1d, nconv, E0, E1, gap = _E0_STATE.unpack(f.read(size))
    return dict(E0_done=e0d, V0_done=v0d, E1_done=e1d, V1_done=v1d, nconv=nconv, E0=E0, E1=E1, gap=gap)


def e0_state_update(st, dirpath=DIRNAME, sym=0, sec=0, momentum=None, eigenvecs=None, _crash_after=None):
    """ckpt_lczsE0_updt (src/model.cc:2659-2746): the new state into <name>1; eigenvec0 recorded from the CG checkpoint when V0
    has just finished, eigenvec0/1 from `eigenvecs` (two host vectors) when V1 has; <name>2 marks the new content complete;
    then the old state is replaced and the stage's own checkpoints are removed.  `_crash_after`: test hook (1: after <name>1,
    2: after <name>2, 3: after the old state was removed)."""
    f0, ev0, ev1 = _e0_names(dirpath, sym, sec, momentum)
    f1, f2 = f0 + "1", f0 + "2"
    if not os.path.exists(f0):
        raise QbgpuError("e0_state_update: no state file (e0_state_init first)")
    _rm(f1)
    _rm(f2)
    blob = _E0_STATE.pack(bool(st["E0_done"]), bool(st["V0_done"]), bool(st["E1_done"]), bool(st["V1_done"]), int(st["nconv"]),
                          float(st["E0"]), float(st["E1"]), float(st["gap"]))
    with open(f1, "wb") as f:
        f.write(blob)
    if _crash_after == 1:
        raise _Interrupted()
    if st["E0_done"] and st["V0_done"] and not st["E1_done"] and not st["V1_done"]:      # record eigenvec0: the CG loop's last V
        for name in sorted(os.listdir(dirpath)):
            if name.startswith("CG_V") and name.endswith(".dat") and name[4:-4].isdigit():
                _rm(ev0)
                with open(_p(dirpath, name), "rb") as a, open(ev0, "wb") as b:
                    b.write(a.read())
    if st["V1_done"]:
        if eigenvecs is None or len(eigenvecs) != 2:
            raise QbgpuError("e0_state_update: V1_done needs the two eigenvectors")
        if not os.path.exists(ev0):
            _csr.vec_disk_write(ev0, eigenvecs[0])
        _csr.vec_disk_write(ev1, eigenvecs[1])
    with open(f2, "wb") as f:                              # before / after this point a restart uses the old / new data
        f.write(blob)
    if _crash_after == 2:
        raise _Interrupted()
    _rm(f0)
    if _crash_after == 3:
        raise _Interrupted()
    with open(f0, "wb") as f:
        f.write(blob)
    lanczos_clean(dirpath)
    cg_clean(dirpath)
    _rm(f1)
    _rm(f2)


def locate_E0_lanczos_checkpointed(mat, nev=1, ncv=1, maxit=1000, every=50, dirpath=DIRNAME, sym=0, sec=0, momentum=None,
                                   max_chunks=None):
    """model<T>::locate_E0_lanczos with enable_ckpt == true (src/model.cc:1124-1316 + 2519-2746), csr_mat branch: the four stages
    E0 / V0 / E1 / V1 with the reference's state file between them and the reference's Lanczos / CG checkpoints inside them
    (lanczos_checkpointed, cg_checkpointed: pieces of `every` steps on the device).  A run that is interrupted -- `max_chunks`
    pieces in total, for tests, or a real crash -- is continued by calling the function again with the same `dirpath`: finished
    stages are skipped, the running one resumes from its own checkpoint, phi0 comes back from eigenvec0*.dat.
    Returns dict(eigenvals, eigenvecs, nconv, finished, state)."""
    if not (0 < nev <= 2 and nev - 1 <= ncv <= nev):
        raise QbgpuError("need 0 < nev <= 2 and nev-1 <= ncv <= nev")
    n, dt = mat.dim, mat.dtype
    f0, ev0, ev1 = _e0_names(dirpath, sym, sec, momentum)
    st = e0_state_init(dirpath, sym, sec, momentum)
    budget = [max_chunks]

    def take():                                            # pieces still allowed in this call (None: no limit)
        return None if budget[0] is None else max(budget[0], 0)

    def spent(k):
        if budget[0] is not None:
            budget[0] -= k

    out = {"eigenvals": [], "eigenvecs": [], "finished": False}
    rnd = lambda seed: _csr.vec_randomize(n, seed, dtype=dt)                  # noqa: E731
    hess = np.zeros(2 * maxit)
    v = np.zeros((4 if ncv > 0 else 2) * n, dtype=dt)
    v[:n] = rnd(1)                                                             # :1165

    def pieces(total_steps_before, total_steps_after):
        return -(-(total_steps_after - total_steps_before) // every) if total_steps_after > total_steps_before else 1

    if not st["E0_done"]:                                                      # :1173-1200
        if take() == 0:
            return dict(out, state=st)
        k0 = lanczos_load(dirpath, maxit, n, dt, "sr_val0")[0]
        m, done = lanczos_checkpointed(mat, v, hess, "sr_val0", maxit, every, dirpath, take(), return_done=True)
        spent(pieces(k0, m))
        if not done:
            return dict(out, state=st)
        ritz, _ = _csr.hess_eigen(hess, maxit, m)
        st.update(E0_done=True, E0=float(ritz[0]), nconv=0)
        out["lanczos_steps"] = m
        e0_state_update(st, dirpath, sym, sec, momentum)
    out["eigenvals"].append(st["E0"])
    if ncv == 0:
        out["finished"] = True
        return dict(out, nconv=st["nconv"], state=st)

    if st["V0_done"] and not st["V1_done"]:                                   # :2614-2627: phi0 back from the disk
        v[2 * n:3 * n] = _csr.vec_disk_read(ev0, n, dt)
    if not st["V0_done"]:                                                      # :1208-1233
        if take() == 0:
            return dict(out, state=st)
        v[2 * n:3 * n] = rnd(1)
        k0 = cg_load(dirpath, maxit, n, dt)[0]
        m, accu, done = cg_checkpointed(mat, st["E0"], v[2 * n:3 * n], v[0:n], v[n:2 * n], v[3 * n:4 * n], maxit, every, dirpath, take(),
                                        return_done=True)
        spent(pieces(k0, m))
        if not done:
            return dict(out, state=st)
        out["cg_steps"], out["cg_accuracy"] = m, accu
        if not any(x.startswith("CG_V") for x in os.listdir(dirpath)):        # (converged without a stored piece: store the result)
            cg_store(dirpath, max(m, 1), v[2 * n:3 * n], v[0:n], v[n:2 * n])
        st.update(V0_done=True, nconv=1)
        e0_state_update(st, dirpath, sym, sec, momentum)

    if nev == 2 and not st["E1_done"]:                                         # :1236-1268
        if take() == 0:
            return dict(out, state=st)
        k0, v_ck, h_ck, _ = lanczos_load(dirpath, maxit, n, dt, "sr_val1")
        if k0 == 0:
            phi0 = v[2 * n:3 * n]
            w = rnd(1)
            w = w - np.vdot(phi0, w) * phi0
            v[:n] = w / np.linalg.norm(w)
            v[n:2 * n] = 0
        hess[:] = 0.0
        v3 = v[:3 * n]
        m, done = lanczos_checkpointed(mat, v3, hess, "sr_val1", maxit, every, dirpath, take(), return_done=True)
        spent(pieces(k0, m))
        if not done:
            return dict(out, state=st)
        ritz, _ = _csr.hess_eigen(hess, maxit, m)
        st.update(E1_done=True, E1=float(ritz[0]), gap=float(ritz[0]) - st["E0"])
        out["lanczos_steps_E1"] = m
        e0_state_update(st, dirpath, sym, sec, momentum)
    if nev == 2:
        out["eigenvals"].append(st["E1"])
        out["gap"] = st["gap"]

    if ncv == 1:                                                               # :1270-1276
        out["eigenvecs"].append(v[2 * n:3 * n].copy())
        out["finished"] = True
        return dict(out, nconv=st["nconv"], state=st)

    if st["V1_done"]:                                                          # :2628-2652
        out["eigenvecs"] = [_csr.vec_disk_read(ev0, n, dt), _csr.vec_disk_read(ev1, n, dt)]
        out["finished"] = True
        return dict(out, nconv=st["nconv"], state=st)
    if take() == 0:                                                            # :1278-1315
        return dict(out, state=st)
    v5 = np.zeros(5 * n, dtype=dt)
    v5[2 * n:3 * n] = v[2 * n:3 * n]
    v5[3 * n:4 * n] = rnd(8)                                                   # seed + 7
    k0 = cg_load(dirpath, maxit, n, dt)[0]
    m, accu, done = cg_checkpointed(mat, st["E1"], v5[3 * n:4 * n], v5[0:n], v5[n:2 * n], v5[4 * n:5 * n], maxit, every, dirpath, take(),
                                    return_done=True)
    spent(pieces(k0, m))
    if not done and m < maxit:
        return dict(out, state=st)
    out["cg_steps_E1"], out["cg_accuracy_E1"] = m, accu
    v0, v1 = v5[2 * n:3 * n].copy(), v5[3 * n:4 * n].copy()
    if st["gap"] < _csr.lanczos_precision:                                     # orthogonalise a degenerate pair, :1302-1311
        v1 = v1 - np.vdot(v0, v1) * v0
        v1 /= np.linalg.norm(v1)
    st.update(V1_done=True, nconv=2)
    e0_state_update(st, dirpath, sym, sec, momentum, eigenvecs=[v0, v1])
    out["eigenvecs"] = [v0, v1]
    out["finished"] = True
    return dict(out, nconv=2, state=st)
