"""ctypes wrappers of the checkers built from the reference's own sources (oracle/_ref/*.so).

Test infrastructure only.  libref_cuda.so = the unchanged reference kernel for sm_100a behind a
1-D-launch harness; libref_host.so = the reference's step/dense/rhs templates compiled for the host.
Both are built in the build container by `make -C oracle ref` and travel to the GPU box prebuilt.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
V, D, I, LL = C.c_void_p, C.c_double, C.c_int, C.c_longlong


def _p(a):
    return None if a is None else a.ctypes.data_as(V)


def have(name):
    return os.path.exists(os.path.join(REF_DIR, name))


_cuda = None


def ref_cuda():
    global _cuda
    if _cuda is None:
        lib = C.CDLL(os.path.join(REF_DIR, "libref_cuda.so"))
        lib.ref_cuda_run204.restype = I
        lib.ref_cuda_run204.argtypes = [I, V, I, V, D, D, V, I, V, V, I, V, V, V, V, V, V, V, V, V]
        _cuda = lib
    return _cuda


def ref_cuda_run204(prm6, y0, t0, tf, tq, sp, forc_blocks, dt_hours, counted=True, want_dense=True):
    """Run the reference kernel.  forc_blocks: list of float32 [nT][ns] (per-link expanded)."""
    y0 = np.ascontiguousarray(y0, np.float64).reshape(-1, 5)
    ns = y0.shape[0]
    tq = np.ascontiguousarray(tq, np.float64)
    nq = tq.shape[0]
    sp = np.ascontiguousarray(sp)
    assert sp.dtype.itemsize == 136 and sp.shape[0] == ns
    forc = (np.concatenate([np.ascontiguousarray(b, np.float32).ravel() for b in forc_blocks])
            if forc_blocks else np.zeros(1, np.float32))
    nT = np.array([b.shape[0] for b in forc_blocks] or [0], np.uint64)
    dt = np.array(list(dt_hours) or [1.0], np.float64)
    prm = np.ascontiguousarray(prm6, np.float64)
    final = np.zeros((ns, 5))
    dense = np.zeros((ns, max(nq, 1), 5)) if want_dense else None  # None: timing runs skip the 40*nq bytes/link round trip
    stiff = np.zeros(ns, np.int32)
    att, eok, jmp = (np.zeros(ns, np.int32) for _ in range(3))
    ms = C.c_float()
    rc = ref_cuda().ref_cuda_run204(1 if counted else 0, _p(prm), ns, _p(y0), t0, tf, _p(tq), nq, _p(sp), _p(forc),
                                    len(forc_blocks), _p(dt), _p(nT), _p(final), _p(dense), _p(stiff), _p(att),
                                    _p(eok), _p(jmp), C.addressof(ms))
    if rc != 0:
        raise RuntimeError(f"ref_cuda_run204 failed rc={rc}")
    out = dict(final=final, dense=dense[:, :nq] if dense is not None else None, stiff=stiff, kernel_ms=ms.value)
    if counted:
        out.update(n_accept=(eok - jmp).astype(np.int64), n_reject=(att - eok).astype(np.int64),
                   n_jump=jmp.astype(np.int64))
    return out


_host = None


def ref_host():
    global _host
    if _host is None:
        lib = C.CDLL(os.path.join(REF_DIR, "libref_host.so"))
        lib.ref_host_step204.argtypes = [V, I, V, D, D, D, V, I, V, V, V]
        lib.ref_host_dense204.argtypes = [V, V, D, D, V]
        lib.ref_host_rhs204.argtypes = [V, I, V, V, I, V]
        lib.ref_host_run204.argtypes = [V, I, V, D, D, V, I, V, V, I, V, V, V, V, V, V, V, V, I, I]
        _host = lib
    return _host


def ref_host_run204(prm6, y0, t0, tf, tq, sp, forc_blocks, dt_hours, threads=1, want_dense=True):
    from concurrent.futures import ThreadPoolExecutor
    y0 = np.ascontiguousarray(y0, np.float64).reshape(-1, 5)
    ns = y0.shape[0]
    tq = np.ascontiguousarray(tq, np.float64)
    nq = tq.shape[0]
    sp = np.ascontiguousarray(sp)
    forc = (np.concatenate([np.ascontiguousarray(b, np.float32).ravel() for b in forc_blocks])
            if forc_blocks else np.zeros(1, np.float32))
    nT = np.array([b.shape[0] for b in forc_blocks] or [0], np.int64)
    dt = np.array(list(dt_hours) or [1.0], np.float64)
    prm = np.ascontiguousarray(prm6, np.float64)
    final = np.zeros((ns, 5))
    dense = np.zeros((ns, nq, 5)) if want_dense else None
    stiff = np.zeros(ns, np.int32)
    na, nr, nj = (np.zeros(ns, np.int64) for _ in range(3))
    L = ref_host()

    def work(ab):
        L.ref_host_run204(_p(prm), ns, _p(y0), t0, tf, _p(tq), nq, _p(sp), _p(forc), len(forc_blocks), _p(dt), _p(nT),
                          _p(final), _p(dense), _p(stiff), _p(na), _p(nr), _p(nj), int(ab[0]), int(ab[1]))

    threads = max(1, min(threads, ns))
    cuts = np.linspace(0, ns, threads + 1).astype(int)
    if threads == 1:
        work((0, ns))
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(work, zip(cuts[:-1], cuts[1:])))
    return dict(final=final, dense=dense, stiff=stiff, n_accept=na, n_reject=nr, n_jump=nj)
