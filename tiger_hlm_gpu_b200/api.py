"""ctypes binding of libhlm_b200.so plus a mirror of the reference's operator names.

Reference surface mirrored here (paths relative to the reference's src/):
  Model trait (UID, N_EQ, Parameters) ........ models/model_204.hpp:15-30
  rk45_api::setModelParameters<T>(p) ......... model_registry.hpp:9-13
  rk45_api::run_rk45<T>(h_y0,t0,tf,tq,d_sp) .. solver/rk45_api.hpp:273-313
  SpatialParams (136-byte record) ............ I_O/parameters_loader.hpp:19-37
Everything numeric happens in the CUDA library; this file only marshals numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, astuple

import numpy as np

ABI_VERSION = 4
_HERE = os.path.dirname(os.path.abspath(__file__))


class HlmError(RuntimeError):
    """Raised for any non-zero status of the C ABI (the reference throws std::runtime_error)."""


def lib_path() -> str:
    # HLM_B200_LIB selects another build of the same library (kernel-tuning experiments)
    return os.environ.get("HLM_B200_LIB") or os.path.join(_HERE, "libhlm_b200.so")


_lib = None

_V, _LL, _D, _I = C.c_void_p, C.c_longlong, C.c_double, C.c_int

# name -> (restype, argtypes); also the list of symbols the CPU test-suite checks against the header
SIGNATURES = {
    "hlm_create": (_I, [_I, C.POINTER(_V)]),
    "hlm_destroy": (None, [_V]),
    "hlm_last_error": (C.c_char_p, []),
    "hlm_abi_version": (_I, []),
    "hlm_set_stream": (_I, [_V, _V]),
    "hlm_synchronize": (_I, [_V]),
    "hlm_model_info": (_I, [_I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "hlm_set_model_parameters": (_I, [_V, _I, _V]),
    "hlm_get_model_parameters": (_I, [_V, _I, _V]),
    "hlm_upload_spatial_params": (_I, [_V, _V, _LL, _LL]),
    "hlm_upload_forcing": (_I, [_V, _I, _D, _LL, _LL, _V]),
    "hlm_upload_forcing_chunk": (_I, [_V, _I, _D, _LL, _LL, _LL, _LL, _V]),
    "hlm_set_forcing_columns": (_I, [_V, _V, _LL]),
    "hlm_clear_forcings": (_I, [_V]),
    "hlm_set_max_attempts": (_I, [_V, _LL]),
    "hlm_set_dense_window_bytes": (_I, [_V, _LL]),
    "hlm_set_reject_limit": (_I, [_V, _I]),
    "hlm_set_precision": (_I, [_V, _I]),
    "hlm_set_output_states": (_I, [_V, C.c_uint]),
    "hlm_set_output_precision": (_I, [_V, _I]),
    "hlm_output_layout": (_I, [_V, _I, C.POINTER(_I), C.POINTER(_I)]),
    "hlm_set_schedule": (_I, [_V, _I]),
    "hlm_set_stiff_fallback": (_I, [_V, _I]),
    "hlm_solve_radau_steps": (_I, [_V, _V]),
    "hlm_run_rk45": (_I, [_V, _I, _V, _LL, _D, _D, _V, _LL, _V, _V, _V, _V, _V, _V]),
    "hlm_solve_begin": (_I, [_V, _I, _V, _LL, _D, _D, _V, _LL]),
    "hlm_solve_restart": (_I, [_V, _D, _D, _V, _LL]),
    "hlm_solve_advance": (_I, [_V, _D, _V, _LL]),
    "hlm_solve_window": (_I, [_V, _LL, _I]),
    "hlm_solve_window_buffer": (_I, [_V, C.POINTER(_V), C.POINTER(_LL), C.POINTER(_LL)]),
    "hlm_solve_fetch_window": (_I, [_V, _V]),
    "hlm_solve_fetch_window_packed": (_I, [_V, _V, C.POINTER(_I)]),
    "hlm_solve_wait_copy": (_I, [_V, _I]),
    "hlm_host_alloc": (_I, [C.POINTER(_V), _LL]),
    "hlm_host_free": (_I, [_V]),
    "hlm_solve_totals": (_I, [_V, _V]),
    "hlm_solve_end": (_I, [_V, _V, _V, _V, _V, _V]),
    "hlm_solve_peek": (_I, [_V, _V, _V, _V]),
    "hlm_route_set_topology": (_I, [_V, _V, _V, _LL, _V, _LL]),
    "hlm_route_clear": (_I, [_V]),
    "hlm_route_set_send_buffer": (_I, [_V, _V]),
    "hlm_route_send_buffer": (_I, [_V, C.POINTER(_V), C.POINTER(_LL)]),
    "hlm_route_pack": (_I, [_V]),
    "hlm_route_gather": (_I, [_V, _V]),
    "hlm_route_peek": (_I, [_V, _V, _V]),
    "hlm_route_peer_alloc": (_I, [_V, _I, _I, _LL, _V]),
    "hlm_route_peer_open": (_I, [_V, _V]),
    "hlm_route_peer_close": (_I, [_V]),
    "hlm_launch_count": (_LL, [_V]),
    "hlm_kernel_time_ms": (_I, [_V, C.POINTER(_D), C.POINTER(_LL)]),
    "hlm_measure_fma_peak": (_I, [_V, _I, C.POINTER(_D)]),
    "hlm_debug_eval": (_I, [_V, _I, _V, _V, _V, _LL]),
}


def load_library():
    """dlopen libhlm_b200.so; fails loudly — there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise HlmError(f"{path} is missing: build it with `make -C tiger_hlm_gpu_b200/csrc` "
                       "(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.hlm_abi_version() != ABI_VERSION:
        raise HlmError(f"ABI mismatch: library {lib.hlm_abi_version()} vs binding {ABI_VERSION}")
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        raise HlmError(f"hlm status {rc}: {load_library().hlm_last_error().decode()}")


#: numpy mirror of SpatialParams, I_O/parameters_loader.hpp:19-37
SPATIAL_PARAMS_DTYPE = np.dtype(
    [("stream", "<i8"), ("next_stream", "<i8")]
    + [(n, "<f8") for n in ("c1", "infil", "perco", "Hu", "lat", "sw", "ss", "n_mann", "slope",
                            "L", "A_h", "alpha3", "alpha4", "melt_f", "temp_thr")])
assert SPATIAL_PARAMS_DTYPE.itemsize == 136


@dataclass
class Parameters:
    """Model::Parameters with the reference's defaults (models/model_204.hpp:22-30)."""
    initialStep: float = 0.01
    rtol: float = 1e-6
    atol: float = 1e-9
    safety: float = 0.9
    minScale: float = 0.2
    maxScale: float = 10.0

    def as_array(self):
        return np.array(astuple(self), dtype=np.float64)


class Model204:
    UID = 204
    N_EQ = 5
    SP_TYPE = SPATIAL_PARAMS_DTYPE
    Parameters = Parameters


class Model200:
    """Hillslope-link runoff (project-defined; the reference names it, README.md:95, without defining it)."""
    UID = 200
    N_EQ = 5
    SP_TYPE = SPATIAL_PARAMS_DTYPE
    Parameters = Parameters


class DummyModel:
    UID = 0
    N_EQ = 5
    SP_TYPE = None
    Parameters = Parameters


def model_info(uid: int):
    n_eq, n_sp, n_forc = _I(), _I(), _I()
    _check(load_library().hlm_model_info(uid, C.byref(n_eq), C.byref(n_sp), C.byref(n_forc)))
    return n_eq.value, n_sp.value, n_forc.value


def _p(a):
    return None if a is None else a.ctypes.data_as(_V)


class Solver:
    """One device context (hlm_ctx).  The reference's process-global state, made an object."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        h = _V()
        _check(self._lib.hlm_create(device, C.byref(h)))
        self._h = h
        self.device = device
        self._keep = []  # host arrays that async copies may still read

    def close(self):
        if getattr(self, "_h", None):
            self._lib.hlm_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- configuration -------------------------------------------------------------------------
    def set_stream(self, cuda_stream: int | None):
        _check(self._lib.hlm_set_stream(self._h, _V(cuda_stream) if cuda_stream else None))

    def synchronize(self):
        _check(self._lib.hlm_synchronize(self._h))
        self._keep.clear()

    def set_model_parameters(self, uid: int, p: Parameters):
        arr = p.as_array()
        _check(self._lib.hlm_set_model_parameters(self._h, uid, _p(arr)))

    def get_model_parameters(self, uid: int) -> Parameters:
        arr = np.zeros(6)
        _check(self._lib.hlm_get_model_parameters(self._h, uid, _p(arr)))
        return Parameters(*arr.tolist())

    def upload_spatial_params(self, sp: np.ndarray):
        sp = np.ascontiguousarray(sp, dtype=SPATIAL_PARAMS_DTYPE)
        _check(self._lib.hlm_upload_spatial_params(self._h, _p(sp), sp.shape[0], sp.dtype.itemsize))
        self._keep.append(sp)

    def upload_forcing(self, j: int, dt_hours: float, data: np.ndarray):
        data = np.ascontiguousarray(data, dtype=np.float32)
        assert data.ndim == 2, "forcing is [nT][ncols]"
        _check(self._lib.hlm_upload_forcing(self._h, j, dt_hours, data.shape[0], data.shape[1], _p(data)))
        self._keep.append(data)

    def upload_forcing_chunk(self, j: int, dt_hours: float, nT_total: int, i0: int, data: np.ndarray):
        """Samples [i0, i0 + len(data)) of an nT_total-sample record (time-chunked residency)."""
        data = np.ascontiguousarray(data, dtype=np.float32)
        assert data.ndim == 2, "forcing chunk is [nT_chunk][ncols]"
        _check(self._lib.hlm_upload_forcing_chunk(self._h, j, dt_hours, nT_total, i0, data.shape[0], data.shape[1], _p(data)))
        self._keep.append(data)

    def set_forcing_columns(self, col):
        if col is None:
            _check(self._lib.hlm_set_forcing_columns(self._h, None, 0))
            return
        col = np.ascontiguousarray(col, dtype=np.int32)
        _check(self._lib.hlm_set_forcing_columns(self._h, _p(col), col.shape[0]))
        self._keep.append(col)

    def clear_forcings(self):
        _check(self._lib.hlm_clear_forcings(self._h))

    def set_max_attempts(self, n: int):
        _check(self._lib.hlm_set_max_attempts(self._h, n))

    def set_reject_limit(self, n: int = 5):
        """More than n consecutive rejections flag a link stiff (5 = the reference's rule)."""
        _check(self._lib.hlm_set_reject_limit(self._h, n))

    def set_dense_window_bytes(self, n: int):
        _check(self._lib.hlm_set_dense_window_bytes(self._h, n))

    def set_stiff_fallback(self, enable: bool):
        """Radau IIA re-integration of links the RK45 path flags stiff (run_rk45's documented behaviour)."""
        _check(self._lib.hlm_set_stiff_fallback(self._h, 1 if enable else 0))

    def solve_radau_steps(self) -> np.ndarray:
        out = np.zeros(self._session[2], np.int64)
        _check(self._lib.hlm_solve_radau_steps(self._h, _p(out)))
        return out

    def set_schedule(self, mode: str = "auto"):
        """'tiles', 'lanes', 'sorted' (tiles of links with equal attempt counts; Model 200) or 'auto' — hlm_set_schedule."""
        _check(self._lib.hlm_set_schedule(self._h, {"auto": 0, "tiles": 1, "lanes": 2, "sorted": 3}[mode]))

    def set_precision(self, bits: int):
        _check(self._lib.hlm_set_precision(self._h, bits))

    def set_output_states(self, states=None):
        """config.yaml's output.states applied on the device: dense records carry only these state indices
        (ascending).  None / empty = all states (the reference's layout)."""
        mask = 0
        for i in (states or []):
            mask |= 1 << int(i)
        _check(self._lib.hlm_set_output_states(self._h, mask))

    def set_output_precision(self, bits: int):
        """Dense records as float64 (default, the reference's type) or float32."""
        _check(self._lib.hlm_set_output_precision(self._h, bits))

    def output_layout(self, uid: int):
        """(columns per dense record, numpy dtype of a value) under the current output settings."""
        n, b = _I(), _I()
        _check(self._lib.hlm_output_layout(self._h, uid, C.byref(n), C.byref(b)))
        return n.value, (np.float32 if b.value == 4 else np.float64)

    # -- the operator ----------------------------------------------------------------------------
    def run_rk45(self, uid: int, y0, t0: float, tf: float, tq, want_dense=True, out_final=None, out_dense=None):
        """hlm_run_rk45 with host buffers.  Returns dict(final, dense, stiff, n_accept, n_reject, n_jump)."""
        n_eq = model_info(uid)[0]
        y0 = np.ascontiguousarray(y0, dtype=np.float64).reshape(-1, n_eq)
        ns = y0.shape[0]
        tq = np.ascontiguousarray(tq if tq is not None else [], dtype=np.float64)
        nq = tq.shape[0]
        final = out_final if out_final is not None else np.zeros((ns, n_eq))
        dense = None
        if want_dense and nq > 0:
            ncol, dt = self.output_layout(uid)
            dense = out_dense if out_dense is not None else np.zeros((ns, nq, ncol), dt)
            assert dense.dtype == dt and dense.size == ns * nq * ncol and dense.flags.c_contiguous
        stiff = np.zeros(ns, np.int32)
        na, nr, nj = (np.zeros(ns, np.int64) for _ in range(3))
        _check(self._lib.hlm_run_rk45(self._h, uid, _p(y0), ns, t0, tf, _p(tq), nq, _p(final), _p(dense),
                                      _p(stiff), _p(na), _p(nr), _p(nj)))
        return dict(final=final, dense=dense, stiff=stiff, n_accept=na, n_reject=nr, n_jump=nj)

    # -- resident session --------------------------------------------------------------------------
    def solve_begin(self, uid: int, y0, t0: float, tf: float, tq):
        n_eq = model_info(uid)[0]
        y0 = np.ascontiguousarray(y0, dtype=np.float64).reshape(-1, n_eq)
        tq = np.ascontiguousarray(tq if tq is not None else [], dtype=np.float64)
        self._session = (uid, n_eq, y0.shape[0], tq.shape[0])
        _check(self._lib.hlm_solve_begin(self._h, uid, _p(y0), y0.shape[0], t0, tf, _p(tq), tq.shape[0]))
        self._keep += [y0, tq]

    def solve_restart(self, t0: float, tf: float, tq):
        tq = np.ascontiguousarray(tq if tq is not None else [], dtype=np.float64)
        uid, n_eq, ns, _ = self._session
        self._session = (uid, n_eq, ns, tq.shape[0])
        _check(self._lib.hlm_solve_restart(self._h, t0, tf, _p(tq), tq.shape[0]))
        self._keep.append(tq)

    def solve_advance(self, tf: float, tq):
        """Continue to a later tf keeping every link's time and step size (hlm_solve_advance)."""
        tq = np.ascontiguousarray(tq if tq is not None else [], dtype=np.float64)
        uid, n_eq, ns, _ = self._session
        self._session = (uid, n_eq, ns, tq.shape[0])
        _check(self._lib.hlm_solve_advance(self._h, tf, _p(tq), tq.shape[0]))
        self._keep.append(tq)

    # -- routed runs ---------------------------------------------------------------------------------
    def route_set_topology(self, up_ptr, up_idx, send_idx=None):
        up_ptr = np.ascontiguousarray(up_ptr, dtype=np.int64)
        up_idx = np.ascontiguousarray(up_idx, dtype=np.int32)
        send_idx = np.ascontiguousarray(send_idx if send_idx is not None else [], dtype=np.int32)
        _check(self._lib.hlm_route_set_topology(self._h, _p(up_ptr), _p(up_idx) if up_idx.size else None,
                                                up_ptr.shape[0] - 1, _p(send_idx) if send_idx.size else None,
                                                send_idx.shape[0]))
        self._route = (up_ptr.shape[0] - 1, send_idx.shape[0])

    def route_clear(self):
        _check(self._lib.hlm_route_clear(self._h))

    def route_set_send_buffer(self, dev_ptr: int | None):
        _check(self._lib.hlm_route_set_send_buffer(self._h, _V(dev_ptr) if dev_ptr else None))

    def route_pack(self):
        _check(self._lib.hlm_route_pack(self._h))

    def route_gather(self, dev_halo: int | None = None):
        _check(self._lib.hlm_route_gather(self._h, _V(dev_halo) if dev_halo else None))

    def route_peer_alloc(self, world: int, rank: int, max_send: int) -> bytes:
        """Allocate this rank's halo vector for the peer-memory exchange; returns its 64-byte CUDA IPC handle."""
        buf = C.create_string_buffer(64)
        _check(self._lib.hlm_route_peer_alloc(self._h, world, rank, max_send, C.cast(buf, _V)))
        return bytes(buf.raw)

    def route_peer_open(self, handles: bytes):
        """Map every rank's halo vector: `handles` = the ranks' 64-byte IPC handles, concatenated in rank order."""
        buf = C.create_string_buffer(handles, len(handles))
        _check(self._lib.hlm_route_peer_open(self._h, C.cast(buf, _V)))

    def route_peer_close(self):
        _check(self._lib.hlm_route_peer_close(self._h))

    def route_peek(self):
        ns, n_send = self._route
        qin, send = np.zeros(ns), np.zeros(n_send)
        _check(self._lib.hlm_route_peek(self._h, _p(qin), _p(send) if n_send else None))
        return qin, send

    def solve_window(self, q_hi: int, want_dense: bool = True):
        _check(self._lib.hlm_solve_window(self._h, q_hi, 1 if want_dense else 0))

    def solve_window_buffer(self):
        ptr, lo, hi = _V(), _LL(), _LL()
        _check(self._lib.hlm_solve_window_buffer(self._h, C.byref(ptr), C.byref(lo), C.byref(hi)))
        return ptr.value, lo.value, hi.value

    def solve_fetch_window(self, host_dense: np.ndarray):
        assert host_dense.dtype == self.output_layout(self._session[0])[1] and host_dense.flags.c_contiguous
        _check(self._lib.hlm_solve_fetch_window(self._h, _p(host_dense)))

    def solve_fetch_window_packed(self, host_win: np.ndarray) -> int:
        """Last window packed as [ns][q_hi - q_lo][columns]; returns the ticket for solve_wait_copy."""
        assert host_win.dtype == self.output_layout(self._session[0])[1] and host_win.flags.c_contiguous
        t = _I()
        _check(self._lib.hlm_solve_fetch_window_packed(self._h, _p(host_win), C.byref(t)))
        return t.value

    def solve_wait_copy(self, ticket: int = -1):
        _check(self._lib.hlm_solve_wait_copy(self._h, ticket))

    def solve_fetch_window_ptr(self, host_ptr: int):
        _check(self._lib.hlm_solve_fetch_window(self._h, _V(host_ptr)))

    def solve_totals(self):
        t = np.zeros(7, np.int64)
        _check(self._lib.hlm_solve_totals(self._h, _p(t)))
        return dict(zip(("n_accept", "n_reject", "n_jump", "active", "done", "stiff", "stalled"), t.tolist()))

    def solve_peek(self):
        uid, n_eq, ns, nq = self._session
        t, h, y = np.zeros(ns), np.zeros(ns), np.zeros((n_eq, ns))
        _check(self._lib.hlm_solve_peek(self._h, _p(t), _p(h), _p(y)))
        return t, h, y.T.copy()

    def solve_end(self):
        uid, n_eq, ns, nq = self._session
        final = np.zeros((ns, n_eq))
        stiff = np.zeros(ns, np.int32)
        na, nr, nj = (np.zeros(ns, np.int64) for _ in range(3))
        _check(self._lib.hlm_solve_end(self._h, _p(final), _p(stiff), _p(na), _p(nr), _p(nj)))
        self._keep.clear()
        return dict(final=final, stiff=stiff, n_accept=na, n_reject=nr, n_jump=nj)

    # -- measurement ---------------------------------------------------------------------------------
    def launch_count(self) -> int:
        return self._lib.hlm_launch_count(self._h)

    def kernel_time_ms(self):
        s, n = _D(), _LL()
        _check(self._lib.hlm_kernel_time_ms(self._h, C.byref(s), C.byref(n)))
        return s.value, n.value

    def measure_fma_peak(self, bits: int = 64) -> float:
        v = _D()
        _check(self._lib.hlm_measure_fma_peak(self._h, bits, C.byref(v)))
        return v.value


    def debug_eval(self, op: int, x, y=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(np.broadcast_to(np.asarray(0.0 if y is None else y, dtype=np.float64), x.shape))
        out = np.zeros_like(x)
        _check(self._lib.hlm_debug_eval(self._h, op, _p(x), _p(y), _p(out), x.size))
        return out


# ---- reference-named free functions (process-global default solver, like the reference's globals) ----
_default: Solver | None = None


def _default_solver() -> Solver:
    global _default
    if _default is None:
        _default = Solver(0)
    return _default


def setModelParameters(model, p: Parameters, solver: Solver | None = None):
    """rk45_api::setModelParameters<Model>(p), model_registry.hpp:9-13."""
    (solver or _default_solver()).set_model_parameters(model.UID, p)


def run_rk45(model, h_y0, t0, tf, h_query_times, d_sp=None, solver: Solver | None = None):
    """rk45_api::run_rk45<Model>(h_y0, t0, tf, h_query_times, d_sp) -> (final, dense).

    solver/rk45_api.hpp:273-313.  `d_sp` is the SpatialParams array (host numpy here; the reference
    passes a device pointer it uploaded itself).  final is [ns*N_EQ], dense [ns*nq*N_EQ] in
    [sys][q][comp] order, flat like the reference's std::vector<double> pair.
    """
    s = solver or _default_solver()
    if d_sp is not None:
        s.upload_spatial_params(d_sp)
    r = s.run_rk45(model.UID, h_y0, t0, tf, h_query_times)
    dense = r["dense"] if r["dense"] is not None else np.zeros(0)
    return r["final"].ravel(), dense.ravel()
