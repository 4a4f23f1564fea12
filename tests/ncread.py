"""Read every variable of a NetCDF file the host code wrote — classic through SciPy's independent reader, NetCDF-4 /
HDF5 (which SciPy cannot open) through the repo's own reader behind libhlm_hostio.so.  Test helper."""
import ctypes as C
import os
import subprocess

import numpy as np
from scipy.io import netcdf_file

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "tiger_hlm_gpu_b200", "host")
_lib = None


def hostio():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", HOST, "build/libhlm_hostio.so"], stdout=subprocess.DEVNULL)
        lib = C.CDLL(os.path.join(HOST, "build", "libhlm_hostio.so"))
        lib.hlmio_last_error.restype = C.c_char_p
        lib.hlmio_variables.restype = C.c_char_p
        lib.hlmio_read_double.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        lib.hlmio_inquire.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


def is_hdf5(path):
    with open(path, "rb") as f:
        return f.read(8) == b"\x89HDF\r\n\x1a\n"


def read_nc(path):
    path = str(path)
    if not is_hdf5(path):
        with netcdf_file(path, "r", mmap=False) as f:
            return {k: np.array(v[:]) for k, v in f.variables.items()}
    io = hostio()
    out = {}
    for name in io.hlmio_variables(path.encode()).decode().split():
        shape = (C.c_longlong * 8)()
        es = C.c_int()
        rank = io.hlmio_inquire(path.encode(), name.encode(), shape, C.byref(es))
        assert rank >= 0, io.hlmio_last_error().decode()
        dims = [shape[k] for k in range(rank)]
        a = np.zeros(dims, np.float64)
        start = (C.c_longlong * max(rank, 1))(*([0] * rank))
        count = (C.c_longlong * max(rank, 1))(*dims)
        assert io.hlmio_read_double(path.encode(), name.encode(), start, count, rank, a.ctypes.data_as(C.c_void_p)) == 0, \
            io.hlmio_last_error().decode()
        out[name] = a
    return out
