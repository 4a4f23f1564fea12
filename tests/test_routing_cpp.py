"""The C++ routing planner (tiger_hlm_gpu_b200/host/hlm_routing.hpp, what a C++ host links) against the Python
mirror (tiger_hlm_gpu_b200/routing.py): the same plan, array by array."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tiger_hlm_gpu_b200 import routing, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "tiger_hlm_gpu_b200", "host")


@pytest.fixture(scope="module")
def io():
    subprocess.check_call(["make", "-C", HOST, "build/libhlm_hostio.so"], stdout=subprocess.DEVNULL)
    lib = C.CDLL(os.path.join(HOST, "build", "libhlm_hostio.so"))
    lib.hlmio_last_error.restype = C.c_char_p
    lib.hlmio_route_plan.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_longlong] + [C.c_void_p] * 7
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def cpp_plan(io, stream, nxt, world, sub):
    n = len(stream)
    stream, nxt = np.ascontiguousarray(stream, np.int64), np.ascontiguousarray(nxt, np.int64)
    order, rank_lo = np.zeros(n, np.int64), np.zeros(world + 1, np.int64)
    up_ptr, up_idx, send_idx = np.zeros(n + world, np.int64), np.zeros(n, np.int32), np.zeros(n, np.int32)
    counts, meta = np.zeros(world, np.int64), np.zeros(3, np.int64)
    rc = io.hlmio_route_plan(_p(stream), _p(nxt), n, world, sub, _p(order), _p(rank_lo), _p(up_ptr), _p(up_idx), _p(send_idx),
                             _p(counts), _p(meta))
    assert rc == 0, io.hlmio_last_error().decode()
    return order, rank_lo, up_ptr, up_idx, send_idx, counts, meta


@pytest.mark.parametrize("ns,sub,world", [(5000, 250, 1), (5000, 250, 3), (20000, 512, 8), (777, 10, 4)])
def test_cpp_plan_equals_python_plan(io, ns, sub, world):
    sp = synthetic.apply_network(synthetic.make_spatial_params(ns), synthetic.make_network(ns, subbasin_links=sub, seed=ns))
    p = routing.plan(sp["stream"], sp["next_stream"], world, subbasin_links=sub)
    order, rank_lo, up_ptr, up_idx, send_idx, counts, meta = cpp_plan(io, sp["stream"], sp["next_stream"], world, sub)
    assert np.array_equal(order, p.order)
    assert meta.tolist() == [p.max_send, p.n_subbasins, p.n_cut_edges]
    a = b = c = 0
    for r, t in enumerate(p.ranks):
        assert (rank_lo[r], rank_lo[r + 1]) == (t.lo, t.hi)
        assert np.array_equal(up_ptr[a:a + t.n_local + 1], t.up_ptr)
        a += t.n_local + 1
        assert np.array_equal(up_idx[b:b + t.up_idx.size], t.up_idx)
        b += t.up_idx.size
        assert counts[r] == t.send_idx.size and np.array_equal(send_idx[c:c + t.send_idx.size], t.send_idx)
        c += t.send_idx.size


def test_cpp_plan_rejects_cycles_and_duplicate_ids(io):
    stream = np.array([1, 2, 3], np.int64)
    z = [np.zeros(8, np.int64) for _ in range(4)]
    i32 = [np.zeros(8, np.int32) for _ in range(2)]
    assert io.hlmio_route_plan(_p(stream), _p(np.array([2, 3, 1], np.int64)), 3, 1, 10, _p(z[0]), _p(z[1]), _p(z[2]), _p(i32[0]),
                               _p(i32[1]), _p(z[3]), _p(np.zeros(3, np.int64))) != 0
    assert b"cycle" in io.hlmio_last_error()
    assert io.hlmio_route_plan(_p(np.array([1, 1, 3], np.int64)), _p(np.array([3, 3, 0], np.int64)), 3, 1, 10, _p(z[0]), _p(z[1]),
                               _p(z[2]), _p(i32[0]), _p(i32[1]), _p(z[3]), _p(np.zeros(3, np.int64))) != 0
    assert b"duplicate" in io.hlmio_last_error()
    with pytest.raises(ValueError, match="cycle"):
        routing.plan(stream, np.array([2, 3, 1]), 1)
