"""Routed run of Model 200 on the CPU oracle: the same per-interval sequence the GPU driver
(tiger_hlm_gpu_b200.routing.RoutedSolver) queues — gather inflow, integrate the interval with every link's
time and step size carried over — with the exchange left to the caller (none, a simulated one, or gloo).
Test infrastructure only."""
import numpy as np

from oracle import oracle as O
from tiger_hlm_gpu_b200 import routing


class OracleRank:
    def __init__(self, topo, sp, forcing, y0, prm, t0, threads=4, device_pow=True, max_attempts=2_000_000):
        self.topo, self.sp, self.forcing, self.prm = topo, sp, forcing, prm
        self.y = np.ascontiguousarray(y0, dtype=np.float64).copy()
        n = self.y.shape[0]
        self.t = np.full(n, float(t0))
        self.h = np.full(n, prm.initialStep)
        self.t_end = float(t0)
        self.kw = dict(threads=threads, device_pow=device_pow, max_attempts=max_attempts)
        self.na = np.zeros(n, np.int64)
        self.nr = np.zeros(n, np.int64)
        self.qin = np.zeros(n)

    def send(self, max_send):
        out = np.zeros(max_send)
        out[: self.topo.send_idx.size] = self.y[self.topo.send_idx, 0]
        return out

    def gather(self, halo):
        self.qin = routing.gather_inflow_reference(self.topo, self.y[:, 0], halo)

    def advance(self, tf, tq):
        r = O.run_rk45(200, self.prm, self.y, self.t_end, tf, tq, sp=self.sp, forcing=self.forcing,
                       inflow=self.qin, state_io=(self.t, self.h), stiff_fallback=True,
                       reject_limit=routing.ROUTED_REJECT_LIMIT, **self.kw)
        assert np.isin(r["stiff"], (0, 3)).all()     # 3 = flagged by the RK45 loop, finished by the implicit fallback
        self.y = r["final"]
        self.t_end = tf
        self.na += r["n_accept"]
        self.nr += r["n_reject"]
        return r


def run_single(sp, forcing, y0, prm, down_stream_plan, t0, tf, dt_couple, queries_per_interval=1, **kw):
    """world = 1 reference run; returns (final, dense [ns][nq][5], tq, n_accept)."""
    topo = down_stream_plan.ranks[0]
    rk = OracleRank(topo, sp, forcing, y0, prm, t0, **kw)
    edges = np.arange(t0, tf + 0.5 * dt_couple, dt_couple)
    dense, tqs = [], []
    for a, b in zip(edges[:-1], edges[1:]):
        tq = a + (b - a) * np.arange(1, queries_per_interval + 1) / queries_per_interval
        rk.gather(None)
        r = rk.advance(b, tq)
        dense.append(r["dense"])
        tqs.append(tq)
    return rk.y, np.concatenate(dense, axis=1), np.concatenate(tqs), rk.na
