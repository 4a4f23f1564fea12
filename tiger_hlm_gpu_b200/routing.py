"""Routed runs: links coupled through upstream channel discharge (BASELINE config 5).

The reference carries the river topology in every parameter record (`stream`, `next_stream`:
I_O/parameters_loader.cpp:74, stream.hpp:31,47) and sketches per-step MPI buffers in data/config.yaml:66-70,
but no equation of it reads another link, so everything here is this project's design; the topology input
(`stream`, `next_stream`) is the only contract taken from the reference.

Scheme.  Time advances in coupling intervals.  Over an interval every link is an independent ODE system
(the RK45 path as it is), its channel fed by the discharge of its upstream links held at the value they had
at the interval's start; between intervals that inflow is gathered again.  The coupling is first order in
the interval length and needs one exchange per interval.

Partition.  Links are cut into sub-basins — connected sub-trees of `next_stream` of about `subbasin_links`
links — and whole sub-basins are dealt to ranks in contiguous runs, so an edge is cut only where a sub-basin
drains into a sub-basin of another rank.  The links at the upstream end of cut edges are the boundary links:
their discharge (one f64 per link per interval) is all that crosses ranks, in one all-gather of a padded
per-rank segment.  Inflow sums run over a link's upstream links in ascending original index, so results are
bit-identical under every partition.

This module is host logic (numpy) plus the per-interval driver; kernels are in csrc/ (route_gather_kernel,
the send-buffer epilogue of rk45_window_kernel).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


# Routed runs flag a link stiff after more than this many consecutive rejections (the reference's 5 is a kink detector
# as much as a stiffness test: hlm_b200.h, hlm_set_reject_limit); the CPU routed run (tests/routed_ref.py) uses the same.
ROUTED_REJECT_LIMIT = 20


def downstream_index(stream: np.ndarray, next_stream: np.ndarray) -> np.ndarray:
    """Index of each link's downstream link, -1 for outlets (next_stream not among `stream`)."""
    stream = np.asarray(stream, dtype=np.int64)
    nxt = np.asarray(next_stream, dtype=np.int64)
    order = np.argsort(stream, kind="stable")
    s_sorted = stream[order]
    if s_sorted.size > 1 and (s_sorted[1:] == s_sorted[:-1]).any():
        raise ValueError("duplicate stream ids")
    pos = np.searchsorted(s_sorted, nxt)
    pos_c = np.minimum(pos, stream.size - 1)
    found = s_sorted[pos_c] == nxt
    down = np.where(found, order[pos_c], -1).astype(np.int64)
    down[down == np.arange(stream.size)] = -1  # a link draining into itself is an outlet
    return down


def _leaf_to_root_frontiers(down: np.ndarray):
    """Kahn's order of a forest, as a list of index arrays (leaves first).  Raises on a cycle."""
    n = down.size
    has = down >= 0
    indeg = np.bincount(down[has], minlength=n)
    frontier = np.flatnonzero(indeg == 0)
    seen = 0
    out = []
    while frontier.size:
        out.append(frontier)
        seen += frontier.size
        par = down[frontier]
        par = par[par >= 0]
        if par.size == 0:
            break
        np.subtract.at(indeg, par, 1)
        cand = np.unique(par)
        frontier = cand[indeg[cand] == 0]
    if seen != n:
        raise ValueError("next_stream has a cycle")
    return out


def subbasins(down: np.ndarray, subbasin_links: int = 4096) -> np.ndarray:
    """Label every link with the index of its sub-basin's outlet link.

    Walking from the leaves, a link closes a sub-basin when the not-yet-closed sub-tree above it (itself
    included) has reached `subbasin_links` links, or when it is an outlet."""
    down = np.asarray(down, dtype=np.int64)
    n = down.size
    fronts = _leaf_to_root_frontiers(down)
    size = np.ones(n, dtype=np.int64)
    cut = np.zeros(n, dtype=bool)
    for f in fronts:
        c = (size[f] >= subbasin_links) | (down[f] < 0)
        cut[f] = c
        keep = ~c
        if keep.any():
            np.add.at(size, down[f[keep]], size[f[keep]])
    root = np.arange(n, dtype=np.int64)
    for f in reversed(fronts):
        open_ = f[~cut[f]]
        root[open_] = root[down[open_]]
    return root


@dataclass
class RankTopology:
    """What one rank hands to hlm_route_set_topology (indices are positions inside the rank's range)."""
    rank: int
    lo: int
    hi: int
    up_ptr: np.ndarray      # int64 [n_local + 1]
    up_idx: np.ndarray      # int32 [nnz]; >= 0 local, < 0 halo slot -(e + 1)
    send_idx: np.ndarray    # int32 [n_send] boundary links, in halo-segment order

    @property
    def n_local(self) -> int:
        return self.hi - self.lo


@dataclass
class RoutePlan:
    world: int
    order: np.ndarray                 # planned position -> original link index
    ranges: list                      # [(lo, hi)] per rank, in planned positions
    ranks: list = field(default_factory=list)
    max_send: int = 0                 # length of every rank's segment of the halo vector
    n_subbasins: int = 0
    n_cut_edges: int = 0

    @property
    def halo_len(self) -> int:
        return self.world * self.max_send

    def inverse_order(self) -> np.ndarray:
        inv = np.empty_like(self.order)
        inv[self.order] = np.arange(self.order.size)
        return inv


def plan(stream, next_stream, world: int, subbasin_links: int = 4096) -> RoutePlan:
    """Partition by sub-basin over `world` ranks and build every rank's local topology."""
    down = downstream_index(stream, next_stream)
    n = down.size
    root = subbasins(down, subbasin_links)
    # sub-basins in the order of their first link; contiguous runs of them per rank, balanced by link count
    roots, first, counts = np.unique(root, return_index=True, return_counts=True)
    by_first = np.argsort(first, kind="stable")
    roots, counts = roots[by_first], counts[by_first]
    if roots.size < world:
        raise ValueError(f"{roots.size} sub-basins cannot be dealt to {world} ranks: lower subbasin_links")
    # contiguous runs: a rank is closed when taking half of the next sub-basin would carry it past its share of
    # the links, or when the sub-basins left are only just enough to give every remaining rank one
    sb_rank = np.zeros(roots.size, dtype=np.int64)
    rank, cum, have = 0, 0, 0
    for i, cnt in enumerate(counts.tolist()):
        if rank < world - 1 and have > 0 and ((2 * cum + cnt) * world > 2 * (rank + 1) * n or roots.size - i <= world - 1 - rank):
            rank += 1
            have = 0
        sb_rank[i] = rank
        cum += cnt
        have += cnt
    rank_of_root = np.empty(n, dtype=np.int64)
    rank_of_root[roots] = sb_rank
    owner = rank_of_root[root]                       # per original link
    order = np.lexsort((np.arange(n), owner))         # by rank, original order inside a rank
    pos = np.empty(n, dtype=np.int64)
    pos[order] = np.arange(n)
    per_rank = np.bincount(owner, minlength=world)
    his = np.cumsum(per_rank)
    ranges = [(int(h - c), int(h)) for h, c in zip(his, per_rank)]

    has = down >= 0
    u = np.flatnonzero(has)                           # upstream end of every edge (original index)
    d = down[u]
    cut_edge = owner[u] != owner[d]
    # boundary links per rank, in planned order; slot = owner * max_send + position in the owner's list
    b_links = np.unique(u[cut_edge])
    b_links = b_links[np.argsort(pos[b_links], kind="stable")]
    b_owner = owner[b_links]
    n_send = np.bincount(b_owner, minlength=world)
    max_send = int(n_send.max()) if b_links.size else 0
    seg_start = np.concatenate([[0], np.cumsum(n_send)[:-1]])
    slot_of = np.full(n, -1, dtype=np.int64)
    slot_of[b_links] = b_owner * max_send + (np.arange(b_links.size) - seg_start[b_owner])

    # edges grouped by downstream link (planned position), upstream ends in ascending ORIGINAL index
    e_order = np.lexsort((u, pos[d]))
    u, d = u[e_order], d[e_order]
    d_pos = pos[d]
    p = RoutePlan(world=world, order=order, ranges=ranges, max_send=max_send,
                  n_subbasins=int(roots.size), n_cut_edges=int(cut_edge.sum()))
    for r, (lo, hi) in enumerate(ranges):
        e_lo, e_hi = np.searchsorted(d_pos, [lo, hi])
        ur, dr = u[e_lo:e_hi], d_pos[e_lo:e_hi] - lo
        local = owner[ur] == r
        idx = np.where(local, pos[ur] - lo, -(slot_of[ur] + 1))
        if (~local).any() and (slot_of[ur[~local]] < 0).any():
            raise AssertionError("cut edge without a halo slot")
        up_ptr = np.zeros(hi - lo + 1, dtype=np.int64)
        np.add.at(up_ptr, dr + 1, 1)
        up_ptr = np.cumsum(up_ptr)
        mine = b_links[b_owner == r]
        p.ranks.append(RankTopology(rank=r, lo=lo, hi=hi, up_ptr=up_ptr, up_idx=idx.astype(np.int32),
                                    send_idx=(pos[mine] - lo).astype(np.int32)))
    return p


def gather_inflow_reference(topo: RankTopology, q_local: np.ndarray, halo: np.ndarray | None) -> np.ndarray:
    """Host statement of route_gather_kernel (sequential sums in list order); tests and CPU drivers."""
    qin = np.zeros(topo.n_local)
    # sequential accumulation, one upstream entry at a time, so rounding matches the kernel
    deg = np.diff(topo.up_ptr)
    for k in range(int(deg.max()) if deg.size else 0):
        rows = np.flatnonzero(deg > k)
        e = topo.up_idx[topo.up_ptr[rows] + k].astype(np.int64)
        loc = e >= 0
        v = np.empty(rows.size)
        v[loc] = q_local[e[loc]]
        if (~loc).any():
            v[~loc] = halo[-(e[~loc] + 1)]
        qin[rows] = qin[rows] + v
    return qin


class RoutedSolver:
    """Per-interval driver of one rank's routed run on the GPU.

    solver: a tiger_hlm_gpu_b200.Solver with parameters/forcings uploaded for this rank's links (planned
    order).  dist: torch.distributed (NCCL) or None for a single rank.  Everything — window kernel (which
    packs the boundary discharge in its epilogue), all-gather, inflow gather — is queued on torch's current
    stream; nothing synchronises with the host between intervals.  exchange = "nccl": the boundary vector is
    all-gathered; "peer": the kernels store boundary discharge straight into every rank's halo vector over NVLink
    (CUDA IPC, ranks on one node) and the only collective left is a one-element all-reduce acting as the barrier
    between intervals.  The implicit fallback is switched on for
    the run: a link the explicit path abandons would otherwise freeze and starve everything downstream."""

    def __init__(self, solver, uid: int, topo: RankTopology, world: int, max_send: int, dist=None, device=None,
                 exchange: str = "nccl"):
        import torch
        self.torch = torch
        self.s, self.uid, self.topo, self.world, self.max_send, self.dist = solver, uid, topo, world, max_send, dist
        self.device = device if device is not None else torch.device("cuda", solver.device)
        # a stream of its own (torch's default stream has handle 0, which hlm_set_stream reads as "the
        # context's own stream"): kernels and collectives are ordered by it, never by the host
        self.stream = torch.cuda.Stream(self.device)
        solver.set_stream(self.stream.cuda_stream)
        solver.route_set_topology(topo.up_ptr, topo.up_idx, topo.send_idx)
        solver.set_stiff_fallback(True)
        solver.set_reject_limit(ROUTED_REJECT_LIMIT)
        self.send = self.halo = self.flag = None
        self.peer = exchange == "peer" and world > 1 and max_send > 0
        if exchange not in ("nccl", "peer"):
            raise ValueError("exchange must be 'nccl' or 'peer'")
        if self.peer:
            mine = solver.route_peer_alloc(world, topo.rank, max_send)
            handles = [None] * world
            dist.all_gather_object(handles, mine)
            solver.route_peer_open(b"".join(handles))
            with torch.cuda.stream(self.stream):
                self.flag = torch.zeros(1, dtype=torch.float32, device=self.device)
        elif world > 1 and max_send > 0:
            with torch.cuda.stream(self.stream):
                self.send = torch.zeros(max_send, dtype=torch.float64, device=self.device)
                self.halo = torch.zeros(world * max_send, dtype=torch.float64, device=self.device)
            solver.route_set_send_buffer(self.send.data_ptr())
        self._started = False
        self.exchanges = 0

    def _exchange_and_gather(self):
        if self.peer:
            # the data is already in every halo vector (stored by the kernels); all that is needed is that every
            # rank's kernel has finished before anyone reads: a stream-ordered barrier
            with self.torch.cuda.stream(self.stream):
                self.dist.all_reduce(self.flag)
            self.exchanges += 1
            self.s.route_gather(None)
        elif self.halo is not None:
            with self.torch.cuda.stream(self.stream):
                self.dist.all_gather_into_tensor(self.halo, self.send)
            self.exchanges += 1
            self.s.route_gather(self.halo.data_ptr())
        else:
            self.s.route_gather(None)

    def begin(self, y0, t0: float, tf: float, tq):
        self.s.solve_begin(self.uid, y0, t0, tf, tq)
        self.s.route_pack()
        self._exchange_and_gather()
        self._started = True
        self._first = True

    def advance(self, tf: float, tq, want_dense: bool = True):
        """Integrate the next coupling interval (the first call integrates the interval given to begin)."""
        if not self._first:
            self._exchange_and_gather()
            self.s.solve_advance(tf, tq)
        self._first = False
        nq = 0 if tq is None else len(tq)
        self.s.solve_window(nq, want_dense and nq > 0)

    def end(self):
        r = self.s.solve_end()
        if self.peer:
            self.dist.barrier()          # peers may still be storing into this rank's halo vector
            self.s.route_peer_close()
        self.s.set_stream(None)
        self.s.set_stiff_fallback(False)
        self.s.set_reject_limit(5)
        return r
