#!/usr/bin/env python
"""bench.py — accepted RK45 system-steps/s of the batched Dormand–Prince path on B200.

Contract: `python bench.py --gpus N --steps K --warmup W` (for N > 1 launched under torchrun, one
rank per GPU).  One JSON line on rank 0.

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): Model204, synthetic
links (SURVEY §8(d) inputs), 1-year hourly precipitation + daily temperature on a forcing grid,
hourly dense output, FP64.  A *step* is one output window = one simulated day for every link of the
rank (24 hourly queries, ~450 accepted RK45 steps per link).  Links are sharded over ranks with no
data-path collective (they are independent); scaling is weak: --links-per-gpu links
on every rank.

  value  accepted system-steps/s with state, parameters and forcings resident in HBM and the dense
         window left on the device (hlm_solve_window), CUDA events on the launching stream,
         max over ranks.
  e2e    the same metric through the reference-facing operator hlm_run_rk45 (the C ABI under
         rk45_api::run_rk45<T>) with HOST buffers: every step uploads y0 from pinned memory and
         downloads final + dense states.
  roofline / cpu_baseline / reference_cuda / clocks: see DESIGN.md §Measurement.

`--impl reference` times the reference's own CPU implementation of the path instead: its
rk45_step / rk45_dense / Model204::rhs templates compiled for the host from the reference sources
(oracle/_ref/libref_host.so) on all host cores, each step a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# STATIC figures from a committed ncu capture, not measured live (ncu counters are not available inside a timed run):
# re-derive them with tools/ncu_summary.py whenever the window kernel changes.  PROFILE names the capture.
PROFILE = {"file": "profiles/r2s_ncu_full_window_kernel_10M.csv", "commit": "62666ca", "dram_bytes_per_link_per_launch": 1408.4,
           "fp64_pipe_utilization": 0.663}
DRAM_BYTES_PER_LINK_PER_LAUNCH = PROFILE["dram_bytes_per_link_per_launch"]
W_MIN_FLOP_PER_ATTEMPT = 561.0  # minimal-algorithm nominal FP64 flop per attempted Model204 step
PRM6 = [1e-6, 1e-6, 1e-9, 0.9, 0.2, 10.0]  # initialStep (main.cpp:633-640), rtol, atol, safety, min/maxScale
DAY = 1440.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="model204", choices=["model204", "model200", "routed"],
                    help="model204 = BASELINE configs[3] (the headline); model200 = configs[2]: the project-defined hillslope-"
                         "link model, unrouted, 1M links, 30 days of forcing; routed = configs[4]: Model 200 links coupled "
                         "through upstream discharge, partitioned by sub-basin, boundary links exchanged over NCCL")
    ap.add_argument("--links-per-gpu", type=int, default=None, help="default 10M (model204) / 1M (model200) / 2.5M (routed)")
    ap.add_argument("--couple-minutes", type=float, default=15.0, help="routed: coupling interval")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "peer"],
                    help="routed: boundary discharge all-gathered over NCCL, or stored into every rank's halo vector by the "
                         "kernels themselves (CUDA IPC peer memory) with a one-element all-reduce as the barrier")
    ap.add_argument("--schedule", default="auto", choices=["auto", "tiles", "lanes", "sorted"],
                    help="how links are dealt to lanes (hlm_set_schedule); auto = sorted tiles (tiles of links with equal attempt counts "
                         "in the previous launch) for Model 200 and routed runs, plain tiles otherwise")
    ap.add_argument("--days", type=int, default=365, help="length of the forcing record / run horizon")
    ap.add_argument("--wet-fraction", type=float, default=0.0,
                    help="share of links started with surface storage so Model204's pow() branch runs")
    ap.add_argument("--precision", type=int, default=64, choices=[64, 32])
    ap.add_argument("--rtol", type=float, default=None, help="override rtol (default 1e-6)")
    ap.add_argument("--atol", type=float, default=None, help="override atol (default 1e-9)")
    ap.add_argument("--stiff-fallback", action="store_true",
                    help="continue links the RK45 path flags stiff with the Radau IIA fallback (always on for model200/routed)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-baselines", action="store_true", help="skip cpu_baseline / reference_cuda legs and every sub-record")
    ap.add_argument("--no-subrecords", action="store_true",
                    help="skip the sub-records of the default line (fp32, wet, strong, e2e_selected, e2e_shim, model200, routed)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU work of the cpu_baseline sample")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples: timed region shorter than the 200 ms sampling period"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def make_inputs(ns: int, days: int, wet_fraction: float, rank: int):
    from tiger_hlm_gpu_b200 import synthetic
    sp = synthetic.make_spatial_params(ns, seed=204 + rank)
    col, ncells = synthetic.make_cells(ns)
    pr, t2m = synthetic.make_forcing_grid(ncells, days, seed=2019 + rank)
    y0 = synthetic.make_y0(ns, wet_fraction, seed=7 + rank)
    return sp, col, ncells, pr, t2m, y0


# ------------------------------------------------------------------------------------------------
def cpu_sample_inputs(ns_sample: int, args):
    """A bounded sample of the SAME workload: the first ns_sample links of rank 0's shard, day 0."""
    from tiger_hlm_gpu_b200 import synthetic
    sp = synthetic.make_spatial_params(ns_sample, seed=204)
    col, ncells = synthetic.make_cells(ns_sample)
    pr, t2m = synthetic.make_forcing_grid(ncells, 2, seed=2019)
    y0 = synthetic.make_y0(ns_sample, args.wet_fraction, seed=7)
    tq = synthetic.hourly_queries(0.0, DAY)
    return sp, col, pr, t2m, y0, tq


def time_cpu(kind: str, ns_sample: int, threads: int, args):
    """Accepted steps/s of the CPU path on `threads` host threads over ns_sample links x 1 day."""
    from tiger_hlm_gpu_b200 import synthetic
    sp, col, pr, t2m, y0, tq = cpu_sample_inputs(ns_sample, args)
    t0 = time.perf_counter()
    if kind == "reference":
        from tests import refs
        blocks = [synthetic.expand_forcing_per_link(pr, col), synthetic.expand_forcing_per_link(t2m, col)]
        r = refs.ref_host_run204(PRM6, y0, 0.0, DAY, tq, sp, blocks, [1.0, 24.0], threads=threads)
    else:
        from oracle import oracle as O
        r = O.run_rk45(204, O.Params.make(*PRM6), y0, 0.0, DAY, tq, sp=sp,
                       forcing=O.Forcing([pr, t2m], [1.0, 24.0], col=col), threads=threads)
    dt = time.perf_counter() - t0
    return float(r["n_accept"].sum()), dt


def cpu_baseline(kind: str, args, target_s: float):
    cores = os.cpu_count() or 1
    pilot_ns = 64 * cores
    acc, dt = time_cpu(kind, pilot_ns, cores, args)
    rate = acc / dt
    ns_sample = int(min(max(pilot_ns, target_s * rate / max(acc / pilot_ns, 1.0)), 4_000_000))
    acc, dt = time_cpu(kind, ns_sample, cores, args)
    return {"value": acc / dt, "unit": "accepted system-steps/s", "cores": cores, "kind": kind,
            "sample": f"{ns_sample} links x 1 simulated day (day 0 of the bench workload, hourly dense output), "
                      f"{int(acc)} accepted steps in {dt:.2f} s wall on {cores} threads",
            "ns_sample": ns_sample, "seconds": dt}


def run_reference_arm(args):
    """--impl reference: the reference's own step/dense/rhs code on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tests import refs
    kind = "reference" if refs.have("libref_host.so") else "port"
    cores = os.cpu_count() or 1
    # size one step for ~4 s so that warmup+steps end within a few minutes
    acc, dt = time_cpu(kind, 64 * cores, cores, args)
    per_link = acc / (64 * cores)
    ns_sample = int(max(64 * cores, 4.0 * (acc / dt) / per_link))
    for _ in range(args.warmup):
        time_cpu(kind, ns_sample, cores, args)
    tot_acc, tot_dt = 0.0, 0.0
    for _ in range(args.steps):
        a, d = time_cpu(kind, ns_sample, cores, args)
        tot_acc += a
        tot_dt += d
    value = tot_acc / tot_dt
    sample = (f"{ns_sample} links x 1 simulated day per step (first links of the bench workload, hourly dense "
              f"output), {cores} host threads; step/dense/rhs are the reference's own templates compiled for the "
              f"host, the driver loop restates solver/rk45_kernel.cu:53-164")
    line = {
        "impl": "reference", "metric": "accepted RK45 system-steps/sec", "value": value,
        "unit": "accepted system-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.links_per_gpu),
        "cpu_baseline": {"value": value, "unit": "accepted system-steps/s", "cores": cores, "kind": kind,
                         "sample": sample, "sample_links_per_step": ns_sample},
        "e2e": {"value": value, "unit": "accepted system-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
W_MIN_FLOP_PER_ATTEMPT_200 = 585.0  # as W_min with Model 200's rhs (36 nominal flop instead of 32), DESIGN.md
UNIT = "accepted system-steps/s"


class Env:
    """Ranks, devices and the two collectives the bench itself needs (barrier, reductions)."""

    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, ms, sums):
        from tiger_hlm_gpu_b200.sharding import reduce_timing
        return reduce_timing(ms, sums, self.dist, self.dev)

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def routed_partition_check(env, args):
    """At N > 1, before anything is timed: a small routed network integrated by the N ranks with the exchange under
    test, and by rank 0 alone; per-link final states and accepted-step counts must be the same bits (inflow sums run
    in ascending original link index, so every partition gives the same arithmetic).  GPU against GPU: a property of
    the path, not a comparison with the test-side restatement."""
    import tiger_hlm_gpu_b200 as hlm
    from tiger_hlm_gpu_b200 import routing, synthetic
    torch, dist, world, rank = env.torch, env.dist, env.world, env.rank
    ns, sub, dt, n_int = 40960, 512, args.couple_minutes, 12
    sp = synthetic.apply_network(synthetic.make_spatial_params(ns, seed=5), synthetic.make_network(ns, subbasin_links=sub, seed=9))
    col, ncells = synthetic.make_cells(ns, links_per_cell=97)
    pr, t2m = synthetic.make_forcing_grid(ncells, 2, seed=31)
    y0 = np.tile(np.array(synthetic.Y0_200), (ns, 1))
    y0[:, 0] = np.random.default_rng(3).uniform(0.05, 5.0, ns)

    def run(world_n, rank_n, dist_n):
        plan = routing.plan(sp["stream"], sp["next_stream"], world_n, subbasin_links=sub)
        topo = plan.ranks[rank_n]
        sel = plan.order[topo.lo:topo.hi]
        solver = hlm.Solver(env.local_rank)
        solver.set_model_parameters(200, hlm.Parameters(*PRM6))
        solver.set_max_attempts(5_000_000)
        solver.upload_spatial_params(sp[sel])
        solver.upload_forcing(0, 1.0, pr)
        solver.upload_forcing(1, 24.0, t2m)
        solver.set_forcing_columns(col[sel])
        rs = routing.RoutedSolver(solver, 200, topo, world_n, plan.max_send, dist_n, exchange=args.exchange)
        for i in range(n_int):
            tq = np.array([dt * (i + 1)])
            if i == 0:
                rs.begin(y0[sel], 0.0, dt, tq)
            rs.advance(dt * (i + 1), tq, want_dense=False)
        r = rs.end()
        n_ex = rs.exchanges
        solver.close()
        return sel, r, plan, n_ex

    sel, r, plan, n_ex = run(world, rank, dist)
    fin = torch.zeros(ns, 5, dtype=torch.float64, device=env.dev)
    na = torch.zeros(ns, dtype=torch.float64, device=env.dev)
    idx = torch.as_tensor(sel, device=env.dev)
    fin[idx] = torch.as_tensor(r["final"], device=env.dev)
    na[idx] = torch.as_tensor(r["n_accept"].astype(np.float64), device=env.dev)
    dist.all_reduce(fin)  # every link is owned by exactly one rank: the sum is a gather
    dist.all_reduce(na)
    out = None
    if rank == 0:
        sel1, r1, _, _ = run(1, 0, None)
        fin1 = np.zeros((ns, 5))
        na1 = np.zeros(ns)
        fin1[sel1] = r1["final"]
        na1[sel1] = r1["n_accept"]
        same = bool(np.array_equal(fin.cpu().numpy(), fin1) and np.array_equal(na.cpu().numpy(), na1))
        out = {"bit_identical_to_one_rank": same, "links": ns, "intervals": n_int, "ranks": world, "exchanges": n_ex,
               "cut_edges": plan.n_cut_edges, "halo_doubles": plan.halo_len, "accepted_steps": float(na1.sum()),
               "what": "final states and accepted-step counts of every link: N ranks with the exchange under test vs rank 0 alone"}
    dist.barrier()
    return out


def routed_record(env, args, K, W, links_per_gpu, with_e2e=True):
    """BASELINE configs[4]: Model 200 on a synthetic river network, partitioned by sub-basin over the ranks,
    boundary discharge exchanged once per coupling interval (NCCL all-gather, or peer-memory stores).  One step =
    one simulated hour (60 / couple_minutes intervals).  Weak scaling: links_per_gpu links per rank.
    Returns the record on rank 0 (None elsewhere)."""
    import tiger_hlm_gpu_b200 as hlm
    from tiger_hlm_gpu_b200 import routing, synthetic

    torch, dist, dev, world, rank = env.torch, env.dist, env.dev, env.world, env.rank
    partition_check = routed_partition_check(env, args) if world > 1 else None
    ns_all = links_per_gpu * world
    dt = args.couple_minutes
    n_int = max(1, int(round(60.0 / dt)))
    days = max(2, (W + 2 * K + 24) // 24 + 1)
    # every rank builds the same network and plan (seeded), then keeps its own part
    sp_all = synthetic.apply_network(synthetic.make_spatial_params(ns_all), synthetic.make_network(ns_all))
    plan = routing.plan(sp_all["stream"], sp_all["next_stream"], world)
    topo = plan.ranks[rank]
    sel = plan.order[topo.lo:topo.hi]
    col_all, ncells = synthetic.make_cells(ns_all)
    pr, t2m = synthetic.make_forcing_grid(ncells, days)
    rng = np.random.default_rng(7)
    y0 = np.tile(np.array(synthetic.Y0_200), (sel.size, 1))
    y0[:, 0] = rng.uniform(0.05, 5.0, ns_all)[sel]
    solver = hlm.Solver(env.local_rank)
    solver.set_model_parameters(200, hlm.Parameters(*PRM6))
    solver.set_schedule(args.schedule)
    solver.set_max_attempts(5_000_000)
    solver.upload_spatial_params(sp_all[sel])
    solver.upload_forcing(0, 1.0, pr)
    solver.upload_forcing(1, 24.0, t2m)
    solver.set_forcing_columns(col_all[sel])
    del sp_all
    fp_peak = solver.measure_fma_peak(64)
    rs = routing.RoutedSolver(solver, 200, topo, world, plan.max_send, dist, exchange=args.exchange)
    stream = rs.stream

    # hourly output like the other workloads, and of the routed quantity: one record of the channel discharge per link
    # at the end of the hour's last coupling interval (hlm_set_output_states: only that state is interpolated and stored)
    solver.set_output_states([0])
    no_tq = np.zeros(0)

    def hour(k, first=False):
        for i in range(n_int):
            tf = 60.0 * k + dt * (i + 1)
            tq = np.array([tf]) if i == n_int - 1 else no_tq
            if first and i == 0:
                rs.begin(y0, 0.0, tf, tq)
            rs.advance(tf, tq)

    for k in range(W):
        hour(k, first=(k == 0))
    solver.synchronize()
    tot0 = solver.solve_totals()
    solver.kernel_time_ms()
    launches0 = solver.launch_count()
    ex0 = rs.exchanges
    sampler = ClockSampler(env.local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    e0.record(stream)
    for k in range(W, W + K):
        hour(k, first=(k == 0))
    e1.record(stream)
    env.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    kern_ms, kern_n = solver.kernel_time_ms()
    launches = solver.launch_count() - launches0
    exchanges = rs.exchanges - ex0
    tot1 = solver.solve_totals()
    radau = int(solver.solve_radau_steps().sum())
    acc = tot1["n_accept"] - tot0["n_accept"]
    attempts = acc + (tot1["n_reject"] - tot0["n_reject"]) + (tot1["n_jump"] - tot0["n_jump"])
    state = dict(tot1)

    # e2e: the same intervals, with the host in the loop as a caller writing hourly output has it: the hour's
    # discharge record of every link is copied to pinned host memory every step, and the first step uploads y0.
    # (Between intervals nothing else crosses PCIe: state stays resident.)
    e2e = None
    if with_e2e:
        win = torch.zeros((sel.size, 1, 1), dtype=torch.float64).pin_memory()
        env.barrier()
        t_start = time.perf_counter()
        acc_before = solver.solve_totals()["n_accept"]
        for k in range(W + K, W + 2 * K):
            hour(k)
            solver.solve_wait_copy(solver.solve_fetch_window_packed(win.numpy()))
        env.barrier()
        e_ms = (time.perf_counter() - t_start) * 1e3
        acc_e = solver.solve_totals()["n_accept"] - acc_before
        e_ms_max, (acc_e_all,) = env.reduce(e_ms, [acc_e])
        e2e = {"value": acc_e_all / (e_ms_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * n_int,
               "d2h_bytes_per_step": int(sel.size) * 8 + 56, "ms_per_step": e_ms_max / K,
               "api": "routing.RoutedSolver over the C ABI (hlm_route_gather + hlm_solve_advance + hlm_solve_window + "
                      "hlm_solve_fetch_window_packed), pinned host buffer; state resident between intervals"}
    peer = rs.peer
    rs.end()
    solver.close()

    ms_max, (acc_all, att_all, launches_all, kern_ms_all, kern_n_all, radau_all) = env.reduce(
        ms, [acc, attempts, launches, kern_ms, kern_n, radau])
    if rank != 0:
        return None
    kern_avg_ms = kern_ms_all / max(kern_n_all, 1)
    att_per_launch = att_all / max(kern_n_all, 1)
    achieved = W_MIN_FLOP_PER_ATTEMPT_200 * att_per_launch / (kern_avg_ms * 1e-3) / 1e12
    return {
        "metric": "accepted RK45 system-steps/sec", "value": acc_all / (ms_max * 1e-3), "unit": UNIT,
        "n_gpus": world, "steps": K, "warmup": args.warmup, "warmup_done": W, "ms_per_step": ms_max / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "routed sub-basin network (BASELINE configs[4]): Model 200 (project-defined), synthetic "
                               "river network, links coupled through upstream discharge held over a coupling interval; "
                               "one step = one simulated hour",
                   "links_per_gpu": links_per_gpu, "links_total": ns_all, "couple_minutes": dt,
                   "intervals_per_step": n_int, "schedule": args.schedule, "sub_basins": plan.n_subbasins, "cut_edges": plan.n_cut_edges,
                   "halo_doubles": plan.halo_len, "rtol": PRM6[1], "atol": PRM6[2],
                   "output": "hourly channel discharge of every link (state 0), interpolated at the hour by the last interval's launch",
                   "reject_limit": routing.ROUTED_REJECT_LIMIT,
                   "parallelism": (f"sub-basins dealt to {world} GPU(s); " +
                                   ("boundary discharge stored into peer memory by the kernels, one barrier per interval" if peer
                                    else "one NCCL all-gather of the boundary vector per interval"))
                                  if world > 1 else "1 GPU, no exchange",
                   "l2": "inputs larger than L2 (state + parameters of the rank's links >> 126 MB)"},
        "accepted_steps_per_step": acc_all / K, "attempts_per_accepted": att_all / max(acc_all, 1.0),
        "implicit_steps_total": radau_all, "exchanges_per_step": exchanges / K, "partition_check": partition_check,
        "link_status_after_run": {k: state[k] for k in ("active", "done", "stiff", "stalled")},
        "e2e": e2e, "gpu_launches": int(launches_all),
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp_peak, "unit": "TFLOP/s", "frac": achieved / fp_peak,
                     "traffic": None, "flop_per_attempt": W_MIN_FLOP_PER_ATTEMPT_200, "attempts_per_launch": att_per_launch,
                     "kernel_ms_avg": kern_avg_ms, "kernel": ("hlm::rk45_lanes_kernel<Model200,double,true>" if args.schedule == "lanes" else
                                "hlm::rk45_window_kernel<Model200,double> over sorted tiles (the first interval: rk45_lanes_kernel)"),
                     "kernel_share_of_step": kern_ms_all / max(world, 1) / ms_max,
                     "peak_source": "measured live (hlm_measure_fma_peak)"},
        "cpu_baseline": None, "clocks": clocks,
    }


def w_min_for(uid):
    return W_MIN_FLOP_PER_ATTEMPT if uid == 204 else W_MIN_FLOP_PER_ATTEMPT_200


def workload_config(args, ns):
    name = ("Model204 hillslope-link runoff, synthetic links (SURVEY 8(d) inputs), 1-year hourly " if args.workload != "model200"
            else "Model 200 (project-defined hillslope-link runoff, unrouted; BASELINE configs[2]), synthetic links, 30 days of hourly ")
    return {"workload": name + "pr + daily t2m forcing grid, hourly dense output; one step = one simulated day "
                        "(24 queries) for every link",
            "links_per_gpu": ns, "days_of_forcing": args.days, "queries_per_step": 24,
            "wet_fraction": args.wet_fraction, "rtol": PRM6[1], "atol": PRM6[2], "initial_step": PRM6[0],
            "parallelism": f"links sharded over {args.gpus} GPU(s), no collective",
            "l2": "inputs larger than L2 (per-step state+parameter+output traffic >> 126 MB)"}


def day_queries(k):
    return k * DAY + 60.0 * np.arange(1, 25)


def resident_steps(env, solver, stream, uid, y0, K, W, sample_clocks=False, want_final=False):
    """W untimed + K timed steps with state, parameters and forcings resident and the dense window left on the
    device.  One step = one simulated day: hlm_solve_restart (new interval from the resident final state, what a
    chained run_rk45 does) + hlm_solve_window (the hot kernel).  The run is driven in day-sized intervals because
    the reference's stiffness threshold scales with (tf - t0).  CUDA events on the launching stream, max over
    ranks; counts summed over ranks."""
    torch = env.torch
    solver.solve_begin(uid, y0, 0.0, DAY, day_queries(0))
    for k in range(W):
        if k:
            solver.solve_restart(k * DAY, (k + 1) * DAY, day_queries(k))
        solver.solve_window(24, True)
    solver.synchronize()
    tot0 = solver.solve_totals()
    solver.kernel_time_ms()  # drop warm-up kernel timings
    launches0 = solver.launch_count()
    sampler = ClockSampler(env.local_rank) if (sample_clocks and env.rank == 0) else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    e0.record(stream)
    for k in range(W, W + K):
        if k:
            solver.solve_restart(k * DAY, (k + 1) * DAY, day_queries(k))
        solver.solve_window(24, True)
    e1.record(stream)
    env.barrier()
    clocks = sampler.stop() if sampler else None
    ms = e0.elapsed_time(e1)
    kern_ms, kern_n = solver.kernel_time_ms()
    launches = solver.launch_count() - launches0
    tot1 = solver.solve_totals()
    acc = tot1["n_accept"] - tot0["n_accept"]
    attempts = acc + (tot1["n_reject"] - tot0["n_reject"]) + (tot1["n_jump"] - tot0["n_jump"])
    end = solver.solve_end()
    final = end["final"] if want_final else None
    ms_max, (acc_all, att_all, launches_all, kern_ms_all, kern_n_all) = env.reduce(ms, [acc, attempts, launches, kern_ms, kern_n])
    return {"ms": ms_max, "acc": acc_all, "attempts": att_all, "launches": launches_all, "kern_ms": kern_ms_all,
            "kern_n": kern_n_all, "state": {k: tot1[k] for k in ("active", "done", "stiff", "stalled")}, "clocks": clocks,
            "final": final, "value": acc_all / (ms_max * 1e-3)}


def roofline_of(r, uid, peak, world):
    """Nominal FP64 (or FP32) roofline of the window kernel from the library's own per-launch CUDA events."""
    kern_avg_ms = r["kern_ms"] / max(r["kern_n"], 1)
    att_per_launch = r["attempts"] / max(r["kern_n"], 1)
    achieved = w_min_for(uid) * att_per_launch / (kern_avg_ms * 1e-3) / 1e12
    return {"achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "flop_per_attempt": w_min_for(uid), "attempts_per_launch": att_per_launch, "kernel_ms_avg": kern_avg_ms,
            "kernel_share_of_step": r["kern_ms"] / max(world, 1) / r["ms"]}


def e2e_steps(env, solver, stream, uid, ns, y0, K, W, states=None, out_bits=64):
    """The same metric through the reference-facing operator hlm_run_rk45 (the C ABI under rk45_api::run_rk45<T>)
    with HOST buffers: every step uploads y0 from pinned memory and downloads final states, codes, counters and the
    day's dense records.  states / out_bits: hlm_set_output_states / hlm_set_output_precision (the records the
    caller keeps, cut down and narrowed on the device)."""
    import ctypes as C
    import tiger_hlm_gpu_b200 as hlm
    torch = env.torch
    nq_w = 24
    solver.set_output_states(states)
    solver.set_output_precision(out_bits)
    ncol, dt = solver.output_layout(uid)
    tdt = torch.float32 if dt == np.float32 else torch.float64
    y_host = torch.from_numpy(np.ascontiguousarray(y0).copy()).pin_memory()
    f_host = torch.zeros((ns, 5), dtype=torch.float64).pin_memory()
    d_host = torch.zeros((ns, nq_w, ncol), dtype=tdt).pin_memory()
    stiff = torch.zeros(ns, dtype=torch.int32).pin_memory()
    na = torch.zeros(ns, dtype=torch.int64).pin_memory()
    lib = hlm.load_library()
    bufs = [y_host, f_host]  # next day's initial state = this day's final state: the two pinned buffers swap roles

    def one_step(k):
        t0 = k * DAY
        tqw = t0 + 60.0 * np.arange(1, nq_w + 1)
        y_in, y_out = bufs[k % 2], bufs[(k + 1) % 2]
        rc = lib.hlm_run_rk45(solver._h, uid, C.c_void_p(y_in.data_ptr()), ns, t0, t0 + DAY,
                              tqw.ctypes.data_as(C.c_void_p), nq_w, C.c_void_p(y_out.data_ptr()),
                              C.c_void_p(d_host.data_ptr()), C.c_void_p(stiff.data_ptr()),
                              C.c_void_p(na.data_ptr()), None, None)
        if rc != 0:
            raise hlm.HlmError(lib.hlm_last_error().decode())
        # the step count for the throughput figure comes from the device-side totals (56 bytes), not from a
        # host pass over the 10 M counters the operator has just returned in `na`
        return solver.solve_totals()["n_accept"]

    try:
        for k in range(W):
            one_step(k)
        env.barrier()
        t_start = time.perf_counter()
        ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ee0.record(stream)
        acc_e = 0
        for k in range(W, W + K):
            acc_e += one_step(k)
        ee1.record(stream)
        env.barrier()
        wall_ms = (time.perf_counter() - t_start) * 1e3
    finally:
        solver.set_output_states(None)
        solver.set_output_precision(64)
    # the call is synchronous at its end (results are in host memory), so wall clock and the event
    # pair bracket the same work; report the larger
    e_ms = max(ee0.elapsed_time(ee1), wall_ms)
    e_ms_max, (acc_e_all,) = env.reduce(e_ms, [acc_e])
    h2d = ns * 5 * 8 + nq_w * 8
    d2h = ns * 5 * 8 + ns * nq_w * ncol * d_host.element_size() + ns * 4 + ns * 8
    rec = {"value": acc_e_all / (e_ms_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": e_ms_max / K,
           "api": f"hlm_run_rk45 (C ABI under rk45_api::run_rk45<Model{uid}>), pinned host buffers"}
    if states is not None or out_bits != 64:
        rec["output"] = {"states": list(states) if states is not None else "all", "value_type": "f32" if out_bits == 32 else "f64",
                         "api": "hlm_set_output_states + hlm_set_output_precision (config.yaml output.states / output.precision)"}
    return rec


def reference_cuda_record(sp, col, pr, t2m, y0, ns, label):
    """The UNCHANGED reference kernel built for sm_100a (oracle/_ref/libref_cuda.so) on the same links and the
    same day as the repo's arm, kernel time only, at the largest size its 32-bit dense index allows."""
    from tests import refs
    from tiger_hlm_gpu_b200 import synthetic
    if not refs.have("libref_cuda.so"):
        return {"unavailable": "oracle/_ref/libref_cuda.so was not built (no /root/reference at build time)"}
    nq = 25
    ns_r = int(min(ns, ((1 << 31) - 1) // (5 * nq)))
    blocks = [synthetic.expand_forcing_per_link(pr[:48], col[:ns_r]), synthetic.expand_forcing_per_link(t2m[:2], col[:ns_r])]
    tq_r = 60.0 * np.arange(0, nq)
    a = (PRM6, y0[:ns_r], 0.0, DAY, tq_r, sp[:ns_r], blocks, [1.0, 24.0])
    refs.ref_cuda_run204(*a, counted=False, want_dense=False)  # warm-up (module load, first touch)
    r = refs.ref_cuda_run204(*a, counted=True, want_dense=False)
    p = refs.ref_cuda_run204(*a, counted=False, want_dense=False)
    return {"value": float(r["n_accept"].sum()) / (p["kernel_ms"] * 1e-3), "unit": UNIT, "kernel_ms": p["kernel_ms"],
            "links": ns_r, "not_finished": int((r["stiff"] != 0).sum()),
            "sample": f"unchanged reference kernel rk45_then_radau_multi<Model204> built for sm_100a, {ns_r} links x day 0 "
                      f"of the bench workload ({label}), 25 hourly queries, 1-D launch <<<ceil(ns/128),128>>>, per-link "
                      f"expanded forcing; kernel time only"}


def model200_record(env, args, K, W):
    """BASELINE configs[2]: Model 200 (project-defined), unrouted, 1 M links, 30 days of forcing, FP64, 1 GPU."""
    import tiger_hlm_gpu_b200 as hlm
    from tiger_hlm_gpu_b200 import routing
    ns, days = 1_000_000, 30
    K = min(K, days - W)
    sp, col, ncells, pr, t2m, y0 = make_inputs(ns, days, 0.0, env.rank)
    y0[:, 0] = np.random.default_rng(7 + env.rank).uniform(0.05, 5.0, ns)  # channel discharge in place of the snow store
    solver = hlm.Solver(env.local_rank)
    stream = env.torch.cuda.Stream(device=env.dev)
    solver.set_stream(stream.cuda_stream)
    solver.set_model_parameters(200, hlm.Parameters(*PRM6))
    solver.set_stiff_fallback(True)
    solver.set_max_attempts(5_000_000)
    # as routed runs and hlm_run do for Model 200: with the reference's limit of 5 consecutive rejections the first day's
    # transient flags 1.5 % of the links stiff that are not, and the implicit fallback spends seconds on them
    solver.set_reject_limit(routing.ROUTED_REJECT_LIMIT)
    solver.upload_spatial_params(sp)
    solver.upload_forcing(0, 1.0, pr)
    solver.upload_forcing(1, 24.0, t2m)
    solver.set_forcing_columns(col)
    peak = solver.measure_fma_peak(64)
    with env.torch.cuda.stream(stream):
        r = resident_steps(env, solver, stream, 200, y0, K, W)
        # the same with water on every link's surface (the surface store's h^(2/3) then runs in every right-hand side; a
        # store that has been wet decays like t^-3/2 and never returns to 0, so this is the long-run state of a real run)
        y0w = y0.copy()
        y0w[:, 2] = 0.01
        rw = resident_steps(env, solver, stream, 200, y0w, K, W)
    solver.close()
    roof = roofline_of(r, 200, peak, 1)
    return {"value": r["value"], "unit": UNIT, "ms_per_step": r["ms"] / K, "steps": K, "links": ns, "days_of_forcing": days,
            "wet": {"value": rw["value"], "ms_per_step": rw["ms"] / K, "link_status_after_run": rw["state"],
                    "what": "every link starts with 0.01 m on its surface store"},
            "attempts_per_accepted": r["attempts"] / max(r["acc"], 1.0), "accepted_steps_per_step": r["acc"] / K,
            "link_status_after_run": r["state"], "roofline_frac": roof["frac"], "kernel_ms_avg": roof["kernel_ms_avg"],
            "schedule": "sorted tiles (auto)", "reject_limit": routing.ROUTED_REJECT_LIMIT, "note": "Model 200 is project-defined (the reference names it, README.md:95, and ships no "
            "definition): parity unpinned; one step = one simulated day, hourly dense output, implicit fallback on"}


def shim_record(args, ns, K, W):
    """e2e through the C++ shim the integration guide tells a reference-side caller to use:
    rk45_api::run_rk45<Model204>() of include/hlm_b200/rk45_api.hpp (host/bench_shim.cpp), same workload."""
    exe = os.path.join(ROOT, "tiger_hlm_gpu_b200", "host", "build", "hlm_bench_shim")
    if not os.path.exists(exe):
        return {"unavailable": "host/build/hlm_bench_shim is not built"}
    out = subprocess.run([exe, str(ns), str(K), str(W)], capture_output=True, text=True, timeout=900)
    if out.returncode != 0:
        return {"error": (out.stderr or out.stdout)[-400:]}
    return json.loads(out.stdout.strip().splitlines()[-1])


def guarded(fn, *a, **kw):
    """A sub-record must never cost the headline line: report the failure in its place."""
    try:
        return fn(*a, **kw)
    except Exception as ex:
        return {"error": repr(ex)[:400]}


def run_routed_arm(args):
    env = Env()
    K, W = args.steps, max(args.warmup, 3)  # never fewer than 3 untimed steps before a timed region (timing rules)
    line = routed_record(env, args, K, W, args.links_per_gpu, with_e2e=not args.no_e2e)
    if env.rank == 0:
        print(json.dumps(line), flush=True)
    env.close()


# ------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.rtol is not None:
        PRM6[1] = args.rtol
    if args.atol is not None:
        PRM6[2] = args.atol
    if args.links_per_gpu is None:
        args.links_per_gpu = {"model204": 10_000_000, "model200": 1_000_000, "routed": 2_500_000}[args.workload]
    if args.workload == "model200":
        if args.days == 365:
            args.days = 30
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the reference names model 200 (README.md:95) but ships no "
                              "definition of it: there is no reference implementation of this workload"}))
            return
    if args.workload == "routed":
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the reference couples no links (SURVEY 8(a) row 9): "
                              "there is no reference implementation of the routed workload"}))
            return
        run_routed_arm(args)
        return
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import tiger_hlm_gpu_b200 as hlm
    from tiger_hlm_gpu_b200 import synthetic
    from tiger_hlm_gpu_b200.sharding import bind_to_gpu_numa_node, shard_range

    env = Env()
    torch, world, rank, local_rank, dev = env.torch, env.world, env.rank, env.local_rank, env.dev
    ns = args.links_per_gpu
    K, W = args.steps, max(args.warmup, 3)  # never fewer than 3 untimed steps before a timed region (timing rules)
    assert W + K <= args.days, "not enough forcing days for warmup+steps"
    Ks = min(K, 5)  # timed steps of the sub-records (wet, fp32, strong, selected-output e2e, model200, routed)
    numa = bind_to_gpu_numa_node(local_rank)  # before any pinned allocation
    subs = not args.no_baselines and not args.no_subrecords

    uid = 200 if args.workload == "model200" else 204
    sp, col, ncells, pr, t2m, y0 = make_inputs(ns, args.days, args.wet_fraction, rank)
    if uid == 200:  # channel discharge in place of the snow store
        y0[:, 0] = np.random.default_rng(7 + rank).uniform(0.05, 5.0, ns)
    solver = hlm.Solver(local_rank)
    # a dedicated non-default stream: the library treats handle 0 as "use my own stream", and torch
    # events only see the stream they are recorded on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    solver.set_stream(stream.cuda_stream)
    solver.set_precision(args.precision)
    solver.set_schedule(args.schedule)
    solver.set_model_parameters(uid, hlm.Parameters(*PRM6))
    solver.set_stiff_fallback(args.stiff_fallback or uid == 200)
    if uid == 200:  # see model200_record
        from tiger_hlm_gpu_b200 import routing
        solver.set_reject_limit(routing.ROUTED_REJECT_LIMIT)
    solver.set_max_attempts(5_000_000)
    solver.upload_spatial_params(sp)
    solver.upload_forcing(0, 1.0, pr)
    solver.upload_forcing(1, 24.0, t2m)
    solver.set_forcing_columns(col)
    fp_peak = solver.measure_fma_peak(args.precision)

    # ---------------- resident arm: `value` ----------------
    main_r = resident_steps(env, solver, stream, uid, y0, K, W, sample_clocks=True)
    value = main_r["value"]

    # ---------------- FP32 arm (BASELINE configs[3]: "FP32 vs FP64"): same days, same links, state and stages in FP32 ----
    fp32 = None
    if args.precision == 64 and subs:
        solver.set_precision(32)
        peak32 = solver.measure_fma_peak(32)
        r32 = resident_steps(env, solver, stream, uid, y0, Ks, W)
        solver.set_precision(64)
        fp32 = {"value": r32["value"], "unit": UNIT, "ms_per_step": r32["ms"] / Ks, "steps": Ks,
                "attempts_per_accepted": r32["attempts"] / max(r32["acc"], 1.0),
                "roofline_frac": roofline_of(r32, uid, peak32, world)["frac"], "peak_tflops": peak32,
                "link_status_after_run": r32["state"],
                "note": "hlm_set_precision(32): FP32 state, stages and error control at the FP64 run's tolerances (rtol 1e-6 is ~8 "
                        "float epsilons).  No reference counterpart (the reference is FP64 only).  Accuracy at exactly these settings is "
                        "asserted in tests/test_gpu_parity.py::test_fp32_mode_at_the_bench_tolerances: within 25*(atol+rtol*|y|) of FP64 "
                        "under time-constant forcing; on this workload, whose states depend on the step sequence through the reference's "
                        "step-start forcing sampling (SURVEY F7), within 8*(atol+rtol*|y|) + 2x FP64's own change under 10x tolerances"}

    # ---------------- wet arm: every link with surface storage, so Model204's pow() runs in every rhs ----------------
    wet = None
    y0_wet = None
    if subs and uid == 204 and args.precision == 64 and args.wet_fraction < 1.0:
        y0_wet = synthetic.make_y0(ns, 1.0, seed=7 + rank)
        rw = resident_steps(env, solver, stream, uid, y0_wet, Ks, W, want_final=True)
        roof_w = roofline_of(rw, uid, fp_peak, world)
        still_wet = int((rw["final"][:, 2] != 0.0).sum())
        (still_wet_all,) = env.reduce(0.0, [still_wet])[1]
        wet = {"value": rw["value"], "unit": UNIT, "ms_per_step": rw["ms"] / Ks, "steps": Ks, "wet_fraction": 1.0,
               "links_with_surface_storage_after_run": int(still_wet_all), "links": ns * world,
               "attempts_per_accepted": rw["attempts"] / max(rw["acc"], 1.0), "link_status_after_run": rw["state"],
               "roofline": {"bound": "fp64", **roof_w},
               "note": "y0 with 0.01 m of surface storage on every link: h_surf decays like t^-3/2 and never reaches 0, so every "
                       "rhs of every attempt runs pow(h_surf, 2/3) as the reference does unconditionally (model_204.hpp:96-104) — "
                       "8 pow per attempt instead of 1.  The nominal flop count (pow = 1 flop) does not see that work, hence the "
                       "lower nominal fraction at a similar FP64-pipe load"}
        rw["final"] = None

    # ---------------- strong scaling of configs[3]: the same 10 M links split over the ranks ----------------
    strong = None
    if subs and world > 1 and uid == 204:
        lo, hi = shard_range(ns, world, rank)
        n_s = hi - lo
        solver.upload_spatial_params(sp[:n_s])
        solver.set_forcing_columns(col[:n_s])
        rs_ = resident_steps(env, solver, stream, uid, y0[:n_s], Ks, W)
        solver.upload_spatial_params(sp)
        solver.set_forcing_columns(col)
        strong = {"value": rs_["value"], "unit": UNIT, "ms_per_step": rs_["ms"] / Ks, "steps": Ks, "links_total": ns,
                  "links_per_gpu": n_s, "scaling": "strong",
                  "note": f"{ns} links in all, rows split with the reference's chunk rule (main.cpp:275-307, sharding.shard_range); "
                          "the persistent grid (444 CTAs) has fewer tiles per CTA to balance, so the tail of a launch weighs more"}

    # ---------------- end-to-end arm: `e2e` ----------------
    e2e = e2e_sel = None
    if not args.no_e2e:
        e2e = e2e_steps(env, solver, stream, uid, ns, y0, K, W)
        if subs:
            e2e_sel = guarded(e2e_steps, env, solver, stream, uid, ns, y0, Ks, W, states=[2, 3], out_bits=32)

    # ---------------- baselines (rank 0, N == 1 only) ----------------
    cpu = ref_cuda = ref_cuda_wet = notebook = shim = None
    if rank == 0 and world == 1 and not args.no_baselines and uid == 204:
        cpu = guarded(cpu_baseline, "port", args, args.cpu_seconds)  # the oracle library always exists; report rather than hide a failure
        ref_cuda = guarded(reference_cuda_record, sp, col, pr, t2m, y0, ns, f"wet_fraction {args.wet_fraction}")
        if y0_wet is not None:
            ref_cuda_wet = guarded(reference_cuda_record, sp, col, pr, t2m, y0_wet, ns, "wet_fraction 1.0")

        def notebook_cpu():
            # the reference's CPU path as the north star names it: the notebook's SciPy integrator, one process per core
            from oracle import notebook_baseline as NB
            cores = os.cpu_count() or 1
            ns_nb = 24 * cores
            sp_n, col_n, pr_n, t2m_n, y0_n, tq_n = cpu_sample_inputs(ns_nb, args)
            _, steps_nb, dt_nb, procs = NB.run(sp_n, pr_n, t2m_n, col_n, y0_n, 0.0, DAY, tq_n[1:], processes=cores)
            return {"value": steps_nb / dt_nb, "unit": UNIT, "cores": procs, "kind": "reference",
                    "sample": f"scipy.integrate.solve_ivp(method='RK45', rtol=1e-6, atol=1e-9, t_eval=hourly) per link as "
                              f"model_dummy_python.ipynb:150-175,935-945 does, {ns_nb} links x 1 simulated day of the bench "
                              f"workload, {steps_nb} accepted steps in {dt_nb:.2f} s wall on {procs} processes"}
        notebook = guarded(notebook_cpu)
    if wet is not None and ref_cuda_wet is not None:
        wet["reference_cuda"] = ref_cuda_wet
    solver.close()
    del sp, col, pr, t2m, y0, y0_wet
    torch.cuda.set_stream(torch.cuda.default_stream(dev))

    # ---------------- other BASELINE configurations, as sub-records of the same line ----------------
    model200 = routed = None
    if subs and uid == 204:
        if rank == 0 and world == 1 and not args.no_e2e:
            shim = guarded(shim_record, args, ns, Ks, W)
        if world == 1:
            model200 = guarded(model200_record, env, args, Ks, W)
        # routed runs at every N: weak scaling, 2.5 M links per GPU, one exchange per coupling interval
        try:
            rr = routed_record(env, args, Ks, W, 2_500_000, with_e2e=False)
            if rr is not None:
                routed = {k: rr[k] for k in ("value", "unit", "ms_per_step", "steps", "exchanges_per_step", "partition_check", "accepted_steps_per_step",
                                             "attempts_per_accepted", "implicit_steps_total", "link_status_after_run", "gpu_launches")}
                routed["config"] = rr["config"]
                routed["kernel_share_of_step"] = rr["roofline"]["kernel_share_of_step"]
                routed["roofline_frac"] = rr["roofline"]["frac"]
                routed["kernel_ms_avg"] = rr["roofline"]["kernel_ms_avg"]
                routed["exchange"] = args.exchange if world > 1 else "none (1 GPU)"
        except Exception as ex:  # every rank takes part in the collectives: a failure here is reported, not hidden
            routed = {"error": repr(ex)[:400]}

    if rank == 0:
        hbm_peak, hbm_src = peaks()
        roof = roofline_of(main_r, uid, fp_peak, world)
        # algorithmic HBM bytes per link per window: state in+out, prepared parameters, forcing column,
        # dense records, counters
        bytes_per_link = (5 + 2) * 8 * 2 + 6 * 4 * 2 + 11 * 8 + 4 + 24 * 5 * 8
        hbm_gbs = bytes_per_link * ns / (roof["kernel_ms_avg"] * 1e-3) / 1e9
        static_ok = args.precision == 64 and args.wet_fraction == 0.0 and uid == 204
        line = {
            "metric": "accepted RK45 system-steps/sec", "value": value, "unit": UNIT,
            "n_gpus": world, "steps": K, "warmup": args.warmup, "warmup_done": W, "ms_per_step": main_r["ms"] / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64" if args.precision == 64 else "f32",
            "data": "synthetic", "config": workload_config(args, ns),
            "accepted_steps_per_step": main_r["acc"] / K, "attempts_per_accepted": main_r["attempts"] / max(main_r["acc"], 1.0),
            "link_status_after_run": main_r["state"],
            "e2e": e2e, "e2e_selected": e2e_sel, "e2e_shim": shim, "fp32": fp32, "wet": wet, "strong": strong,
            "model200": model200, "routed": routed, "gpu_launches": int(main_r["launches"]),
            "roofline": {"bound": "fp64" if args.precision == 64 else "fp32", **roof,
                         # STATIC: dram__bytes_read.sum + dram__bytes_write.sum of one committed ncu --set full capture of this
                         # command (10 M links, 24 queries), per link x links; goes stale when the kernel changes
                         "traffic": (DRAM_BYTES_PER_LINK_PER_LAUNCH * ns if static_ok else None),
                         "traffic_source": {"kind": "static", **PROFILE} if static_ok else None,
                         "peak_source": "measured live: register-resident FMA microbenchmark (hlm_measure_fma_peak); "
                                        "MEASURED_PEAKS.json holds no FP64/FP32 vector peak",
                         # instruction-level utilisation of the bounding pipe, from the same committed capture (static)
                         "fp64_pipe_utilization_ncu": ({"value": PROFILE["fp64_pipe_utilization"], "kind": "static",
                                                        "metric": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
                                                        "source": PROFILE["file"], "commit": PROFILE["commit"]} if static_ok else None),
                         "kernel": f"hlm::rk45_window_kernel<Model{uid},{'double' if args.precision == 64 else 'float'}>",
                         "hbm": {"achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak,
                                 "peak_source": hbm_src, "algorithmic_bytes_per_link_per_launch": bytes_per_link}},
            "cpu_baseline": cpu, "notebook_cpu": notebook, "reference_cuda": ref_cuda, "clocks": main_r["clocks"], "host_binding": numa,
        }
        print(json.dumps(line), flush=True)
    env.close()


if __name__ == "__main__":
    main()
