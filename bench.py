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

DRAM_BYTES_PER_LINK_PER_LAUNCH = 1396.0  # ncu --set full of `python bench.py`: (2.642 GB read + 11.318 GB write) / 10 M links, profiles/r1h_*
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
    ap.add_argument("--schedule", default="auto", choices=["auto", "tiles", "lanes"],
                    help="how links are dealt to lanes (hlm_set_schedule); auto = lanes for routed runs, tiles otherwise")
    ap.add_argument("--days", type=int, default=365, help="length of the forcing record / run horizon")
    ap.add_argument("--wet-fraction", type=float, default=0.0,
                    help="share of links started with surface storage so Model204's pow() branch runs")
    ap.add_argument("--precision", type=int, default=64, choices=[64, 32])
    ap.add_argument("--rtol", type=float, default=None, help="override rtol (default 1e-6)")
    ap.add_argument("--atol", type=float, default=None, help="override atol (default 1e-9)")
    ap.add_argument("--stiff-fallback", action="store_true",
                    help="continue links the RK45 path flags stiff with the Radau IIA fallback (always on for model200/routed)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-baselines", action="store_true", help="skip cpu_baseline / reference_cuda legs")
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
        "config": dict(workload_config(args, args.links_per_gpu), sample_links_per_step=ns_sample),
        "cpu_baseline": {"value": value, "unit": "accepted system-steps/s", "cores": cores, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": "accepted system-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
W_MIN_FLOP_PER_ATTEMPT_200 = 585.0  # as W_min with Model 200's rhs (36 nominal flop instead of 32), DESIGN.md


def run_routed_arm(args):
    """BASELINE configs[4]: Model 200 on a synthetic river network, partitioned by sub-basin over the ranks,
    boundary discharge exchanged with one NCCL all-gather per coupling interval.  One step = one simulated
    hour (60 / couple_minutes intervals).  Weak scaling: links_per_gpu links per rank."""
    import torch
    import tiger_hlm_gpu_b200 as hlm
    from tiger_hlm_gpu_b200 import routing, synthetic
    from tiger_hlm_gpu_b200.sharding import reduce_timing

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    K, W = args.steps, max(args.warmup, 3)  # never fewer than 3 untimed steps before a timed region (timing rules)
    ns_all = args.links_per_gpu * world
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
    solver = hlm.Solver(local_rank)
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

    def hour(k, first=False):
        for i in range(n_int):
            tf = 60.0 * k + dt * (i + 1)
            tq = np.array([tf])
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
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for k in range(W, W + K):
        hour(k, first=(k == 0))
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    kern_ms, kern_n = solver.kernel_time_ms()
    launches = solver.launch_count() - launches0
    tot1 = solver.solve_totals()
    radau = int(solver.solve_radau_steps().sum())
    acc = tot1["n_accept"] - tot0["n_accept"]
    attempts = acc + (tot1["n_reject"] - tot0["n_reject"]) + (tot1["n_jump"] - tot0["n_jump"])
    state = dict(tot1)

    # e2e: the same intervals, with the host in the loop as a caller writing hourly output has it: the hour's
    # last dense record (discharge + stores of every link) is copied to pinned host memory every step, and the
    # first step uploads y0.  (Between intervals nothing else crosses PCIe: state stays resident.)
    e2e = None
    if not args.no_e2e:
        win = torch.zeros((sel.size, 1, 5), dtype=torch.float64).pin_memory()
        barrier()
        t_start = time.perf_counter()
        acc_before = solver.solve_totals()["n_accept"]
        for k in range(W + K, W + 2 * K):
            hour(k)
            solver.solve_wait_copy(solver.solve_fetch_window_packed(win.numpy()))
        barrier()
        e_ms = (time.perf_counter() - t_start) * 1e3
        acc_e = solver.solve_totals()["n_accept"] - acc_before
        e_ms_max, (acc_e_all,) = reduce_timing(e_ms, [acc_e], dist, dev)
        e2e = {"value": acc_e_all / (e_ms_max * 1e-3), "unit": "accepted system-steps/s", "h2d_bytes_per_step": 8 * n_int,
               "d2h_bytes_per_step": int(sel.size) * 40 + 56, "ms_per_step": e_ms_max / K,
               "api": "routing.RoutedSolver over the C ABI (hlm_route_gather + hlm_solve_advance + hlm_solve_window + "
                      "hlm_solve_fetch_window_packed), pinned host buffer; state resident between intervals"}
    rs.end()

    ms_max, (acc_all, att_all, launches_all, kern_ms_all, kern_n_all, radau_all) = reduce_timing(
        ms, [acc, attempts, launches, kern_ms, kern_n, radau], dist, dev)
    if rank == 0:
        kern_avg_ms = kern_ms_all / max(kern_n_all, 1)
        att_per_launch = att_all / max(kern_n_all, 1)
        achieved = W_MIN_FLOP_PER_ATTEMPT_200 * att_per_launch / (kern_avg_ms * 1e-3) / 1e12
        line = {
            "metric": "accepted RK45 system-steps/sec", "value": acc_all / (ms_max * 1e-3), "unit": "accepted system-steps/s",
            "n_gpus": world, "steps": K, "warmup": args.warmup, "warmup_done": W, "ms_per_step": ms_max / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "routed sub-basin network (BASELINE configs[4]): Model 200 (project-defined), synthetic "
                                   "river network, links coupled through upstream discharge held over a coupling interval; "
                                   "one step = one simulated hour",
                       "links_per_gpu": args.links_per_gpu, "links_total": ns_all, "couple_minutes": dt,
                       "intervals_per_step": n_int, "schedule": args.schedule, "sub_basins": plan.n_subbasins, "cut_edges": plan.n_cut_edges,
                       "halo_doubles": plan.halo_len, "rtol": PRM6[1], "atol": PRM6[2],
                       "parallelism": (f"sub-basins dealt to {world} GPU(s); " +
                                       ("boundary discharge stored into peer memory by the kernels, one barrier per interval" if rs.peer
                                        else "one NCCL all-gather of the boundary vector per interval"))
                                      if world > 1 else "1 GPU, no exchange",
                       "l2": "inputs larger than L2 (state + parameters of the rank's links >> 126 MB)"},
            "accepted_steps_per_step": acc_all / K, "attempts_per_accepted": att_all / max(acc_all, 1.0),
            "implicit_steps_total": radau_all, "exchanges_per_step": (rs.exchanges - ex0) / max(1, 2 * K if e2e else K),
            "link_status_after_run": {k: state[k] for k in ("active", "done", "stiff", "stalled")},
            "e2e": e2e, "gpu_launches": int(launches_all),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp_peak, "unit": "TFLOP/s", "frac": achieved / fp_peak,
                         "traffic": None, "flop_per_attempt": W_MIN_FLOP_PER_ATTEMPT_200, "attempts_per_launch": att_per_launch,
                         "kernel_ms_avg": kern_avg_ms, "kernel": "hlm::rk45_window_kernel<Model200,double>",
                         "kernel_share_of_step": kern_ms_all / max(world, 1) / ms_max,
                         "peak_source": "measured live (hlm_measure_fma_peak)"},
            "cpu_baseline": None, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    solver.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


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

    import torch
    import tiger_hlm_gpu_b200 as hlm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    ns = args.links_per_gpu
    K, W = args.steps, max(args.warmup, 3)  # never fewer than 3 untimed steps before a timed region (timing rules)
    assert W + K <= args.days, "not enough forcing days for warmup+steps"
    from tiger_hlm_gpu_b200.sharding import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local_rank)  # before any pinned allocation

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
    solver.set_max_attempts(5_000_000)
    solver.upload_spatial_params(sp)
    solver.upload_forcing(0, 1.0, pr)
    solver.upload_forcing(1, 24.0, t2m)
    solver.set_forcing_columns(col)
    fp_peak = solver.measure_fma_peak(args.precision)

    # ---------------- resident arm: `value` ----------------
    # One step = one simulated day: hlm_solve_restart (new interval from the resident final state,
    # what a chained run_rk45 does) + hlm_solve_window (the hot kernel).  The run is driven in
    # day-sized intervals because the reference's stiffness threshold scales with (tf - t0).
    def day_queries(k):
        return k * DAY + 60.0 * np.arange(1, 25)

    solver.solve_begin(uid, y0, 0.0, DAY, day_queries(0))
    for k in range(W):
        if k:
            solver.solve_restart(k * DAY, (k + 1) * DAY, day_queries(k))
        solver.solve_window(24, True)
    solver.synchronize()
    tot0 = solver.solve_totals()
    solver.kernel_time_ms()  # drop warm-up kernel timings
    launches0 = solver.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for k in range(W, W + K):
        if k:
            solver.solve_restart(k * DAY, (k + 1) * DAY, day_queries(k))
        solver.solve_window(24, True)
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    kern_ms, kern_n = solver.kernel_time_ms()
    launches = solver.launch_count() - launches0
    tot1 = solver.solve_totals()
    acc = tot1["n_accept"] - tot0["n_accept"]
    attempts = acc + (tot1["n_reject"] - tot0["n_reject"]) + (tot1["n_jump"] - tot0["n_jump"])
    state = dict(tot1)
    solver.solve_end()

    from tiger_hlm_gpu_b200.sharding import reduce_timing
    ms_max, (acc_all, att_all, launches_all, kern_ms_all, kern_n_all) = reduce_timing(
        ms, [acc, attempts, launches, kern_ms, kern_n], dist, dev)
    value = acc_all / (ms_max * 1e-3)

    # ---------------- FP32 arm (BASELINE configs[3]: "FP32 vs FP64"): same days, same links, state and stages in FP32 ----
    fp32 = None
    if args.precision == 64 and not args.no_baselines:
        solver.set_precision(32)
        peak32 = solver.measure_fma_peak(32)
        solver.solve_begin(uid, y0, 0.0, DAY, day_queries(0))
        for k in range(W):
            if k:
                solver.solve_restart(k * DAY, (k + 1) * DAY, day_queries(k))
            solver.solve_window(24, True)
        solver.synchronize()
        t32a = solver.solve_totals()
        solver.kernel_time_ms()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        f0.record(stream)
        for k in range(W, W + K):
            solver.solve_restart(k * DAY, (k + 1) * DAY, day_queries(k))
            solver.solve_window(24, True)
        f1.record(stream)
        barrier()
        ms32 = f0.elapsed_time(f1)
        t32b = solver.solve_totals()
        solver.solve_end()
        solver.set_precision(64)
        acc32 = t32b["n_accept"] - t32a["n_accept"]
        att32 = acc32 + (t32b["n_reject"] - t32a["n_reject"]) + (t32b["n_jump"] - t32a["n_jump"])
        ms32_max, (acc32_all, att32_all) = reduce_timing(ms32, [acc32, att32], dist, dev)
        fp32 = {"value": acc32_all / (ms32_max * 1e-3), "unit": "accepted system-steps/s", "ms_per_step": ms32_max / K,
                "attempts_per_accepted": att32_all / max(acc32_all, 1.0),
                "roofline_frac": w_min_for(uid) * att32_all / world / (ms32_max * 1e-3) / 1e12 / peak32, "peak_tflops": peak32,
                "link_status_after_run": {k: t32b[k] for k in ("active", "done", "stiff", "stalled")},
                "note": "hlm_set_precision(32): FP32 state, stages and error control; no reference counterpart (the reference is "
                        "FP64 only), compared with FP64 at rtol 1e-4 in tests/test_gpu_parity.py"}

    # ---------------- end-to-end arm: `e2e` ----------------
    e2e = None
    if not args.no_e2e:
        nq_w = 24
        y_host = torch.from_numpy(y0.copy()).pin_memory()
        f_host = torch.zeros((ns, 5), dtype=torch.float64).pin_memory()
        d_host = torch.zeros((ns, nq_w, 5), dtype=torch.float64).pin_memory()
        stiff = torch.zeros(ns, dtype=torch.int32).pin_memory()
        na = torch.zeros(ns, dtype=torch.int64).pin_memory()
        lib = hlm.load_library()
        import ctypes as C

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

        for k in range(W):
            one_step(k)
        barrier()
        t_start = time.perf_counter()
        ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ee0.record(stream)
        acc_e = 0
        for k in range(W, W + K):
            acc_e += one_step(k)
        ee1.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t_start) * 1e3
        # the call is synchronous at its end (results are in host memory), so wall clock and the event
        # pair bracket the same work; report the larger
        e_ms = max(ee0.elapsed_time(ee1), wall_ms)
        e_ms_max, (acc_e_all,) = reduce_timing(e_ms, [acc_e], dist, dev)
        h2d = ns * 5 * 8 + nq_w * 8
        d2h = ns * 5 * 8 + ns * nq_w * 5 * 8 + ns * 4 + ns * 4
        e2e = {"value": acc_e_all / (e_ms_max * 1e-3), "unit": "accepted system-steps/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e_ms_max / K,
               "api": f"hlm_run_rk45 (C ABI under rk45_api::run_rk45<Model{uid}>), pinned host buffers"}

    # ---------------- baselines (rank 0, N == 1 only) ----------------
    cpu = None
    ref_cuda = None
    if rank == 0 and world == 1 and not args.no_baselines and uid == 204:
        try:
            cpu = cpu_baseline("port", args, args.cpu_seconds)
        except Exception as ex:  # the oracle library always exists; report rather than hide a failure
            cpu = {"error": repr(ex)}
        try:
            from tests import refs
            from tiger_hlm_gpu_b200 import synthetic
            if refs.have("libref_cuda.so"):
                ns_r = min(ns, 1 << 20)
                blocks = [synthetic.expand_forcing_per_link(pr[:48], col[:ns_r]),
                          synthetic.expand_forcing_per_link(t2m[:2], col[:ns_r])]
                tq_r = 60.0 * np.arange(0, 25)
                refs.ref_cuda_run204(PRM6, y0[:ns_r], 0.0, DAY, tq_r, sp[:ns_r], blocks, [1.0, 24.0], counted=False)
                r = refs.ref_cuda_run204(PRM6, y0[:ns_r], 0.0, DAY, tq_r, sp[:ns_r], blocks, [1.0, 24.0], counted=True)
                p = refs.ref_cuda_run204(PRM6, y0[:ns_r], 0.0, DAY, tq_r, sp[:ns_r], blocks, [1.0, 24.0], counted=False)
                ref_cuda = {"value": float(r["n_accept"].sum()) / (p["kernel_ms"] * 1e-3),
                            "unit": "accepted system-steps/s", "kernel_ms": p["kernel_ms"],
                            "sample": f"unchanged reference kernel rk45_then_radau_multi<Model204> built for sm_100a, "
                                      f"{ns_r} links x day 0 of the bench workload, 25 hourly queries, 1-D launch "
                                      f"<<<ceil(ns/128),128>>>, per-link expanded forcing; kernel time only"}
        except Exception as ex:
            ref_cuda = {"error": repr(ex)}

    notebook = None
    if rank == 0 and world == 1 and not args.no_baselines and uid == 204:
        # the reference's CPU path as the north star names it: the notebook's SciPy integrator, one process per core
        try:
            from oracle import notebook_baseline as NB
            cores = os.cpu_count() or 1
            ns_nb = 24 * cores
            sp_n, col_n, pr_n, t2m_n, y0_n, tq_n = cpu_sample_inputs(ns_nb, args)
            _, steps_nb, dt_nb, procs = NB.run(sp_n, pr_n, t2m_n, col_n, y0_n, 0.0, DAY, tq_n[1:], processes=cores)
            notebook = {"value": steps_nb / dt_nb, "unit": "accepted system-steps/s", "cores": procs, "kind": "reference",
                        "sample": f"scipy.integrate.solve_ivp(method='RK45', rtol=1e-6, atol=1e-9, t_eval=hourly) per link as "
                                  f"model_dummy_python.ipynb:150-175,935-945 does, {ns_nb} links x 1 simulated day of the bench "
                                  f"workload, {steps_nb} accepted steps in {dt_nb:.2f} s wall on {procs} processes"}
        except Exception as ex:
            notebook = {"error": repr(ex)}

    if rank == 0:
        hbm_peak, hbm_src = peaks()
        kern_avg_ms = kern_ms_all / max(kern_n_all, 1)
        att_per_launch = att_all / max(kern_n_all, 1)
        w_min = W_MIN_FLOP_PER_ATTEMPT if uid == 204 else W_MIN_FLOP_PER_ATTEMPT_200
        achieved_tflops = w_min * att_per_launch / (kern_avg_ms * 1e-3) / 1e12
        # algorithmic HBM bytes per link per window: state in+out, prepared parameters, forcing column,
        # dense records, counters
        bytes_per_link = (5 + 2) * 8 * 2 + 6 * 4 * 2 + 11 * 8 + 4 + 24 * 5 * 8
        hbm_gbs = bytes_per_link * ns / (kern_avg_ms * 1e-3) / 1e9
        line = {
            "metric": "accepted RK45 system-steps/sec", "value": value, "unit": "accepted system-steps/s",
            "n_gpus": world, "steps": K, "warmup": args.warmup, "warmup_done": W, "ms_per_step": ms_max / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64" if args.precision == 64 else "f32",
            "data": "synthetic", "config": workload_config(args, ns),
            "accepted_steps_per_step": acc_all / K, "attempts_per_accepted": att_all / max(acc_all, 1.0),
            "link_status_after_run": {k: state[k] for k in ("active", "done", "stiff", "stalled")},
            "e2e": e2e, "fp32": fp32, "gpu_launches": int(launches_all),
            "roofline": {"bound": "fp64" if args.precision == 64 else "fp32",
                         "achieved": achieved_tflops, "peak": fp_peak, "unit": "TFLOP/s",
                         "frac": achieved_tflops / fp_peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel
                         # (profiles/r1h_ncu_full_window_kernel_10M.csv: this command, 10 M links, 24 queries), per link
                         "traffic": (DRAM_BYTES_PER_LINK_PER_LAUNCH * ns if args.precision == 64 and args.wet_fraction == 0.0 and uid == 204 else None),
                         "traffic_source": "ncu --set full capture of this command (profiles/r1h_ncu_full_window_kernel_10M.csv), per link x links",
                         "peak_source": "measured live: register-resident FMA microbenchmark (hlm_measure_fma_peak); "
                                        "MEASURED_PEAKS.json holds no FP64/FP32 vector peak",
                         # instruction-level utilisation of the bounding pipe, from the committed ncu capture of this
                         # command (not measured live: ncu counters are not available inside a timed run)
                         "fp64_pipe_utilization_ncu": ({"value": 0.615, "metric": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
                                                        "source": "profiles/r1h_ncu_full_window_kernel_10M.csv"}
                                                       if uid == 204 and args.precision == 64 else None),
                         "flop_per_attempt": w_min, "attempts_per_launch": att_per_launch,
                         "kernel_ms_avg": kern_avg_ms, "kernel": f"hlm::rk45_window_kernel<Model{uid},{'double' if args.precision == 64 else 'float'}>",
                         "hbm": {"achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak,
                                 "peak_source": hbm_src, "algorithmic_bytes_per_link_per_launch": bytes_per_link}},
            "cpu_baseline": cpu, "notebook_cpu": notebook, "reference_cuda": ref_cuda, "clocks": clocks, "host_binding": numa,
        }
        print(json.dumps(line), flush=True)
    solver.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
