"""Day 0 of the unrouted Model 200 bench workload: wall time, links the RK45 path flags stiff, implicit steps.
usage: python tools/probe_model200_day0.py [links]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import tiger_hlm_gpu_b200 as hlm
from tiger_hlm_gpu_b200 import synthetic

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
sp = synthetic.make_spatial_params(ns)
col, ncells = synthetic.make_cells(ns)
pr, t2m = synthetic.make_forcing_grid(ncells, 3)
y0 = np.tile(np.array(synthetic.Y0_200), (ns, 1))
y0[:, 0] = np.random.default_rng(7).uniform(0.05, 5.0, ns)
tq = synthetic.hourly_queries(0.0, 1440.0)
for reject_limit in (5, 20):
    with hlm.Solver(0) as s:
        s.set_model_parameters(200, hlm.Parameters(initialStep=1e-6))
        s.set_stiff_fallback(True)
        s.set_max_attempts(5_000_000)
        s.set_reject_limit(reject_limit)
        s.upload_spatial_params(sp)
        s.upload_forcing(0, 1.0, pr)
        s.upload_forcing(1, 24.0, t2m)
        s.set_forcing_columns(col)
        s.set_output_states([0])
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.time()
            r = s.run_rk45(200, y0, 0.0, 1440.0, tq)
            torch.cuda.synchronize()
            dt = time.time() - t0
            codes, counts = np.unique(r["stiff"], return_counts=True)
            print(f"reject_limit {reject_limit} rep {rep}: {dt * 1e3:.1f} ms wall, stiff codes {dict(zip(codes.tolist(), counts.tolist()))}, "
                  f"accepted {int(r['n_accept'].sum())}, keys {[k for k in r if 'impl' in k or 'radau' in k]}", flush=True)
