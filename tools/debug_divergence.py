"""Locate the first step at which the CUDA path and the CPU oracle differ (DummyModel)."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import tiger_hlm_gpu_b200 as hlm
from oracle import oracle as O

s = hlm.Solver(0)
s.set_model_parameters(0, hlm.Parameters())
s.set_max_attempts(100000)
y0 = np.ones((1, 5))
tr, r = O.trace(0, O.Params.make(), y0, 0.0, 5.0)
# oracle post-step states
y = y0[0].copy()
states = []
for (t, h, err, acc) in tr:
    yo, e, k = O.step(0, None, 0, y, h, 1e-6, 1e-9, 0.0, 0.0)
    if acc:
        y = yo
        states.append((t + h, y.copy(), err, k.copy(), h))
tq = np.linspace(0, 5, 4001)[1:]
s.solve_begin(0, y0, 0.0, 5.0, tq)
seen = set()
for q in range(1, len(tq) + 1):
    s.solve_window(q, False)
    t, h, yy = s.solve_peek()
    key = float(t[0])
    if key in seen:
        continue
    seen.add(key)
    m = [st for st in states if st[0] == key]
    if not m:
        print('GPU t=%r not an oracle step end; h=%r' % (key, float(h[0])))
        near = min(states, key=lambda st: abs(st[0] - key))
        print('   nearest oracle t=%r diff %g' % (near[0], near[0] - key))
        continue
    st = m[0]
    i = states.index(st)
    h_next = tr[np.nonzero(tr[:, 0] == key)[0][0], 1] if (tr[:, 0] == key).any() else None
    dy = (yy[0].view(np.int64) - st[1].view(np.int64)).tolist()
    print('step %2d t=%.6f  y ulp diff %s   h_gpu=%r h_oracle_next=%r' % (i, key, dy, float(h[0]), h_next))
s.solve_end()
