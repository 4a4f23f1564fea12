import sys, os, zlib; sys.path.insert(0, os.getcwd())
import numpy as np
import tiger_hlm_gpu_b200 as hlm
s = hlm.Solver(0)
rng = np.random.default_rng(0)
os.makedirs('gpurun_out', exist_ok=True)
for yexp, lo, hi in ((0.2, -3, 12), (2.0/3.0, -8, 0)):
    x = 10 ** rng.uniform(lo, hi, 2_000_000)
    g = s.debug_eval(0, x, yexp)
    c = np.power(x, yexp)
    d = (g.view(np.int64) - c.view(np.int64))
    print('pow y=%.4f mismatch rate %.3e  max ulp %d' % (yexp, (d != 0).mean(), np.abs(d).max()))
    idx = np.nonzero(d)[0][:2000]
    np.savez_compressed('gpurun_out/pow_mismatch_%d.npz' % int(yexp*100), x=x[idx], g=g[idx], c=c[idx])
# also save a big sample of (x, libdevice pow) for offline emulation checks
x = 10 ** rng.uniform(-3, 12, 200000)
np.savez_compressed('gpurun_out/pow_sample_02.npz', x=x, g=s.debug_eval(0, x, 0.2))
x = 10 ** rng.uniform(-8, 0, 200000)
np.savez_compressed('gpurun_out/pow_sample_23.npz', x=x, g=s.debug_eval(0, x, 2.0/3.0))
# MUFU.RCP64H over all 2^20 high-mantissa patterns in [1,2), plus low-word sensitivity
hi = (np.arange(1 << 20, dtype=np.uint64) | np.uint64(0x3ff00000)) << np.uint64(32)
x = hi.view(np.float64)
r = s.debug_eval(1, x)
x2 = (hi | np.uint64(0xffffffff)).view(np.float64)
r2 = s.debug_eval(1, x2)
print('rcp64h low-word independent:', np.array_equal(r, r2), ' low 32 bits of result all zero:', not (r.view(np.uint64) & np.uint64(0xffffffff)).any())
rh = (r.view(np.uint64) >> np.uint64(32)).astype(np.uint32)
open('gpurun_out/rcp64h_1_2.bin', 'wb').write(zlib.compress(rh.tobytes(), 9))
print('table bytes compressed', os.path.getsize('gpurun_out/rcp64h_1_2.bin'))
# other exponents: is the mantissa result exponent-independent?
for e in (0x400, 0x3fe, 0x432, 0x3c0):
    hi_e = (np.arange(0, 1 << 20, 997, dtype=np.uint64) | np.uint64(e << 20)) << np.uint64(32)
    re = s.debug_eval(1, hi_e.view(np.float64))
    base = rh[::997].astype(np.int64)
    got = (re.view(np.uint64) >> np.uint64(32)).astype(np.int64)
    print('exp %x: mantissa-part equal:' % e, np.array_equal(got & 0xfffff, base & 0xfffff), ' exp delta', set(((got >> 20) - (base >> 20)).tolist()))

# ---- which operation makes GPU and CPU oracle differ? ------------------------------------------
from oracle import oracle as O
tr, r = O.trace(0, O.Params.make(), np.ones((1, 5)), 0.0, 5.0)
x = 1.0 / (tr[:, 2] + 1e-16)
g = s.debug_eval(0, x, 0.2)
c = np.power(x, 0.2)
print('dummy trace: attempts', len(tr), 'controller pow mismatches', int((g != c).sum()))
print(np.c_[tr[:, :3], x, (g.view(np.int64) - c.view(np.int64))][:15])
s.set_model_parameters(0, hlm.Parameters())
s.set_max_attempts(100000)
tq = np.linspace(0, 5, 2001)[1:]
gg = s.run_rk45(0, np.ones((1, 5)), 0.0, 5.0, tq)
oo = O.run_rk45(0, O.Params.make(), np.ones((1, 5)), 0.0, 5.0, tq)
bad = np.nonzero((gg['dense'][0] != oo['dense'][0]).any(axis=1))[0]
print('dummy: first differing query', bad[:3], 'at t', tq[bad[:3]], 'step ends', (tr[:, 0] + tr[:, 1])[:6])
print('gpu final', gg['final'][0].tolist()); print('cpu final', oo['final'][0].tolist())
print('diff ulps first bad row', (gg['dense'][0][bad[0]].view(np.int64) - oo['dense'][0][bad[0]].view(np.int64)) if len(bad) else None)
np.savez_compressed('gpurun_out/dummy_debug.npz', trace=tr, gd=gg['dense'], od=oo['dense'], tq=tq)
