"""Tuning aid (variant build with -DMS_PHOT_STAMPS): per-pair clock64 stamps of CTA 0 of the fused photometry kernel."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from minesweeper_b200 import synth
from minesweeper_b200.likelihood import likelihood
import bench
class A: pass
a = A(); a.small_grid = False; a.hidden = 128; a.batch = 1 << 20
fitargs, fitpars, runbools, box = bench.make_workload(a)
L = likelihood(fitargs, fitpars, runbools, precision='tc', device=0, verbose=False)
eng = L.engine
th = torch.from_numpy(synth.draw_theta_phot(a.batch, seed=3, box=box)).cuda()
out = torch.empty(a.batch, dtype=torch.float64, device='cuda')
for _ in range(3):
    eng.lnlike_batch(th, L._colmap, L._fixed, L._flags, want_derived=False, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.lnlike_batch(th, L._colmap, L._fixed, L._flags, want_derived=False, out=out); e1.record()
torch.cuda.synchronize()
print('step ms', e0.elapsed_time(e1))
buf = (C.c_longlong * 8192)()
eng.lib.ms_debug_read_stamps.restype = C.c_int
eng.lib.ms_debug_read_stamps.argtypes = [C.c_void_p, C.c_int]
print('rc', eng.lib.ms_debug_read_stamps(buf, 8192))
s = np.frombuffer(buf, dtype=np.int64).reshape(-1, 8)
t0 = s[0, 0]
npairs = 28
print('pair: pro_wait_start pro_start pro_end | epi_top epi_xfull_passed epi_end   (clk since kernel start of CTA 0)')
for i in range(npairs):
    r = s[i, :6] - t0
    print(i, r[:3], '|', r[3:6], ' pro_dur', r[2] - r[1], 'epi_dur', r[5] - r[3], 'epi_wait_x', r[4] - r[3])

cb = (C.c_longlong * (148 * 8))()
eng.lib.ms_debug_read_cta.restype = C.c_int
eng.lib.ms_debug_read_cta.argtypes = [C.c_void_p]
print('rc', eng.lib.ms_debug_read_cta(cb))
c = np.frombuffer(cb, dtype=np.int64).reshape(148, 8)
clk = c[:, 1] - c[:, 0]; ns = c[:, 3] - c[:, 2]
print('per-CTA clk: min %d median %d max %d' % (clk.min(), np.median(clk), clk.max()))
print('per-CTA ns : min %d median %d max %d' % (ns.min(), np.median(ns), ns.max()))
print('clock rate GHz (clk/ns): median %.3f min %.3f max %.3f' % (np.median(clk / ns), (clk / ns).min(), (clk / ns).max()))
print('kernel span ns (first start .. last end):', c[:, 3].max() - c[:, 2].min(), ' start spread ns', c[:, 2].max() - c[:, 2].min())
o = np.argsort(c[:, 3])
print('last 5 CTAs to finish (cta, sm, ns, clk):', [(int(i), int(c[i, 4]), int(ns[i]), int(clk[i])) for i in o[-5:]])
print('first 5 CTAs to finish:', [(int(i), int(c[i, 4]), int(ns[i]), int(clk[i])) for i in o[:5]])
