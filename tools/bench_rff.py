"""Config 4 grid stage (F=1000 shared features, S=100 samples, 2^20-point grid): CUDA-event time of ipp_rff_eval_argmax."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from informative_path_planning_b200 import _lib as L, aq_library as aq
r4 = np.random.default_rng(4)
F_, S_ = 1000, int(sys.argv[1]) if len(sys.argv) > 1 else 100
W4 = r4.normal(size=(F_, 2)); b4 = 2 * np.pi * r4.uniform(size=(F_, 1)); T4 = r4.normal(size=(F_, S_))
gx, gy = np.meshgrid(r4.uniform(1, 9, 1024), r4.uniform(1, 9, 1024))
grid_d = L.to_device(np.vstack([gx.ravel(), gy.ravel()]).T)
for _ in range(3):
    aq.rff_eval_argmax_shared(W4, b4, T4, 100.0, grid_d)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); mx, am = aq.rff_eval_argmax_shared(W4, b4, T4, 100.0, grid_d); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
xs_d, ys_d = grid_d[:1024, 0].contiguous(), grid_d[::1024, 1].contiguous()
tm = []
for _ in range(12):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); mx2, am2 = aq.rff_eval_argmax_mesh(W4, b4, T4, 100.0, xs_d, ys_d); e1.record(); torch.cuda.synchronize()
    tm.append(e0.elapsed_time(e1))
print('  mesh form: %.3f ms (best of 10), same arg-max: %s, max |dmax| %.2e' % (min(tm[2:]), bool(torch.equal(am, am2)), float((mx - mx2).abs().max())))
ms = min(ts)
print('rff grid stage F=%d S=%d G=2^20: %.3f ms (best of 10; median %.3f) = %.1f TFLOP/s on 2 G F S' % (
    F_, S_, ms, sorted(ts)[5], 2.0 * (1 << 20) * F_ * S_ / ms / 1e9))
