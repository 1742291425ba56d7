"""Tensor-mode (3xTF32 or sliced INT8, env MODE=tf32x3|i8x6) vs FP64 predict: variance error statistics and throughput
by N (BASELINE config 3 sweep)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import informative_path_planning_b200 as pkg
from informative_path_planning_b200 import _lib as L
R = (0., 10., 0., 10.)
out = []
KC = int(os.environ.get('KC', '1'))
MODE = os.environ.get('MODE', 'tf32x3')
for N in [int(a) for a in (sys.argv[1:] or ['333', '1024', '2048', '4096', '8192'])]:
    rng = np.random.default_rng(N)
    X = rng.uniform(0, 10, (N, 2))
    z = 10 * np.sin(0.7 * X[:, :1]) * np.cos(0.5 * X[:, 1:]) + rng.normal(0, np.sqrt(0.5), (N, 1))
    m = pkg.OnlineGPModel(R, 1.0, 100.0, noise=0.5)
    m.add_data(X, z)
    m.tf32_acc_kblocks = KC
    M = int(os.environ.get('M', str(1 << 20)))
    q = L.to_device(rng.uniform(0, 10, (M, 2)))
    res = {'N': N, 'M': M, 'acc_kblocks': KC}
    for prec in ('fp64', MODE):
        mu, var = m.predict_device(q, precision=prec)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            mu, var = m.predict_device(q, precision=prec)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        res[prec + '_ms'] = ms
        res[prec + '_queries_per_s'] = M / ms * 1e3
        res[prec] = (mu.cpu().numpy(), var.cpu().numpy())
    (mu64, v64), (mu32, v32) = res.pop('fp64'), res.pop(MODE)
    err = np.abs(v32 - v64)
    res.update(mean_max_abs_diff=float(np.abs(mu32 - mu64).max()), var_abs_err_max=float(err.max()), var_abs_err_median=float(np.median(err)),
               var_rel_err_max=float((err / v64).max()), var_rel_err_median=float(np.median(err / v64)),
               var_err_over_prior_max=float(err.max() / 100.0), bias=float(np.mean(v32 - v64)), speedup=res['fp64_ms'] / res[MODE + '_ms'])
    T = (N + 127) // 128
    T64 = (N + 63) // 64
    res['mode'] = MODE
    res['tf32_executed_tflops'] = 3 * 2.0 * 128 * 128 * T * (T + 1) / 2 * M / (res[MODE + '_ms'] / 1e3) / 1e12 if MODE == 'tf32x3' else None
    if MODE in ('i8x7', 'i8x6', 'i8x5'):
        ns, bn = {'i8x7': (7, 64), 'i8x6': (6, 64), 'i8x5': (5, 96)}[MODE]
        kblocks = sum(-(-min((I + 1) * bn, N) // 64) for I in range(-(-N // bn)))
        res['i8_executed_tops'] = ns * (ns + 1) / 2 * 2.0 * 128 * bn * 64 * kblocks * (M / 128) / (res[MODE + '_ms'] / 1e3) / 1e12
    else:
        res['i8_executed_tops'] = None
    print(json.dumps(res)); out.append(res)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/%s_sweep.json' % MODE, 'w'), indent=1)
