"""Small stand-alone workloads, one per kernel we want an `ncu --set full` capture of:

    python tools/ncu_targets.py <target> [reps]
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c 1 -o gpurun_out/<name> \
        python tools/ncu_targets.py <target>

targets: rbf (rbf_queries_kernel, N=4096 x 32768 queries), rbf_i8 (rbf_queries_i8_kernel), rff (rff_gemm_argmax_kernel,
config 4: F=1000, S=100, 2^20 grid points), append (append_update_kernel, k=4 onto N=4096), infogain
(info_gain_schur_kernel, 16384 paths x 4 points, N=900), rollout (the per-rollout device pass of the native tree search,
N=900 / N=4096), factor (potrf_diag_kernel + GEMMs, N=4096).  Each runs its op `reps` times (default 3).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from informative_path_planning_b200 import _lib as L
from informative_path_planning_b200 import aq_library as aq
from informative_path_planning_b200 import gpmodel_library as gp

R = (0., 10., 0., 10.)


def model(n, noise=0.5):
    rng = np.random.default_rng(0)
    X = rng.uniform(0, 10, (n, 2))
    z = 10 * np.sin(0.7 * X[:, :1]) * np.cos(0.5 * X[:, 1:]) + rng.normal(0, 0.7, (n, 1))
    m = gp.OnlineGPModel(R, 1.0, 100.0, noise=noise)
    m.add_data(X, z)
    return m, rng


def main():
    target = sys.argv[1]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    torch.cuda.set_device(0)
    if target in ('rbf', 'rbf_i8'):
        m, rng = model(4096)
        xq = L.to_device(rng.uniform(0, 10, (32768, 2)))
        for _ in range(reps):
            m.predict_device(xq, precision='fp64' if target == 'rbf' else 'i8x6')
    elif target == 'rff':
        r4 = np.random.default_rng(4)
        F_, S_ = 1000, 100
        W4 = r4.normal(size=(F_, 2)); b4 = 2 * np.pi * r4.uniform(size=(F_, 1)); T4 = r4.normal(size=(F_, S_))
        gx, gy = np.meshgrid(r4.uniform(1, 9, 1024), r4.uniform(1, 9, 1024))
        grid_d = L.to_device(np.vstack([gx.ravel(), gy.ravel()]).T)
        for _ in range(reps):
            aq.rff_eval_argmax_shared(W4, b4, T4, 100.0, grid_d)
    elif target == 'rffmesh':
        r4 = np.random.default_rng(4)
        F_, S_ = 1000, 100
        W4 = r4.normal(size=(F_, 2)); b4 = 2 * np.pi * r4.uniform(size=(F_, 1)); T4 = r4.normal(size=(F_, S_))
        xs_d, ys_d = L.to_device(r4.uniform(1, 9, 1024)), L.to_device(r4.uniform(1, 9, 1024))
        for _ in range(reps):
            aq.rff_eval_argmax_mesh(W4, b4, T4, 100.0, xs_d, ys_d)
    elif target == 'append':
        import copy
        m, rng = model(4096)
        for _ in range(reps):
            f = copy.copy(m)
            f.add_data(rng.uniform(1, 9, (4, 2)), rng.normal(0, 10, (4, 1)))
    elif target == 'infogain':
        m, rng = model(900)
        P, k = 16384, 4
        q = rng.uniform(0.5, 9.5, (P * k, 2))
        offs = np.arange(0, P * k + 1, k)
        for _ in range(reps):
            aq.info_gain_paths(m, q, offs)
    elif target in ('rollout', 'rollout4096'):
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import bench_mcts
        r = bench_mcts.run('gpu', 900 if target == 'rollout' else 4096, 'mean', budget=40)
        print(r['rollouts_per_s'], 'rollouts/s')
    elif target == 'factor':
        m, rng = model(4096)
        for _ in range(reps):
            m._factor(want_linv=False)
        import ctypes
        clk = np.zeros(4, dtype=np.int64)
        L.check(L.lib.ipp_potrf_debug_clocks(clk.ctypes.data_as(ctypes.c_void_p)))
        print('last diagonal block, SM clocks: load %d, pivots %d, inverse %d, store %d' % tuple(clk))
    else:
        raise SystemExit('unknown target ' + target)
    torch.cuda.synchronize()
    print('ok', target)


if __name__ == '__main__':
    main()
