"""Wall-time split of the native tree search: tree walk on the host vs device pass (ipp_tree_last_timing)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_mcts
from informative_path_planning_b200 import _lib as L
for n in (100, 900, 4096):
    for f_rew in ('mean', 'mes'):
        bench_mcts.run('gpu', n, f_rew, budget=20)
        r = min((bench_mcts.run('gpu', n, f_rew) for _ in range(3)), key=lambda x: x['seconds'])
        w, d = ctypes.c_double(), ctypes.c_double()
        L.check(L.lib.ipp_tree_last_timing(ctypes.byref(w), ctypes.byref(d)))
        print('N=%d %s: %.0f rollouts/s, step %.1f ms = walk %.1f ms + device %.1f ms + rest %.1f ms' % (
            n, f_rew, r['rollouts_per_s'], r['seconds'] * 1e3, w.value / 1e3, d.value / 1e3,
            r['seconds'] * 1e3 - (w.value + d.value) / 1e3))
        lu = ctypes.c_double()
        L.check(L.lib.ipp_tree_last_launch_us(ctypes.byref(lu)))
        print('    of the device time, inside the launch calls: %.1f ms' % (lu.value / 1e3))
        clk = np.zeros(32, dtype=np.int64)
        L.check(L.lib.ipp_rollout_debug_clocks(clk.ctypes.data_as(ctypes.c_void_p)))
        print('    finishing CTA of the last rollout, SM clocks: tiles %d, hand-over %d, partial sum + C0 %d, conditioning %d' % tuple(clk[:4]))
        print('      tile phase: setup %d, W loads + column image %d, row image %d, barrier %d, mean + contraction %d, other tiles + fold %d, 8-warp sum + store %d' % tuple(clk[4:11]))
        lv = clk[17:25]
        print('      conditioning: levels at', [int(v - clk[16]) for v in lv if v > 0], 'scoring at', int(clk[25] - clk[16]), 'end at', int(clk[26] - clk[16]))
