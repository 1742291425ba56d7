"""tools/mcts_split.py for one belief size: python tools/mcts_split_n.py 4096 [mean|mes]"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_mcts
from informative_path_planning_b200 import _lib as L
n = int(sys.argv[1]); f_rew = sys.argv[2] if len(sys.argv) > 2 else 'mean'
bench_mcts.run('gpu', n, f_rew, budget=20)
r = min((bench_mcts.run('gpu', n, f_rew) for _ in range(3)), key=lambda x: x['seconds'])
w, d, lu = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
L.check(L.lib.ipp_tree_last_timing(ctypes.byref(w), ctypes.byref(d)))
L.check(L.lib.ipp_tree_last_launch_us(ctypes.byref(lu)))
clk = np.zeros(32, dtype=np.int64)
L.check(L.lib.ipp_rollout_debug_clocks(clk.ctypes.data_as(ctypes.c_void_p)))
print('N=%d %s: %.0f rollouts/s, step %.1f ms = walk %.1f + device %.1f (launch calls %.1f) ms' % (
    n, f_rew, r['rollouts_per_s'], r['seconds'] * 1e3, w.value / 1e3, d.value / 1e3, lu.value / 1e3))
print('  phases: tiles %d, hand-over %d, sum + C0 %d, conditioning %d' % tuple(clk[:4]))
print('  tile phase detail [4..10]:', [int(v) for v in clk[4:11]])
print('  conditioning detail: start-to-level clocks', [int(clk[17 + i] - clk[16]) for i in range(5)], 'scoring', int(clk[26] - clk[25]), 'levels-to-scoring', int(clk[25] - clk[16]))
