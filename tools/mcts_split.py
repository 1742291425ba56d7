"""Wall-time split of the native tree search: tree walk on the host vs device pass (ipp_tree_last_timing)."""
import ctypes, os, sys
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
