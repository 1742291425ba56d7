"""cProfile of one planning step of the device-backed tree search (where does the host time go?)."""
import cProfile, pstats, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_mcts
n_obs = int(sys.argv[1]) if len(sys.argv) > 1 else 900
f_rew = sys.argv[2] if len(sys.argv) > 2 else 'mean'
bench_mcts.run('gpu', n_obs, f_rew, budget=20)
pr = cProfile.Profile()
pr.enable()
r = bench_mcts.run('gpu', n_obs, f_rew)
pr.disable()
print(r['rollouts_per_s'], 'rollouts/s')
pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
