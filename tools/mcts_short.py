import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_mcts
r = bench_mcts.run('gpu', int(sys.argv[1]) if len(sys.argv) > 1 else 900, 'mean', budget=40)
print(r['rollouts_per_s'])
