import sys, os, tempfile
sys.path.insert(0, '/root/repo')
os.chdir(tempfile.mkdtemp())
import time
from informative_path_planning_b200 import envmodel_library as envlib, evaluation_library as evalib, robot_library as roblib
from informative_path_planning_b200 import obstacles as obslib
from informative_path_planning_b200 import mcts_library as mc
ranges = (0., 10., 0., 10.)
world = envlib.Environment(ranges=ranges, NUM_PTS=20, variance=100.0, lengthscale=1.0, noise=0.5, seed=0, dim=2)
robot = roblib.Robot(sample_world=world.sample_value, start_loc=(5.0, 5.0, 0.0), start_time=0, extent=ranges, dimension=2,
                     init_lengthscale=1.0, init_variance=100.0, noise=0.5, path_generator='dubins', frontier_size=15,
                     horizon_length=1.5, turning_radius=0.05, sample_step=0.5, f_rew='mes', nonmyopic=True, tree_type='dpw',
                     computation_budget=250, rollout_length=5, obstacle_world=obslib.FreeWorld(),
                     evaluation=evalib.Evaluation(world=world, reward_function='mes'))
t0 = time.perf_counter()
robot.planner(T=150)
dt = time.perf_counter() - t0
print('150-step mission with metrics: %.2f s, N=%d, native searches %d, final MSE %.3f, dist %.1f' % (
    dt, robot.GP.xvals.shape[0], mc.cMCTS.native_searches, robot.eval.metrics['MSE'][149], robot.dist))
robot.eval.plot_metrics()
print(sorted(os.listdir('figures/mes')))
