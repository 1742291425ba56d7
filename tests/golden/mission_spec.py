"""Shared description of the seeded golden missions (tests/golden/missions.npz): used by make_golden.py (which runs the
REFERENCE's Robot on them, in the build container) and by tests/test_gpu_mission.py (which runs this package's Robot on
the device).  No reference access here."""
import numpy as np

RANGES = (0., 10., 0., 10.)

MISSION_METRICS = ['aquisition_function', 'simple_regret', 'sample_regret_loc', 'sample_regret_val', 'max_loc_error',
                   'max_val_error', 'instant_regret', 'max_val_regret', 'mes_reward_robot', 'mes_reward_omni',
                   'info_gain_reward', 'MSE', 'hotspot_error', 'current_highest_obs', 'distance_traveled']


def mission_field(x):
    return (25.0 * np.exp(-0.5 * ((x[:, 0] - 7.0) ** 2 + (x[:, 1] - 6.5) ** 2) / 1.5 ** 2)
            + 8.0 * np.sin(0.8 * x[:, 0]) * np.cos(0.6 * x[:, 1]))[:, None]


MISSIONS = {   # name: (f_rew, nonmyopic, tree_type, path_generator, obstacle world, T, budget, rollout length)
    'myopic_mean_fan': ('mean', False, 'dpw', 'default', 'free', 8, 0, 0),
    'myopic_mean_dubins_blocks': ('mean', False, 'dpw', 'dubins', 'block', 8, 0, 0),
    'myopic_mes_fan': ('mes', False, 'dpw', 'default', 'free', 4, 0, 0),
    'nonmyopic_mean_dpw': ('mean', True, 'dpw', 'default', 'free', 4, 25, 3),
    'nonmyopic_mean_belief_dubins': ('mean', True, 'belief', 'dubins', 'channel', 3, 20, 3),
    # dpw tree over Dubins candidates: the configurations the native tree search (csrc/tree.cu) takes over
    'nonmyopic_mean_dpw_dubins': ('mean', True, 'dpw', 'dubins', 'free', 4, 60, 4),
    'nonmyopic_mes_dpw_dubins_blocks': ('mes', True, 'dpw', 'dubins', 'block', 3, 40, 3),
    'nonmyopic_ei_dpw_dubins_channel': ('exp_improve', True, 'dpw', 'dubins', 'channel', 3, 40, 3),
    # the remaining rewards: information gain (Schur form, batched), expected improvement, and the heuristic rewards of the
    # reference's current sweep (run_sim_all.sh:17), whose rollouts skip the belief updates here
    'myopic_info_gain_fan': ('info_gain', False, 'dpw', 'default', 'free', 4, 0, 0),
    'myopic_ei_dubins': ('exp_improve', False, 'dpw', 'dubins', 'free', 5, 0, 0),
    'myopic_naive_dubins': ('naive', False, 'dpw', 'dubins', 'free', 4, 0, 0),
    'nonmyopic_naive_dpw_dubins': ('naive', True, 'dpw', 'dubins', 'block', 3, 40, 3),
    'nonmyopic_naive_value_dpw_dubins': ('naive_value', True, 'dpw', 'dubins', 'free', 3, 40, 3),
    # information gain inside the tree search (robot_library.py:95-100 accepts it with nonmyopic=True): rollouts on the
    # information-gain belief, natively for the dpw tree; hotspot_info (= info_gain + UCB): two passes per rollout, natively too
    'nonmyopic_info_gain_dpw_dubins': ('info_gain', True, 'dpw', 'dubins', 'free', 3, 40, 3),
    'nonmyopic_hotspot_dpw_dubins': ('hotspot_info', True, 'dpw', 'dubins', 'block', 3, 30, 3),
}


def mission_prior():
    """A few scattered prior observations: without them the robot's first observations lie on one straight line, the
    candidate fan is mirror-symmetric about it and the best two paths tie to the last bit (an ill-posed arg-max)."""
    x = np.random.default_rng(5).uniform(1.0, 9.0, (6, 2))
    return x, mission_field(x)


def mission_kwargs(name, evaluation, obstacle_world, sample_world):
    f_rew, nonmyopic, tree_type, gen, _, T, budget, rl = MISSIONS[name]
    return dict(sample_world=sample_world, start_loc=(5.0, 5.0, 0.0), start_time=0, extent=RANGES, dimension=2,
                kernel_file=None, kernel_dataset=None, prior_dataset=mission_prior(), init_lengthscale=1.0, init_variance=100.0,
                noise=0.5, path_generator=gen, goal_only=False, frontier_size=15, horizon_length=1.5,
                turning_radius=0.05, sample_step=0.5, evaluation=evaluation, f_rew=f_rew, create_animation=False,
                learn_params=False, nonmyopic=nonmyopic, discretization=(20, 20), use_cost=False, MIN_COLOR=-25.,
                MAX_COLOR=25., computation_budget=budget, rollout_length=rl, obstacle_world=obstacle_world,
                tree_type=tree_type)


def mission_world(model_cls):
    import types
    g = np.linspace(0.5, 9.5, 10)
    gx, gy = np.meshgrid(g, g)
    Xw = np.vstack([gx.ravel(), gy.ravel()]).T
    wm = model_cls(RANGES, 1.0, 100.0, noise=0.5, dimension=2)
    wm.add_data(Xw, mission_field(Xw))
    return types.SimpleNamespace(dim=2, GP=wm, x1min=0., x1max=10., x2min=0., x2max=10.)


def mission_sample_world(x):
    return mission_field(x) + np.random.normal(loc=0, scale=np.sqrt(0.5))      # one scalar draw per batch (Q9)


def mission_obstacles(ob, kind):
    """ob: an obstacles module (the reference's, or this package's)."""
    ext = list(RANGES)
    if kind == 'block':
        return ob.BlockWorld(ext, num_blocks=2, dim_blocks=(1.5, 1.5), centers=[(6.5, 5.0), (4.0, 7.0)])
    if kind == 'channel':
        return ob.ChannelWorld(ext, (6.5, 5.0), 3., 0.6)
    return ob.FreeWorld()
