"""Drop-in for the reference's robot_library.Robot (reference :36-337) without the matplotlib figures.

The primary caller of the hot path: builds the OnlineGPModel belief, and every
mission step (planner :256-337) finds the predicted maximum over the 100 x 100 grid (predict_max :222-254), chooses a
trajectory -- myopically (choose_trajectory :155-208: all candidate paths of the pose scored in ONE device call) or with
the tree search (cMCTS, one device call per rollout) --, updates the Evaluation metrics, observes along the chosen path
and appends the block to the belief.  Random numbers are drawn from the global np.random / random in the reference's
order, so a seeded mission visits the same poses as the reference's.

Not carried over: visualize_trajectory / visualize_reward / visualize_world_model draw no figures; they return the
100 x 100 grid data the reference plots (one batched predict / reward call).  Path generators: 'dubins' and the default
straight fan ('equal_dubins' is broken in the reference, the 'fully_reachable_*' generators are out of scope).
"""
import logging
import os

import numpy as np

from . import aq_library as aqlib
from . import gpmodel_library as gplib
from . import mcts_library as mctslib
from . import obstacles as obslib
from . import paths_library as pathlib

logger = logging.getLogger('robot')

_BATCHED = ('mean', 'mes', 'exp_improve', 'info_gain', 'hotspot_info')


class Robot(object):
    def __init__(self, **kwargs):
        self.ranges = kwargs['extent']
        self.dimension = kwargs['dimension']
        self.create_animation = kwargs.get('create_animation', False)
        self.eval = kwargs['evaluation']
        self.loc = kwargs['start_loc']
        self.time = kwargs['start_time']
        self.sample_world = kwargs['sample_world']
        self.f_rew = kwargs['f_rew']
        self.frontier_size = kwargs['frontier_size']
        self.discretization = kwargs.get('discretization')
        self.tree_type = kwargs['tree_type']
        self.path_option = kwargs['path_generator']

        self.nonmyopic = kwargs['nonmyopic']
        self.comp_budget = kwargs['computation_budget']
        self.roll_length = kwargs['rollout_length']
        self.horizon_length = kwargs['horizon_length']
        self.sample_step = kwargs['sample_step']
        self.turning_radius = kwargs['turning_radius']
        self.goal_only = kwargs.get('goal_only', False)
        self.obstacle_world = kwargs.get('obstacle_world') or obslib.FreeWorld()
        self.learn_params = kwargs.get('learn_params', False)
        self.use_cost = kwargs.get('use_cost', False)
        self.MIN_COLOR = kwargs.get('MIN_COLOR')
        self.MAX_COLOR = kwargs.get('MAX_COLOR')

        self.maxes = []
        self.current_max = -1000
        self.current_max_loc = [0, 0]
        self.max_locs = None
        self.max_val = None
        self.target = None
        self.noise = kwargs['noise']
        self.trajectory = []
        self.dist = 0

        table = {'hotspot_info': aqlib.hotspot_info_UCB, 'mean': aqlib.mean_UCB, 'info_gain': aqlib.info_gain,
                 'mes': aqlib.mves, 'exp_improve': aqlib.exp_improvement, 'naive': aqlib.naive,
                 'naive_value': aqlib.naive_value}
        if self.f_rew not in table:
            raise ValueError('Only \'hotspot_info\' and \'mean\' and \'info_gain\' and \'mes\' and \'exp_improve\' '
                             'reward fucntions supported.')
        self.aquisition_function = table[self.f_rew]
        if self.f_rew == 'naive':
            self.sample_num, self.sample_radius = 3, 1.5
        elif self.f_rew == 'naive_value':
            self.sample_num, self.sample_radius = 3, 3.0

        self.GP = gplib.OnlineGPModel(ranges=self.ranges, lengthscale=kwargs['init_lengthscale'],
                                      variance=kwargs['init_variance'], noise=self.noise, dimension=self.dimension)
        kernel_dataset, prior_dataset = kwargs.get('kernel_dataset'), kwargs.get('prior_dataset')
        kernel_file = kwargs.get('kernel_file')
        if kernel_dataset is not None and prior_dataset is not None:                 # reference :122-135
            data = np.vstack([prior_dataset[0], kernel_dataset[0]])
            observations = np.vstack([prior_dataset[1], kernel_dataset[1]])
            self.GP.train_kernel(data, observations, kernel_file)
        elif kernel_dataset is not None:
            self.GP.train_kernel(kernel_dataset[0], kernel_dataset[1], kernel_file)
        elif kernel_file is not None:
            self.GP.load_kernel()
        if prior_dataset is not None:
            self.GP.add_data(prior_dataset[0], prior_dataset[1])

        gen_args = (self.frontier_size, self.horizon_length, self.turning_radius, self.sample_step, self.ranges,
                    self.obstacle_world)
        if self.path_option == 'dubins':
            self.path_generator = pathlib.Dubins_Path_Generator(*gen_args)
        elif self.path_option in ('equal_dubins', 'fully_reachable_goal', 'fully_reachable_step'):
            raise ValueError('path_generator %r is not carried over (see DESIGN.md, out of scope); use \'dubins\' or '
                             '\'default\'' % (self.path_option,))
        else:
            self.path_generator = pathlib.Path_Generator(*gen_args)

    # ---- myopic choice: reference :155-208 ---------------------------------------------------------------
    def choose_trajectory(self, t):
        value = {}
        param = None
        if self.f_rew == 'mes':
            self.max_val, self.max_locs, self.target = aqlib.sample_max_vals(
                self.GP, t=t, visualize=True, f_rew=self.f_rew, obstacles=self.obstacle_world)
        elif self.f_rew in ('naive', 'naive_value'):
            self.max_val, self.max_locs, self.target = aqlib.sample_max_vals(
                self.GP, t=t, obstacles=self.obstacle_world, visualize=True, f_rew=self.f_rew, nK=int(self.sample_num))
            param = ((self.max_val, self.max_locs, self.target), self.sample_radius)
        self.predict_max(t=t)

        paths, true_paths = self.path_generator.get_path_set(self.loc)
        if self.f_rew == 'mes':
            param = (self.max_val, self.max_locs, self.target)
        elif self.f_rew == 'exp_improve':
            param = [self.current_max] if len(self.maxes) == 0 else self.maxes

        if len(paths) > 0:
            if self.f_rew in _BATCHED and self.GP.xvals is not None:
                value = mctslib.evaluate_root_actions(self.GP, paths, self.f_rew, t, param)       # ONE device call
            else:
                for path, points in paths.items():
                    value[path] = self.aquisition_function(time=t, xvals=np.array(points), robot_model=self.GP, param=param)
            if self.use_cost:
                for path in value:
                    cost = float(self.path_generator.path_cost(true_paths[path]))
                    value[path] = value[path] / (cost if cost != 0.0 else 100.0)
        try:
            top = max(value.values())
            best_key = np.random.choice([key for key in value.keys() if value[key] == top])
            return paths[best_key], true_paths[best_key], value[best_key], paths, value, self.max_locs
        except Exception:
            return None

    def collect_observations(self, xobs):
        """reference :210-220"""
        zobs = self.sample_world(xobs)
        self.GP.add_data(xobs, zobs)
        for z, x in zip(zobs, xobs):
            if z[0] > self.current_max:
                self.current_max = z[0]
                self.current_max_loc = [x[0], x[1]]

    def predict_max(self, t=0):
        """reference :222-254: arg-max of the posterior mean over the 100 x 100 grid (one batched device predict)."""
        return mctslib.predict_max(self.GP, self.ranges, t=t, time=self.time)

    # ---- mission loop: reference :256-337 ----------------------------------------------------------------
    def planner(self, T):
        self.trajectory = []
        self.dist = 0
        for t in range(T):
            self.time = t
            logger.info('[{}] Current Location: {}'.format(t, self.loc))
            pred_loc, pred_val = self.predict_max()
            logger.info('Current predicted max and value: {} \t {}'.format(pred_loc, pred_val))

            if not self.nonmyopic:
                sampling_path, best_path, best_val, all_paths, all_values, max_locs = self.choose_trajectory(t=t)
            else:
                if self.f_rew in ('naive', 'naive_value'):
                    param = (self.sample_num, self.sample_radius)
                elif self.f_rew == 'exp_improve':
                    param = self.current_max
                else:
                    param = None
                mcts = mctslib.cMCTS(self.comp_budget, self.GP, self.loc, self.roll_length, self.path_generator,
                                     self.aquisition_function, self.f_rew, t, aq_param=param, use_cost=self.use_cost,
                                     tree_type=self.tree_type)
                (sampling_path, best_path, best_val, all_paths, all_values, self.max_locs, self.max_val,
                 self.target) = mcts.choose_trajectory(t=t)

            start = self.loc
            for m in best_path:
                self.dist += np.sqrt((start[0] - m[0]) ** 2 + (start[1] - m[1]) ** 2)
                start = m

            if self.eval is not None:
                self.eval.update_metrics(t=len(self.trajectory), robot_model=self.GP, all_paths=all_paths,
                                         selected_path=sampling_path, value=best_val, max_loc=pred_loc, max_val=pred_val,
                                         params=[self.current_max, self.current_max_loc, self.max_val, self.max_locs],
                                         dist=self.dist)
            if best_path is None:
                break

            data = np.array(sampling_path)
            x1, x2 = data[:, 0], data[:, 1]
            if self.dimension == 2:
                xlocs = np.vstack([x1, x2]).T
            elif self.dimension == 3:
                xlocs = np.vstack([x1, x2, t * np.ones(len(x1))]).T
            else:
                raise ValueError('Only 2D or 3D worlds supported!')
            self.collect_observations(xlocs)

            if t < T // 3 and self.learn_params:      # reference :328 is Python 2: T/3 on ints floors
                self.GP.train_kernel()
            self.trajectory.append(best_path)
            self.loc = sampling_path[-1]

        out_dir = os.path.join('./figures', str(self.f_rew))
        if not os.path.exists(out_dir):
            os.makedirs(out_dir)
        np.savetxt(os.path.join(out_dir, 'robot_model.csv'), (self.GP.xvals[:, 0], self.GP.xvals[:, 1], self.GP.zvals[:, 0]))

    # ---- the grids the reference plots (no figures) ----------------------------------------------------------
    def _grid(self):
        x1vals = np.linspace(self.ranges[0], self.ranges[1], 100)
        x2vals = np.linspace(self.ranges[2], self.ranges[3], 100)
        x1, x2 = np.meshgrid(x1vals, x2vals, sparse=False, indexing='xy')
        if self.dimension == 2:
            return x1, x2, np.vstack([x1.ravel(), x2.ravel()]).T
        return x1, x2, np.vstack([x1.ravel(), x2.ravel(), self.time * np.ones(len(x1.ravel()))]).T

    def visualize_world_model(self, screen=True, filename='SUMMARY'):
        """The data of reference :496-535: (x1, x2, posterior mean on the 100 x 100 grid, variance grid)."""
        x1, x2, data = self._grid()
        observations, var = self.GP.predict_value(data)
        return x1, x2, observations.reshape(x1.shape), var.reshape(x1.shape)

    def visualize_trajectory(self, screen=True, filename='SUMMARY', best_path=None, maxes=None, all_paths=None,
                             all_vals=None):
        """The data of reference :344-433 (same grid predict as visualize_world_model)."""
        return self.visualize_world_model(screen, filename)

    def visualize_reward(self, screen=False, filename='REWARD', t=0):
        """The data of reference :435-493: the acquisition function's per-point values (FVECTOR) on the grid."""
        x1, x2, data = self._grid()
        if self.f_rew == 'mes':
            param = (self.max_val, self.max_locs, self.target)
        elif self.f_rew == 'exp_improve':
            param = [self.current_max] if len(self.maxes) == 0 else self.maxes
        elif self.f_rew in ('naive', 'naive_value'):
            param = ((self.max_val, self.max_locs, self.target), self.sample_radius)
        else:
            param = None
        reward = self.aquisition_function(time=t, xvals=data, robot_model=self.GP, param=param, FVECTOR=True)
        return x1, x2, np.asarray(reward).reshape(x1.shape)

    def plot_information(self):
        self.eval.plot_metrics()
