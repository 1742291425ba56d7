"""CPU restatement of the mission loop.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/robot_library.py: Robot.__init__ :36-153 (model + path generator), choose_trajectory :155-208,
collect_observations :210-220, predict_max :222-254, planner :256-337 (without the figures and the csv dump).  Built on the
other oracle modules (OnlineGPModelOracle, aq_oracle, mcts_oracle, evaluation_oracle); global np.random / random are
consumed in the reference's order.  `path_generator` is passed in as an object: the straight fan of mcts_oracle, or any
generator with get_path_set(pose) (the `dubins` C library is not available, see paths_library of the package).
"""
import numpy as np

from . import aq_oracle as aqlib
from . import mcts_oracle as mctslib
from .gpmodel_oracle import OnlineGPModelOracle


class RobotOracle(object):
    def __init__(self, sample_world, start_loc, extent, init_lengthscale, init_variance, noise, path_generator, f_rew,
                 nonmyopic, computation_budget, rollout_length, tree_type, evaluation=None, prior_dataset=None,
                 dimension=2):
        self.ranges, self.dimension, self.loc, self.time = extent, dimension, start_loc, 0
        self.sample_world, self.f_rew, self.eval = sample_world, f_rew, evaluation
        self.nonmyopic, self.comp_budget, self.roll_length, self.tree_type = nonmyopic, computation_budget, rollout_length, tree_type
        self.maxes, self.current_max, self.current_max_loc = [], -1000, [0, 0]
        self.max_locs = self.max_val = self.target = None
        self.aquisition_function = {'hotspot_info': aqlib.hotspot_info_UCB, 'mean': aqlib.mean_UCB,       # :93-113
                                    'info_gain': aqlib.info_gain, 'mes': aqlib.mves,
                                    'exp_improve': aqlib.exp_improvement, 'naive': aqlib.naive,
                                    'naive_value': aqlib.naive_value}[f_rew]
        if f_rew == 'naive':                                                                                # :105-112
            self.sample_num, self.sample_radius = 3, 1.5
        elif f_rew == 'naive_value':
            self.sample_num, self.sample_radius = 3, 3.0
        self.GP = OnlineGPModelOracle(ranges=extent, lengthscale=init_lengthscale, variance=init_variance, noise=noise,
                                      dimension=dimension)                                                  # :119
        if prior_dataset is not None:                                                                       # :138-139
            self.GP.add_data(prior_dataset[0], prior_dataset[1])
        self.path_generator = path_generator
        self.trajectory, self.dist = [], 0

    def choose_trajectory(self, t):                                                                         # :155-208
        value = {}
        param = None
        if self.f_rew == 'mes':
            self.max_val, self.max_locs, self.target = aqlib.sample_max_vals(self.GP, t=t)
        elif self.f_rew in ('naive', 'naive_value'):                                                        # :168-170
            self.max_val, self.max_locs, self.target = aqlib.sample_max_vals(self.GP, t=t, f_rew=self.f_rew,
                                                                             nK=int(self.sample_num))
            param = ((self.max_val, self.max_locs, self.target), self.sample_radius)
        self.predict_max(t=t)
        paths, true_paths = self.path_generator.get_path_set(self.loc)
        for path, points in paths.items():
            if self.f_rew == 'mes':
                param = (self.max_val, self.max_locs, self.target)
            elif self.f_rew == 'exp_improve':
                param = [self.current_max] if len(self.maxes) == 0 else self.maxes
            value[path] = self.aquisition_function(time=t, xvals=points, robot_model=self.GP, param=param)
        try:
            best_key = np.random.choice([key for key in value.keys() if value[key] == max(value.values())])
            return paths[best_key], true_paths[best_key], value[best_key], paths, value, self.max_locs
        except Exception:
            return None

    def collect_observations(self, xobs):                                                                   # :210-220
        zobs = self.sample_world(xobs)
        self.GP.add_data(xobs, zobs)
        for z, x in zip(zobs, xobs):
            if z[0] > self.current_max:
                self.current_max = z[0]
                self.current_max_loc = [x[0], x[1]]

    def predict_max(self, t=0):                                                                             # :222-254
        if self.GP.xvals is None:
            return np.array([0., 0.]), 0.
        x1vals = np.linspace(self.ranges[0], self.ranges[1], 100)
        x2vals = np.linspace(self.ranges[2], self.ranges[3], 100)
        x1, x2 = np.meshgrid(x1vals, x2vals, sparse=False, indexing='xy')
        data = np.vstack([x1.ravel(), x2.ravel()]).T
        observations, var = self.GP.predict_value(data)
        return data[np.argmax(observations), :], np.max(observations)

    def planner(self, T):                                                                                   # :256-337
        self.trajectory, self.dist = [], 0
        for t in range(T):
            self.time = t
            pred_loc, pred_val = self.predict_max()
            if not self.nonmyopic:
                sampling_path, best_path, best_val, all_paths, all_values, max_locs = self.choose_trajectory(t=t)
            else:
                if self.f_rew in ('naive', 'naive_value'):                                                  # :278-285
                    param = (self.sample_num, self.sample_radius)
                else:
                    param = self.current_max if self.f_rew == 'exp_improve' else None
                mcts = mctslib.cMCTSOracle(self.comp_budget, self.GP, self.loc, self.roll_length, self.path_generator,
                                           self.aquisition_function, self.f_rew, t, aq_param=param, tree_type=self.tree_type)
                (sampling_path, best_path, best_val, all_paths, all_values, self.max_locs, self.max_val,
                 self.target) = mcts.choose_trajectory(t=t)
            start = self.loc
            for m in best_path:
                self.dist += np.sqrt((start[0] - m[0]) ** 2 + (start[1] - m[1]) ** 2)
                start = m
            if self.eval is not None:
                self.eval(t=len(self.trajectory), robot_model=self.GP, all_paths=all_paths, selected_path=sampling_path,
                          value=best_val, max_loc=pred_loc, max_val=pred_val, dist=self.dist, current_max=self.current_max)
            data = np.array(sampling_path)
            self.collect_observations(np.vstack([data[:, 0], data[:, 1]]).T)
            self.trajectory.append(best_path)
            self.loc = sampling_path[-1]
