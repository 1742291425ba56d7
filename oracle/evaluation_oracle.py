"""CPU restatement of the metrics caller.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/evaluation_library.py:
  mean_reward :113-125, hotspot_info_reward :135-148, info_gain_reward :150-152, inst_regret :155-175,
  simple_regret :177-188, sample_regret :190-206, max_error :208-214, hotspot_error :216-243, MSE :248-258,
  update_metrics :262-314 (the prediction-dependent entries).
One GP prediction per path, in the reference's order; random test points from the global np.random exactly as the
reference draws them.  `world` is any object with .dim, .GP (or .models), .x1min/.x1max/.x2min/.x2max whose models
are oracle GP models; 2-D worlds only need .GP.
"""
import numpy as np
import scipy.spatial

from . import aq_oracle as aqlib


class EvaluationOracle(object):
    def __init__(self, world, reward_function='mean'):
        self.world = world
        if world.dim == 2:                                               # :36-41
            self.max_loc = world.GP.xvals[np.argmax(world.GP.zvals), :]
            self.max_val = np.max(world.GP.zvals)
        else:
            self.max_loc = world.models[0].xvals[np.argmax(world.models[0].zvals), 0:-1]
            self.max_val = np.max(world.models[0].zvals)
        self.reward_function = reward_function
        self.f_rew = {'hotspot_info': self.hotspot_info_reward, 'mean': self.mean_reward,      # :82-105
                      'info_gain': self.info_gain_reward, 'mes': self.mean_reward,
                      'exp_improve': self.mean_reward, 'naive': None, 'naive_value': None}[reward_function]

    def _world_predict(self, time, x1, x2):
        if self.world.dim == 2:
            return self.world.GP.predict_value(np.vstack([x1, x2]).T)
        return self.world.models[time].predict_value(np.vstack([x1, x2, time * np.ones(len(x1))]).T)

    def mean_reward(self, time, xvals, robot_model):
        data = np.array(xvals)
        mu, var = self._world_predict(time, data[:, 0], data[:, 1])
        return np.sum(mu)

    def info_gain_reward(self, time, xvals, robot_model):
        return aqlib.info_gain(time, xvals, robot_model)

    def hotspot_info_reward(self, time, xvals, robot_model):
        data = np.array(xvals)
        mu, var = self._world_predict(time, data[:, 0], data[:, 1])
        return self.info_gain_reward(time, xvals, robot_model) + 1.0 * np.sum(mu)

    def inst_regret(self, t, all_paths, selected_path, robot_model, param=None):
        value_omni = {}
        for path, points in all_paths.items():
            if param is None:
                value_omni[path] = self.f_rew(time=t, xvals=points, robot_model=robot_model)
            else:
                value_omni[path] = aqlib.mves(time=t, xvals=points, robot_model=robot_model,
                                              param=(self.max_val).reshape(1, 1))
        value_max = value_omni[max(value_omni, key=value_omni.get)]
        if param is None:
            value_selected = self.f_rew(time=t, xvals=selected_path, robot_model=robot_model)
        else:
            value_selected = aqlib.mves(time=t, xvals=selected_path, robot_model=robot_model,
                                        param=(self.max_val).reshape(1, 1))
        return value_max - value_selected, value_selected, value_max

    def simple_regret(self, xvals):
        error = 0.0
        for point in xvals:
            error += np.linalg.norm(np.array(point[0:-1]) - self.max_loc)
        return error / float(len(xvals))

    def sample_regret(self, robot_model):
        if robot_model.xvals is None:
            return 0., 0.
        gv = np.reshape(np.array(self.max_val), (1, 1))
        gl = np.reshape(np.array(self.max_loc), (1, 2))
        x = robot_model.xvals if robot_model.dimension == 2 else robot_model.xvals[:, 0:-1]
        return (np.mean(scipy.spatial.distance.cdist(gl, x)),
                np.mean(scipy.spatial.distance.cdist(gv, robot_model.zvals)))

    def max_error(self, max_loc, max_val):
        return np.linalg.norm(max_loc[0:-1] - self.max_loc), np.linalg.norm(max_val - self.max_val)

    def _draw(self, time, robot_model, NTEST):
        x1 = np.random.random_sample((NTEST, 1)) * (self.world.x1max - self.world.x1min) + self.world.x1min
        x2 = np.random.random_sample((NTEST, 1)) * (self.world.x2max - self.world.x2min) + self.world.x2min
        x1 = x1.reshape((NTEST,))
        x2 = x2.reshape((NTEST,))
        pred_world, _ = self._world_predict(time, x1, x2)
        if self.world.dim == 2:
            pred_robot, _ = robot_model.predict_value(np.vstack([x1, x2]).T)
        else:
            pred_robot, _ = robot_model.predict_value(np.vstack([x1, x2, time * np.ones(NTEST)]).T)
        return pred_world, pred_robot

    def hotspot_error(self, time, robot_model, NTEST=100, NHS=100):
        pred_world, pred_robot = self._draw(time, robot_model, NTEST)
        order = np.argsort(pred_world, axis=None)
        return ((pred_world[order[0:NHS]] - pred_robot[order[0:NHS]]) ** 2).mean()

    def MSE(self, time, robot_model, NTEST=100):
        pred_world, pred_robot = self._draw(time, robot_model, NTEST)
        return ((pred_world - pred_robot) ** 2).mean()

    def step_metrics(self, t, robot_model, all_paths, selected_path, max_loc, max_val):
        """The prediction-dependent entries of update_metrics, in its order (:271-300)."""
        m = {}
        m['simple_regret'] = self.simple_regret(selected_path)
        m['sample_regret_loc'], m['sample_regret_val'] = self.sample_regret(robot_model)
        m['max_loc_error'], m['max_val_error'] = self.max_error(max_loc, max_val)
        if self.reward_function in ('naive', 'naive_value'):                                    # :280-284
            m['instant_regret'] = m['max_val_regret'] = m['mes_reward_robot'] = m['mes_reward_omni'] = -1.
        else:
            m['instant_regret'], _, _ = self.inst_regret(t, all_paths, selected_path, robot_model)
            m['max_val_regret'], m['mes_reward_robot'], m['mes_reward_omni'] = self.inst_regret(
                t, all_paths, selected_path, robot_model, param='info_regret')
        m['info_gain_reward'] = self.info_gain_reward(t, selected_path, robot_model)
        m['MSE'] = self.MSE(t, robot_model, NTEST=200)
        m['hotspot_error'] = self.hotspot_error(t, robot_model, NTEST=200, NHS=100)
        return m
