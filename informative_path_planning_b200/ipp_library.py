"""The legacy monolith's call sites for the hot path (reference ipp_library.py), CUDA-backed.

ipp_library.py is the older single-file copy of the stack that demo.ipynb still imports; the north star names it as a
call site.  What matters of it for the hot path:
  GPModel (:32-148)        GPy wrapper WITHOUT noise_var -> the regression runs with GPy's default likelihood variance 1.0
                           and ignores .noise (:145); predict_value(xvals, TEMP=False) always includes the likelihood;
                           add_data_and_temp_model builds a throw-away model on (data + candidate points)
  rewards (:1485-1959)     info_gain / mean_UCB / hotspot_info_UCB / exp_improvement: the same closed forms as aq_library,
                           but WITHOUT its no-data early-outs (with no observations they score the prior);
                           mves (:1786-1836): entropy_of_n(var) - entropy_of_tn(None, y*, mu, var), which is algebraically
                           the same utility as aq_library.mves (-log Phi(g) + g phi(g) / (2 Phi(g)), g = (y* - mu) / var)
                           but summed ONCE (no double add): exactly half of aq_library.mves
  MCTS.get_reward (:689-708)  concatenates the sample points of the whole action sequence and calls the reward ONCE on the
                           unmodified belief -- already the batched seam: one device call per rollout (`sequence_reward`).
Everything else of the monolith (Environment, path generators, Robot, Evaluation, plotting) is superseded by the split
libraries and out of scope.
"""
import copy

import numpy as np

from . import aq_library as aqlib
from . import gpmodel_library as gplib


class GPModel(gplib.GPModel):
    """reference ipp_library.py:32-148"""

    def __init__(self, ranges, lengthscale, variance, noise=0.0001, dimension=2, kernel='rbf'):
        if dimension != 2:
            raise ValueError('Environment must have dimension 2 \'rbf\'')
        super(GPModel, self).__init__(ranges, lengthscale, variance, noise, dimension, kernel)
        self.dim = dimension
        self.temp_model = None

    def _noise_var(self):
        return 1.0                                   # GPRegression(X, Y, kern) without noise_var (:145): GPy's default

    def predict_value(self, xvals, TEMP=False):      # :73-103 (include_likelihood = True always)
        assert xvals.shape[0] >= 1
        assert xvals.shape[1] == self.dim
        target = self.temp_model if TEMP else self
        if target is None or target.model is None:
            return np.zeros((xvals.shape[0], 1)), np.ones((xvals.shape[0], 1)) * self.variance
        return gplib.GPModel.predict_value(target, xvals, include_noise=True)

    def add_data_and_temp_model(self, xvals, zvals):  # :105-124
        temp = copy.copy(self)
        temp.temp_model = None
        gplib.GPModel.add_data(temp, xvals, zvals)
        self.temp_model = temp


def _queries(xvals):
    data = np.array(xvals, dtype=float)
    return np.vstack([data[:, 0], data[:, 1]]).T


def _ucb_terms(time, queries, robot_model):
    mu, var = robot_model.predict_value(queries)
    return np.sum(mu), np.sqrt(aqlib.ucb_beta(time)) * np.sum(np.fabs(var))


def mean_UCB(time, xvals, robot_model, param=None):              # :1529-1544
    if robot_model.xvals is not None:
        return aqlib.mean_UCB(time, xvals, robot_model)           # one fused device pass
    a, b = _ucb_terms(time, _queries(xvals), robot_model)         # no data: the legacy code scores the prior
    return a + b


def info_gain(time, xvals, robot_model, param=None):              # :1485-1526
    return aqlib.info_gain(time, xvals, robot_model)


def hotspot_info_UCB(time, xvals, robot_model, param=None):       # :1547-1562
    if robot_model.xvals is not None:
        return aqlib.hotspot_info_UCB(time, xvals, robot_model)
    a, b = _ucb_terms(time, _queries(xvals), robot_model)
    return info_gain(time, xvals, robot_model) + a + b


def exp_improvement(time, xvals, robot_model, param=None):       # :1936-1959
    return aqlib.exp_improvement(time, xvals, robot_model, param)


def mves(time, xvals, robot_model, param):                        # :1786-1836
    if param[0] is None:
        return 1.0
    if robot_model.xvals is None:                                 # prior everywhere: closed form on the host
        q = _queries(xvals)
        mean, var = robot_model.predict_value(q)
        maxes = np.asarray(param[0], dtype=float).reshape(-1)
        total = sum(np.sum(entropy_of_n(var) - entropy_of_tn(None, m, mean, var)) for m in maxes)
        return total / maxes.shape[0]
    return 0.5 * aqlib.mves(time, np.asarray(xvals, dtype=float), robot_model, param)


def entropy_of_n(var):                                            # :1838-1839
    return np.log(np.sqrt(2.0 * np.pi * var))


def entropy_of_tn(a, b, mu, var):                                 # :1841-1869
    from scipy.stats import norm
    if a is None:
        Phi_alpha, phi_alpha, alpha = 0, 0, 0
    else:
        alpha = (a - mu) / var
        Phi_alpha, phi_alpha = norm.cdf(alpha), norm.pdf(alpha)
    if b is None:
        Phi_beta, phi_beta, beta = 1, 0, 0
    else:
        beta = (b - mu) / var
        Phi_beta, phi_beta = norm.cdf(beta), norm.pdf(beta)
    Z = Phi_beta - Phi_alpha
    return np.log(Z * np.sqrt(2.0 * np.pi * var)) + (alpha * phi_alpha - beta * phi_beta) / (2.0 * Z)


def sequence_reward(aquisition_function, f_rew, t, sample_lists, robot_model, max_param=None, current_max=None):
    """MCTS.get_reward (:689-708): the samples of every node of the sequence concatenated, ONE reward call on the
    unmodified belief.  sample_lists: list of point lists (one per tree node of the sequence)."""
    obs = [p for samples in sample_lists for p in samples]
    if f_rew in ('mes', 'maxs-mes'):
        return aquisition_function(time=t, xvals=obs, robot_model=robot_model, param=max_param)
    if f_rew == 'exp_improve':
        return aquisition_function(time=t, xvals=obs, robot_model=robot_model, param=[current_max])
    return aquisition_function(time=t, xvals=obs, robot_model=robot_model)
