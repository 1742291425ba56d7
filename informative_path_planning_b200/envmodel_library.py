"""Drop-in for the reference's envmodel_library.Environment (reference :27-260) without the plotting: a rectangular
Gaussian world drawn from the GP prior on a NUM_PTS x NUM_PTS grid and served back through sample_value.

SURVEY.md 8f rank 4 (world synthesis).  The arithmetic goes through this package's GPModel (the GPy-wrapping model of
gpmodel_library.py:25-185): one prior draw at the first grid point, then ONE joint posterior sample of the remaining
NUM_PTS^2 - 1 points (full covariance on the device, np.random.multivariate_normal on the host so that the reference's
random stream is consumed identically), regenerated until the maximum lies inside the 5 % border (:150-154).
Quirk kept: sample_value adds one SCALAR noise draw to every point of the batch (:260).
"""
import copy
import logging

import numpy as np

from .gpmodel_library import GPModel
from .obstacles import FreeWorld

logger = logging.getLogger('robot')


class Environment(object):
    def __init__(self, ranges, NUM_PTS, variance, lengthscale, noise=0.0001, visualize=False, seed=None, dim=2,
                 model=None, MIN_COLOR=-25.0, MAX_COLOR=25.0, obstacle_world=None, time_duration=None, gp_cls=GPModel):
        self.variance = variance
        self.lengthscale = lengthscale
        self.dim = dim
        self.noise = noise
        self.obstacle_world = FreeWorld() if obstacle_world is None else obstacle_world
        self.time_duration = time_duration
        logger.info('Environment seed: {}'.format(seed))
        self.x1min, self.x1max = float(ranges[0]), float(ranges[1])
        self.x2min, self.x2max = float(ranges[2]), float(ranges[3])
        if model is not None:                                              # reference :65-66 (pre-built world)
            self.GP = model
            return
        x1 = np.linspace(self.x1min, self.x1max, NUM_PTS)
        x2 = np.linspace(self.x2min, self.x2max, NUM_PTS)
        x1vals, x2vals = np.meshgrid(x1, x2, sparse=False, indexing='xy')
        bb = ((ranges[1] - ranges[0]) * 0.05, (ranges[3] - ranges[2]) * 0.05)   # maxima may not sit on the border
        ranges = (ranges[0] + bb[0], ranges[1] - bb[0], ranges[2] + bb[1], ranges[3] - bb[1])
        self.models = {}
        if self.time_duration is None:
            self.time_duration = 1
        for T in range(self.time_duration):
            maxima = [self.x1min, self.x2min]
            while (maxima[0] < ranges[0] or maxima[0] > ranges[1] or maxima[1] < ranges[2] or maxima[1] > ranges[3] or
                   self.obstacle_world.in_obstacle(maxima, buff=0.0)):
                self.GP = gp_cls(ranges=ranges, lengthscale=lengthscale, variance=variance, dimension=self.dim)
                if self.dim == 2:
                    data = np.vstack([x1vals.ravel(), x2vals.ravel()]).T
                else:
                    data = np.vstack([x1vals.ravel(), x2vals.ravel(), T * np.ones(len(x1vals.ravel()))]).T
                if T == 0:
                    xsamples = np.reshape(np.array(data[0, :]), (1, dim))
                    mean, var = self.GP.predict_value(xsamples, include_noise=False)
                    if seed is not None:
                        np.random.seed(seed)
                        seed += 1
                    zsamples = np.reshape(np.random.normal(loc=0, scale=np.sqrt(var)), (1, 1))
                    self.GP.add_data(xsamples, zsamples)
                    np.random.seed(seed)
                    observations = self.GP.posterior_samples(data[1:, :], full_cov=True, size=1)
                    self.GP.add_data(data[1:, :], observations)
                else:
                    np.random.seed(seed)
                    observations = self.models[T - 1].posterior_samples(data, full_cov=True, size=1)
                    self.GP.add_data(data, observations)
                maxima = self.GP.xvals[np.argmax(self.GP.zvals), :]
            self.models[T] = copy.deepcopy(self.GP)
        maxind = np.argmax(self.GP.zvals)
        self.max_val = self.GP.zvals[maxind, :]
        self.max_loc = self.GP.xvals[maxind, :]
        logger.info('Environment initialized with bounds X1: ({}, {})  X2: ({}, {})'.format(
            self.x1min, self.x1max, self.x2min, self.x2max))

    def sample_value(self, xvals, time=None):
        """Noisy sample of the world at xvals (M, dim) -> (M, 1); reference :241-260."""
        assert xvals.shape[0] >= 1
        assert xvals.shape[1] == self.dim
        if self.dim == 2:
            mean, var = self.GP.predict_value(xvals, include_noise=False)
        elif self.dim == 3:
            T = xvals[1, self.dim - 1]
            mean, var = self.models[T].predict_value(xvals, include_noise=False)
        else:
            raise ValueError('Model dimension must be 2 or 3!')
        return mean + np.random.normal(loc=0, scale=np.sqrt(self.noise))
