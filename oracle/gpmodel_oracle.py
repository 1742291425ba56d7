"""CPU restatement of the belief model (L1 of SURVEY.md).  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/gpmodel_library.py:
  GPModel            :25-132   (GPy-wrapping model: rebuild on every add_data, Cholesky-form predict)
  OnlineGPModel      :187-328  (explicit (K + s2 I)^-1, block-inverse append, numpy predict)
  posterior_samples  :130-132, :331-364

float64 throughout; no clipping of negative variances (reference has none).
"""
import numpy as np

from . import gpy_restated as G


class GPModelOracle(object):
    """gpmodel_library.py:25-132."""

    def __init__(self, ranges, lengthscale, variance, noise=0.0001, dimension=2, kernel='rbf', period=None):
        self.noise = noise
        self.lengthscale = lengthscale
        self.variance = variance
        self.ranges = ranges
        self.xvals = None
        self.zvals = None
        if dimension == 2:                                   # :51-53
            self.dimension = dimension
            self.asymmetric = False
        elif dimension == 3:                                 # :55-59
            if len(lengthscale) < dimension:
                raise ValueError('Lengthscale vector must have same length as dimension.')
            self.dimension = dimension
            self.asymmetric = True
        else:                                                # :60-62
            raise ValueError('Environment must have dimension 2 or 3')
        if kernel == 'rbf':                                  # :64-65
            self.kern = G.RBF(input_dim=self.dimension, lengthscale=lengthscale, variance=variance,
                              ARD=self.asymmetric)
        else:                                                # :66-77 ('rbf-period' is out of scope)
            raise ValueError('Kernel type must by \'rbf\'')
        self.model = None

    def predict_value(self, xvals, include_noise=True):      # :82-103
        assert xvals.shape[0] >= 1
        assert xvals.shape[1] == self.dimension
        n_points = xvals.shape[0]
        if self.model is None:
            return np.zeros((n_points, 1)), np.ones((n_points, 1)) * self.variance
        return self.model.predict(xvals, full_cov=False, include_likelihood=include_noise)

    def add_data(self, xvals, zvals):                        # :106-128
        self.xvals = xvals if self.xvals is None else np.vstack([self.xvals, xvals])
        self.zvals = zvals if self.zvals is None else np.vstack([self.zvals, zvals])
        # ":124 if self.model == None or True" -> always a brand-new GPRegression
        self.model = G.GPRegression(np.array(self.xvals), np.array(self.zvals), self.kern, noise_var=self.noise)

    def posterior_samples(self, xvals, size=10, full_cov=True):   # :130-132
        return self.model.posterior_samples_f(xvals, size, full_cov=full_cov)


class OnlineGPModelOracle(GPModelOracle):
    """gpmodel_library.py:187-328."""

    def __init__(self, ranges, lengthscale, variance, noise=0.0001, dimension=2, kernel='rbf', update_legacy=False):
        super(OnlineGPModelOracle, self).__init__(ranges, lengthscale, variance, noise, dimension, kernel)
        self._K = None
        self._woodbury_inv = None
        self._woodbury_vector = None
        self.update_legacy = update_legacy

    # :208-228
    def init_model(self, xvals, zvals):
        self.xvals = xvals
        self.zvals = zvals
        self._K = self.kern.K(self.xvals)
        Ky = self._K.copy()
        G.diag_add(Ky, self.noise + 1e-8)
        Wi, _, _, _ = G.pdinv(Ky)
        self._woodbury_inv = Wi
        self._woodbury_vector = np.dot(self._woodbury_inv, self.zvals)

    # :230-279 (incremental=True is the only branch add_data reaches)
    def update_model(self, xvals, zvals, incremental=True):
        assert self.xvals is not None and self.zvals is not None
        Kx = self.kern.K(self.xvals, xvals)
        S = self.kern.K(xvals, xvals)
        self._K = np.block([[self._K, Kx], [Kx.T, S]])
        self.xvals = np.vstack([self.xvals, xvals])
        self.zvals = np.vstack([self.zvals, zvals])
        if incremental:
            Pinv = self._woodbury_inv
            Q = Kx
            R = Kx.T
            M = S - np.dot(np.dot(R, Pinv), Q)                                  # :252
            G.diag_add(M, self.noise + 1e-8)                                    # :254
            M, _, _, _ = G.pdinv(M)                                             # :255
            Pnew = Pinv + np.dot(np.dot(np.dot(np.dot(Pinv, Q), M), R), Pinv)   # :257 (association order kept)
            Qnew = -np.dot(np.dot(Pinv, Q), M)                                  # :258
            Rnew = -np.dot(np.dot(M, R), Pinv)                                  # :259
            self._woodbury_inv = np.block([[Pnew, Qnew], [Rnew, M]])            # :262-265
        else:
            Ky = self._K.copy()
            G.diag_add(Ky, self.noise + 1e-8)
            self._woodbury_inv, _, _, _ = G.pdinv(Ky)
        self._woodbury_vector = np.dot(self._woodbury_inv, self.zvals)          # :273

    # :281-301
    def add_data(self, xvals, zvals):
        if self.xvals is None:
            assert self.zvals is None
            self.init_model(xvals, zvals)
        else:
            assert self.zvals is not None
            self.update_model(xvals, zvals)
        if self.update_legacy:
            self.model = G.GPRegression(np.array(self.xvals), np.array(self.zvals), self.kern, noise_var=self.noise)

    # :303-328
    def predict_value(self, xvals, include_noise=True, full_cov=False):
        assert xvals.shape[0] >= 1
        assert xvals.shape[1] == self.dimension
        n_points = xvals.shape[0]
        if self.xvals is None:
            return np.zeros((n_points, 1)), np.ones((n_points, 1)) * self.variance
        Kx = self.kern.K(self.xvals, xvals)
        mu = np.dot(Kx.T, self._woodbury_vector)
        if mu.ndim == 1:
            mu = mu.reshape(-1, 1)
        if full_cov:
            var = self.kern.K(xvals) - np.dot(Kx.T, np.dot(self._woodbury_inv, Kx))
        else:
            var = (self.kern.Kdiag(xvals) - np.sum(np.dot(self._woodbury_inv.T, Kx) * Kx, 0))[:, None]
        if include_noise:
            var = var + self.noise          # the reference adds the scalar to every entry, also for full_cov
        return mu, var

    # :331-364 (output_dim == 1 branch; global np.random like the reference)
    def posterior_samples(self, xvals, size=10, full_cov=True):
        m, v = self.predict_value(xvals, include_noise=True, full_cov=full_cov)
        if not full_cov:
            return np.random.multivariate_normal(m.flatten(), np.diag(v.flatten()), size).T
        return np.random.multivariate_normal(m.flatten(), v, size).T

    @property
    def woodbury_inv(self):
        return self._woodbury_inv

    @property
    def woodbury_vector(self):
        return self._woodbury_vector


def predict_chunked(model, xq, include_noise=True, chunk=8192):
    """The same arithmetic as OnlineGPModelOracle.predict_value, over row chunks of xq so the
    N x M temporaries stay bounded (used by the CPU-baseline timing in bench.py)."""
    mus, vs = [], []
    for s in range(0, xq.shape[0], chunk):
        mu, v = model.predict_value(xq[s:s + chunk], include_noise=include_noise)
        mus.append(mu)
        vs.append(v)
    return np.vstack(mus), np.vstack(vs)
