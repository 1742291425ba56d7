"""Restatement of the GPy pieces the reference's hot path calls.  TEST INFRASTRUCTURE.

GPy is a third-party dependency of the reference that is NOT under
/root/reference and is not pinned (README.md:34 names it without a version;
Python-2.7 era, i.e. GPy <= 1.9.x).  The call sites this file serves:

  GPy.kern.RBF(...).K / .Kdiag      gpmodel_library.py:65,213,234,239,251,313,318,322
  GPy.util.linalg.pdinv             gpmodel_library.py:219,255,270
  GPy.util.diag.add                 gpmodel_library.py:218,254,269
  GPy.util.linalg.{jitchol,dpotri,dpotrs,dtrtrs,symmetrify,tdot}  :18, :400-466
  GPy.models.GPRegression           gpmodel_library.py:125,168,298 (+ .predict :102,
                                    .posterior_samples_f :131, .set_XY :128)

Published algorithm restated (GPy 1.9.x, kern/src/stationary.py, kern/src/rbf.py,
util/linalg.py, inference/latent_function_inference/exact_gaussian_inference.py,
inference/latent_function_inference/posterior.py):

  RBF:      r = sqrt(clip(|x|^2 + |y|^2 - 2 x.y, 0, inf)) / l   (ARD: inputs divided by l first)
            K(X)  forces the diagonal of r^2 to exactly 0
            K = variance * exp(-0.5 * r**2);  Kdiag = variance
  jitchol:  dpotrf(lower); on failure up to 5 retries with jitter mean(diag)*1e-6*10^i
  pdinv:    L = jitchol(A); Ai = dpotri(L) symmetrified; Li = dtrtri(L); logdet = 2 sum log diag L
  GPRegression (exact Gaussian inference): Ky = K + (noise + 1e-8) I; LW = chol(Ky);
            alpha = dpotrs(LW, Y); predict: mu = Kx^T alpha, var = Kdiag - sum(dtrtrs(LW,Kx)^2, 0),
            + noise when include_likelihood.

These are assumptions about an absent dependency; they cannot be verified in this
image.  Sensitivity is covered in tests/test_oracle.py.
"""
import types

import numpy as np
from scipy import linalg as sla
from scipy.linalg import lapack


# ----------------------------------------------------------------------------- linalg
def force_F_ordered(A):
    if A.flags['F_CONTIGUOUS']:
        return A
    return np.asfortranarray(A)


def symmetrify(A, upper=False):
    """In place: copy the lower triangle onto the upper (or the reverse)."""
    n = A.shape[0]
    iu = np.triu_indices(n, 1)
    if upper:
        A.T[iu] = A[iu]
    else:
        A[iu] = A.T[iu]
    return A


def jitchol(A, maxtries=5):
    A = np.ascontiguousarray(A)
    L, info = lapack.dpotrf(A, lower=1)
    if info == 0:
        return L
    diagA = np.diag(A)
    if np.any(diagA <= 0.):
        raise np.linalg.LinAlgError("not pd: non-positive diagonal elements")
    jitter = diagA.mean() * 1e-6
    num_tries = 1
    while num_tries <= maxtries and np.isfinite(jitter):
        try:
            return sla.cholesky(A + np.eye(A.shape[0]) * jitter, lower=True)
        except Exception:
            jitter *= 10
        finally:
            num_tries += 1
    raise np.linalg.LinAlgError("not positive definite, even with jitter.")


def dtrtri(L):
    L = force_F_ordered(L)
    return lapack.dtrtri(L, lower=1)[0]


def dtrtrs(A, B, lower=1, trans=0, unitdiag=0):
    A = np.asfortranarray(A)
    return lapack.dtrtrs(A, B, lower=lower, trans=trans, unitdiag=unitdiag)


def dpotrs(A, B, lower=1):
    A = force_F_ordered(A)
    return lapack.dpotrs(A, B, lower=lower)


def dpotri(A, lower=1):
    A = force_F_ordered(A)
    R, info = lapack.dpotri(A, lower=lower)
    symmetrify(R)
    return R, info


def pdinv(A, *args):
    L = jitchol(A, *args)
    logdet = 2. * np.sum(np.log(np.diag(L)))
    Li = dtrtri(L)
    Ai, _ = dpotri(L, lower=1)
    symmetrify(Ai)
    return Ai, L, Li, logdet


def tdot(mat):
    return np.dot(mat, mat.T)


def diag_add(A, b, offset=0):
    """GPy.util.diag.add: in place A[i,i] += b."""
    assert offset == 0
    n = A.shape[0]
    A.flat[::n + 1] += b
    return A


# ----------------------------------------------------------------------------- kernel
class RBF(object):
    def __init__(self, input_dim, variance=1., lengthscale=None, ARD=False, **kw):
        self.input_dim = input_dim
        self.ARD = ARD
        self.variance = float(np.asarray(variance).reshape(-1)[0])
        if lengthscale is None:
            lengthscale = 1.0
        ls = np.asarray(lengthscale, dtype=float).reshape(-1)
        if ARD:
            if ls.size == 1:
                ls = np.ones(input_dim) * ls[0]
            # GPy asserts lengthscale.size == input_dim for ARD kernels
            assert ls.size == input_dim, "bad number of lengthscales"
        else:
            assert ls.size == 1, "Only 1 lengthscale needed for non-ARD kernel"
        self.lengthscale = ls

    def _unscaled_dist(self, X, X2=None):
        if X2 is None:
            Xsq = np.sum(np.square(X), 1)
            r2 = -2. * tdot(X) + (Xsq[:, None] + Xsq[None, :])
            r2.flat[::r2.shape[0] + 1] = 0.
            r2 = np.clip(r2, 0, np.inf)
            return np.sqrt(r2)
        X1sq = np.sum(np.square(X), 1)
        X2sq = np.sum(np.square(X2), 1)
        r2 = -2. * np.dot(X, X2.T) + (X1sq[:, None] + X2sq[None, :])
        r2 = np.clip(r2, 0, np.inf)
        return np.sqrt(r2)

    def _scaled_dist(self, X, X2=None):
        if self.ARD:
            if X2 is not None:
                X2 = X2 / self.lengthscale
            return self._unscaled_dist(X / self.lengthscale, X2)
        return self._unscaled_dist(X, X2) / self.lengthscale

    def K(self, X, X2=None):
        r = self._scaled_dist(np.asarray(X, dtype=float), None if X2 is None else np.asarray(X2, dtype=float))
        return self.variance * np.exp(-0.5 * r ** 2)

    def Kdiag(self, X):
        ret = np.empty(X.shape[0])
        ret[:] = self.variance
        return ret


# ----------------------------------------------------------------------------- GPRegression
class GPRegression(object):
    """Exact-inference GP regression with a Gaussian likelihood (GPy default noise_var=1.0)."""

    def __init__(self, X, Y, kernel=None, noise_var=1., **kw):
        self.kern = kernel
        self.noise_var = float(noise_var)
        self.set_XY(X, Y)

    def set_XY(self, X=None, Y=None):
        self.X = np.asarray(X, dtype=float)
        self.Y = np.asarray(Y, dtype=float)
        K = self.kern.K(self.X)
        Ky = K.copy()
        diag_add(Ky, self.noise_var + 1e-8)
        Wi, LW, LWi, W_logdet = pdinv(Ky)
        alpha, _ = dpotrs(LW, self.Y, lower=1)
        self._woodbury_chol = LW
        self._woodbury_vector = alpha
        self._woodbury_inv = Wi

    def _raw_predict(self, Xnew, full_cov=False):
        Kx = self.kern.K(self.X, Xnew)
        mu = np.dot(Kx.T, self._woodbury_vector)
        if mu.ndim == 1:
            mu = mu.reshape(-1, 1)
        tmp = dtrtrs(self._woodbury_chol, Kx)[0]
        if full_cov:
            var = self.kern.K(Xnew) - tdot(tmp.T)
        else:
            var = (self.kern.Kdiag(Xnew) - np.square(tmp).sum(0))[:, None]
        return mu, var

    def predict(self, Xnew, full_cov=False, include_likelihood=True, **kw):
        mu, var = self._raw_predict(np.asarray(Xnew, dtype=float), full_cov=full_cov)
        if include_likelihood:
            if full_cov:
                var = var + np.eye(var.shape[0]) * self.noise_var
            else:
                var = var + self.noise_var
        return mu, var

    def posterior_samples_f(self, X, size=10, full_cov=True, **kw):
        m, v = self._raw_predict(np.asarray(X, dtype=float), full_cov=full_cov)
        if full_cov:
            return np.random.multivariate_normal(m.flatten(), v, size).T
        return np.random.multivariate_normal(m.flatten(), np.diag(v.flatten()), size).T


# ----------------------------------------------------------------------------- stub module tree
def as_gpy_module():
    """Build a module object shaped like the parts of ``GPy`` the reference imports
    (used only by tests/golden/make_golden.py to execute the reference source)."""
    GPy = types.ModuleType('GPy')
    kern = types.ModuleType('GPy.kern')
    kern.RBF = RBF

    def _no_periodic(*a, **k):
        raise NotImplementedError("StdPeriodic is out of scope (SURVEY.md section 8)")
    kern.StdPeriodic = _no_periodic
    util = types.ModuleType('GPy.util')
    linalg = types.ModuleType('GPy.util.linalg')
    for name in ('pdinv', 'dpotrs', 'dpotri', 'symmetrify', 'jitchol', 'dtrtrs', 'tdot', 'dtrtri'):
        setattr(linalg, name, globals()[name])
    diag = types.ModuleType('GPy.util.diag')
    diag.add = diag_add
    util.linalg = linalg
    util.diag = diag
    models = types.ModuleType('GPy.models')
    models.GPRegression = GPRegression
    inference = types.ModuleType('GPy.inference')
    lfi = types.ModuleType('GPy.inference.latent_function_inference')
    lfi.exact_gaussian_inference = types.ModuleType('GPy.inference.latent_function_inference.exact_gaussian_inference')
    inference.latent_function_inference = lfi
    GPy.kern, GPy.util, GPy.models, GPy.inference = kern, util, models, inference
    mods = {
        'GPy': GPy, 'GPy.kern': kern, 'GPy.util': util, 'GPy.util.linalg': linalg, 'GPy.util.diag': diag,
        'GPy.models': models, 'GPy.inference': inference,
        'GPy.inference.latent_function_inference': lfi,
        'GPy.inference.latent_function_inference.exact_gaussian_inference': lfi.exact_gaussian_inference,
    }
    return mods
