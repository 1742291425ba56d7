"""Drop-in reward / acquisition callables (reference aq_library.py), CUDA-backed.

Same call signatures as the reference -- f(time, xvals, robot_model, param=None[, FVECTOR=False]),
always usable with keyword arguments (mcts_library.py:387-396, robot_library.py:200-202) -- and the
same return types (numpy scalar, or an (M, 1) array with FVECTOR=True).  `robot_model` must be one of
this package's GPModel / OnlineGPModel objects: the posterior, the acquisition closed forms and the
per-path reduction run in one device pass through libipp_b200.so; there is no CPU path.

Batched forms used by the planners (one GPU call for a whole set of candidate paths):
  reward_paths(f_rew, time, paths, robot_model, param)  -> np.ndarray [P]

Quirks of the reference kept on purpose: UCB/EI scale by the variance (Q1), MES divides
by the variance and double-adds the utility (Q2), info_gain multiplies K by the variance again and
scales by 2*pi*e (Q6), RFF weights use std sqrt(1/lengthscale) (Q7).
"""
import ctypes
from functools import partial

import numpy as np
import scipy.optimize
import torch

from . import _lib as L
from .gpmodel_library import GPModel

_TWO_PI_E = 2 * np.pi * np.e


def _require_model(robot_model):
    if not isinstance(robot_model, GPModel):
        raise TypeError('robot_model must be an informative_path_planning_b200 GPModel/OnlineGPModel '
                        '(got %r); there is no CPU fallback' % type(robot_model))


def _queries(time, xvals, robot_model):
    """The reference's query construction (e.g. aq_library.py:84-91): columns 0,1 (+ time when D=3)."""
    data = np.array(xvals, dtype=float)
    x1 = data[:, 0]
    x2 = data[:, 1]
    if robot_model.dimension == 2:
        return data, np.vstack([x1, x2]).T
    return data, np.vstack([x1, x2, time * np.ones(len(x1))]).T


def ucb_beta(time):
    delta = 0.9
    d = 20
    pit = np.pi ** 2 * (time + 1) ** 2 / 6.
    return 2 * np.log(d * pit / delta)


_offsets_cache = {}


def _single_path_offsets(k, dev):
    """[0, k] as an int64 device tensor; cached (one tiny H2D per distinct k instead of one per call)."""
    key = (int(k), dev.index)
    t = _offsets_cache.get(key)
    if t is None:
        t = torch.tensor([0, k], dtype=torch.int64, device=dev)
        _offsets_cache[key] = t
    return t


def reward_device(f_rew, time, xvals, robot_model, param=None):
    """Scalar reward of one path as a 0-d CUDA tensor, WITHOUT a device->host sync (planner-internal:
    the tree only needs the value when the rollout is back-propagated).  Returns None when the
    reference's host-side early-outs apply (no data / no maxima), so the caller falls back to the
    public callable."""
    if robot_model.xvals is None:
        return None
    _, queries = _queries(time, xvals, robot_model)
    xq = L.to_device(queries)
    offs = _single_path_offsets(xq.shape[0], xq.device)
    if f_rew == 'mean':
        r, _ = robot_model.reward_paths_device(L.REWARD_UCB, xq, offs, params=[np.sqrt(ucb_beta(time))])
    elif f_rew == 'exp_improve':
        eta = 0.5 if param is None else sum(param) / len(param)
        r, _ = robot_model.reward_paths_device(L.REWARD_EI, xq, offs, params=[float(eta)])
    elif f_rew == 'mes':
        if param[0] is None:
            return None
        maxes_d = param[0] if isinstance(param[0], torch.Tensor) else L.to_device(np.asarray(param[0], dtype=float).reshape(-1))
        r, _ = robot_model.reward_paths_device(L.REWARD_MES, xq, offs, maxes_d=maxes_d)
    else:
        return None
    return r[0]


# ------------------------------------------------------------------------------------------ rewards
def mean_UCB(time, xvals, robot_model, param=None, FVECTOR=False):
    """reference aq_library.py:82-114"""
    _require_model(robot_model)
    data, queries = _queries(time, xvals, robot_model)
    if robot_model.xvals is None:
        return np.ones((data.shape[0], 1)) if FVECTOR else 1.0
    xq = L.to_device(queries)
    reward, points = robot_model.reward_paths_device(L.REWARD_UCB, xq, _single_path_offsets(xq.shape[0], xq.device),
                                                     params=[np.sqrt(ucb_beta(time))], want_points=FVECTOR)
    if FVECTOR:
        return points.cpu().numpy().reshape(-1, 1)
    return np.float64(reward.cpu().numpy()[0])


def exp_improvement(time, xvals, robot_model, param=None):
    """reference aq_library.py:556-582"""
    _require_model(robot_model)
    _, queries = _queries(time, xvals, robot_model)
    eta = 0.5 if param is None else sum(param) / len(param)
    xq = L.to_device(queries)
    reward, _ = robot_model.reward_paths_device(L.REWARD_EI, xq, _single_path_offsets(xq.shape[0], xq.device),
                                                params=[float(eta)])
    return np.float64(reward.cpu().numpy()[0])


def mves(time, xvals, robot_model, param, FVECTOR=False):
    """reference aq_library.py:281-337"""
    _require_model(robot_model)
    maxes = param[0]
    if maxes is None:
        return np.ones((xvals.shape[0], 1)) if FVECTOR else 1.0
    _, queries = _queries(time, xvals, robot_model)
    xq = L.to_device(queries)
    maxes_d = L.to_device(np.asarray(maxes, dtype=float).reshape(-1))
    reward, points = robot_model.reward_paths_device(L.REWARD_MES, xq, _single_path_offsets(xq.shape[0], xq.device),
                                                     maxes_d=maxes_d, want_points=True)
    if FVECTOR:
        # f_m = (sum_i u_mi + sum_i sum_m' u_m'i) / S  (the reference's "f += utility; f += sum(utility)")
        pts = points.cpu().numpy().reshape(-1, 1)
        return (pts + np.sum(pts)) / maxes_d.numel()
    return np.float64(reward.cpu().numpy()[0])


def _logdet_small_device(robot_model, queries, scale_variance):
    """logdet(I + K2(q, q)) for the no-data branch of info_gain, on the device (k x k Cholesky)."""
    v = float(np.asarray(robot_model.variance, dtype=float).reshape(-1)[0])
    kern2 = L.make_rbf(robot_model.dimension, v * scale_variance, robot_model.lengthscale)
    xq = L.to_device(queries)
    k = xq.shape[0]
    ld = L.round_up(k, 2)
    A = torch.empty((k, ld), dtype=torch.float64, device=xq.device)
    L.check(L.lib.ipp_rbf_cross(ctypes.byref(kern2), L.ptr(xq), k, L.ptr(xq), k, L.ptr(A), ld, L.stream_ptr()))
    A[:, :k] += torch.eye(k, dtype=torch.float64, device=xq.device)
    scal = torch.zeros(2, dtype=torch.float64, device=xq.device)
    info = torch.zeros(1, dtype=torch.int32, device=xq.device)
    work = L.workspace(L.lib.ipp_potrf_workspace(k), slot='potrf')
    L.check(L.lib.ipp_potrf_lower(L.ptr(A), k, ld, L.ptr(scal), L.ptr(info), L.ptr(work), L.stream_ptr()))
    if int(info.item()) != 0:
        raise np.linalg.LinAlgError('I + v K(q) is not positive definite')
    return float(scal[0].item())


IG_KMAX = 32          # csrc/infogain.cu: longest path the one-warp-per-path Schur kernel factors in shared memory


def _info_gain_long_path(robot_model, xq):
    """logdet S for one path of more than IG_KMAX points (the reference takes any length, aq_library.py:55-70): the same
    Schur block S = I + K2(q,q) - K2(q,X) (I + vK)^-1 K2(X,q), assembled with the FP64 GEMM and factored by the blocked
    device Cholesky."""
    kern2, binv, ld, _ = robot_model._info_gain_state()
    k, N, dev = xq.shape[0], robot_model.n_obs, xq.device
    ldn = L.round_up(N, 16)
    Ct = torch.zeros((k, ldn), dtype=torch.float64, device=dev)
    L.check(L.lib.ipp_rbf_cross(ctypes.byref(kern2), L.ptr(xq), k, L.ptr(robot_model._X_d), N, L.ptr(Ct), ldn, L.stream_ptr()))
    G = torch.empty((k, ldn), dtype=torch.float64, device=dev)
    L.check(L.lib.ipp_dgemm_nt(k, N, N, 1.0, L.ptr(Ct), ldn, L.ptr(binv), ld, 0.0, L.ptr(G), ldn, None, 0, L.stream_ptr()))
    lds = L.round_up(k, 2)
    S = torch.empty((k, lds), dtype=torch.float64, device=dev)
    L.check(L.lib.ipp_rbf_cross(ctypes.byref(kern2), L.ptr(xq), k, L.ptr(xq), k, L.ptr(S), lds, L.stream_ptr()))
    S[:, :k] += torch.eye(k, dtype=torch.float64, device=dev)
    L.check(L.lib.ipp_dgemm_nt(k, k, N, -1.0, L.ptr(G), ldn, L.ptr(Ct), ldn, 1.0, L.ptr(S), lds, None, 0, L.stream_ptr()))
    scal = torch.zeros(2, dtype=torch.float64, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    work = L.workspace(L.lib.ipp_potrf_workspace(k), slot='potrf')
    L.check(L.lib.ipp_potrf_lower(L.ptr(S), k, lds, L.ptr(scal), L.ptr(info), L.ptr(work), L.stream_ptr()))
    if int(info.item()) != 0:
        raise np.linalg.LinAlgError('info_gain: Schur block not positive definite')
    return float(scal[0].item())


def info_gain_paths(robot_model, queries, offsets):
    """2*pi*e * (logdet(I + v K([X; q_p])) - logdet(I + v K(X))) for every path p, as one device pass
    (Schur-complement form of reference aq_library.py:55-70); paths longer than IG_KMAX points one by one through
    _info_gain_long_path."""
    kern2, binv, ld, _ = robot_model._info_gain_state()
    xq = L.to_device(queries)
    N = robot_model.n_obs
    offsets = np.asarray(offsets, dtype=np.int64)
    P = offsets.shape[0] - 1
    lengths = np.diff(offsets)
    long_paths = np.nonzero(lengths > IG_KMAX)[0]
    result = np.empty(P)
    short = np.nonzero(lengths <= IG_KMAX)[0]
    if long_paths.size:
        for p in long_paths:
            result[p] = _info_gain_long_path(robot_model, xq[offsets[p]:offsets[p + 1]].contiguous())
        if short.size == 0:
            return _TWO_PI_E * result
        keep = np.concatenate([np.arange(offsets[p], offsets[p + 1]) for p in short])
        xq_s = xq[torch.as_tensor(keep, device=xq.device)].contiguous()
        offs_s = np.concatenate([[0], np.cumsum(lengths[short])])
    else:
        xq_s, offs_s = xq, offsets
    M, Ps = xq_s.shape[0], offs_s.shape[0] - 1
    offs = torch.as_tensor(np.asarray(offs_s, dtype=np.int64)).to(xq.device)
    out = torch.empty(Ps, dtype=torch.float64, device=xq.device)
    info = torch.zeros(1, dtype=torch.int32, device=xq.device)
    work = L.workspace(L.lib.ipp_info_gain_workspace(N, M))
    L.check(L.lib.ipp_info_gain_paths(ctypes.byref(kern2), L.ptr(robot_model._X_d), N, L.ptr(binv), ld, L.ptr(xq_s),
                                      L.ptr(offs), Ps, M, L.ptr(out), L.ptr(info), L.ptr(work), L.stream_ptr()))
    if int(info.item()) != 0:
        raise np.linalg.LinAlgError('info_gain: Schur block not positive definite')
    result[short] = out.cpu().numpy()
    return _TWO_PI_E * result


def info_gain(time, xvals, robot_model, param=None):
    """reference aq_library.py:34-80"""
    _require_model(robot_model)
    _, queries = _queries(time, xvals, robot_model)
    if robot_model.xvals is None:
        v = float(np.asarray(robot_model.variance, dtype=float).reshape(-1)[0])
        return np.float64(0.5 * _logdet_small_device(robot_model, queries, v))
    return np.float64(info_gain_paths(robot_model, queries, [0, queries.shape[0]])[0])


def hotspot_info_UCB(time, xvals, robot_model, param=None):
    """reference aq_library.py:117-135 (LAMBDA = 1).  The reference calls predict_value itself (not mean_UCB), so with
    no data it sees the prior (zeros, variance * ones): info_gain + sqrt(beta_t) * k * variance, not mean_UCB's 1.0."""
    if robot_model.xvals is None:
        k = np.array(xvals, dtype=float).shape[0]
        v = float(np.asarray(robot_model.variance, dtype=float).reshape(-1)[0])
        return info_gain(time, xvals, robot_model) + 0.0 + np.sqrt(ucb_beta(time)) * np.sum(np.fabs(np.ones((k, 1)) * v))
    return info_gain(time, xvals, robot_model) + mean_UCB(time, xvals, robot_model)


def naive(time, xvals, robot_model, param, FVECTOR=False):
    """reference aq_library.py:339-367 (pure geometry on a handful of points: host numpy, as in the reference)"""
    _, max_locs, _ = param[0]
    if max_locs is None:
        return np.zeros((xvals.shape[0], 1)) if FVECTOR else 0.0
    data = np.array(xvals, dtype=float)
    f = np.zeros((data.shape[0], 1))
    for i in range(max_locs.shape[0]):
        dist = np.sqrt(np.square(data[:, 0] - max_locs[i][0]) + np.square(data[:, 1] - max_locs[i][1]))
        f += (dist <= param[1]).astype(float).reshape(f.shape)
    f = f / max_locs.shape[0]
    return f if FVECTOR else np.sum(f)


def naive_value(time, xvals, robot_model, param, FVECTOR=False):
    """reference aq_library.py:369-413"""
    max_vals, _, funcs = param[0]
    if max_vals is None or funcs is None:
        return np.zeros((xvals.shape[0], 1)) if FVECTOR else 0.0
    data, queries = _queries(time, xvals, robot_model)
    f = np.zeros((data.shape[0], 1))
    for i in range(max_vals.shape[0]):
        mean = funcs[i](queries)
        f += (np.fabs(mean - max_vals[i][0]) <= param[1]).astype(float).reshape(f.shape)
    f = f / float(max_vals.shape[0])
    return f if FVECTOR else np.sum(f)


def naive_levels(f_rew, time, level_points, robot_model, param):
    """[naive / naive_value (time, x, robot_model, param) for x in level_points] -- the rewards of one tree-search rollout.
    Neither reward reads the belief (naive: distances to the sampled maximisers; naive_value: the sampled functions
    themselves), so a rollout needs none of its simulated belief updates.  naive_value evaluates each sampled function
    ONCE on all points of the rollout (one device call per function instead of one per function and level)."""
    if f_rew == 'naive':
        return [naive(time, x, robot_model, param) for x in level_points]
    max_vals, _, funcs = param[0]
    if max_vals is None or funcs is None:
        return [0.0 for _ in level_points]
    queries = [_queries(time, x, robot_model)[1] for x in level_points]
    offs = np.cumsum([0] + [q.shape[0] for q in queries])
    allq = np.vstack(queries)
    means = [funcs[i](allq) for i in range(max_vals.shape[0])]
    out = []
    for l in range(len(queries)):
        f = np.zeros((queries[l].shape[0], 1))
        for i in range(max_vals.shape[0]):
            mean = means[i][offs[l]:offs[l + 1]]
            f += (np.fabs(mean - max_vals[i][0]) <= param[1]).astype(float).reshape(f.shape)
        f = f / float(max_vals.shape[0])
        out.append(np.sum(f))
    return out


# ------------------------------------------------------------------------------------------ batched rewards
def flatten_paths(paths, dimension=2, time=0):
    """paths: sequence of per-path point lists/arrays (k_p, >=2) -> (queries (M, D), offsets (P+1,) int64)."""
    arrs = [np.asarray(p, dtype=float)[:, :2] for p in paths]
    offsets = np.zeros(len(arrs) + 1, dtype=np.int64)
    offsets[1:] = np.cumsum([a.shape[0] for a in arrs])
    q = np.vstack(arrs) if arrs else np.zeros((0, 2))
    if dimension == 3:
        q = np.hstack([q, time * np.ones((q.shape[0], 1))])
    return q, offsets


def reward_paths(f_rew, time, paths, robot_model, param=None, offsets=None):
    """All candidate paths in ONE device call.  f_rew in {'mean', 'exp_improve', 'mes', 'info_gain',
    'hotspot_info'}; `param` as the reference passes it to the matching callable.  `paths` is either a
    sequence of point lists, or an (M, D) array together with `offsets`.  Returns float64 [P]; equals
    [f(time, p, robot_model, param) for p in paths]."""
    _require_model(robot_model)
    if offsets is None:
        queries, offsets = flatten_paths(paths, robot_model.dimension, time)
    else:
        queries = np.asarray(paths, dtype=float)
        offsets = np.asarray(offsets, dtype=np.int64)
    P = offsets.shape[0] - 1
    if robot_model.xvals is None:
        if f_rew in ('mean',):
            return np.ones(P)
        if f_rew in ('info_gain', 'hotspot_info', 'exp_improve'):
            fn = {'info_gain': info_gain, 'exp_improve': exp_improvement, 'hotspot_info': hotspot_info_UCB}[f_rew]
            return np.array([fn(time, queries[offsets[p]:offsets[p + 1]], robot_model, param) for p in range(P)])
    if f_rew == 'mes' and param[0] is None:
        return np.ones(P)
    if f_rew in ('info_gain', 'hotspot_info'):
        ig = info_gain_paths(robot_model, queries, offsets)
        if f_rew == 'info_gain':
            return ig
    xq = L.to_device(queries)
    offs = torch.as_tensor(offsets).to(xq.device)
    if f_rew in ('mean', 'hotspot_info'):
        r, _ = robot_model.reward_paths_device(L.REWARD_UCB, xq, offs, params=[np.sqrt(ucb_beta(time))])
        r = r.cpu().numpy()
        return r + ig if f_rew == 'hotspot_info' else r
    if f_rew == 'exp_improve':
        eta = 0.5 if param is None else sum(param) / len(param)
        r, _ = robot_model.reward_paths_device(L.REWARD_EI, xq, offs, params=[float(eta)])
        return r.cpu().numpy()
    if f_rew == 'mes':
        maxes_d = L.to_device(np.asarray(param[0], dtype=float).reshape(-1))
        r, _ = robot_model.reward_paths_device(L.REWARD_MES, xq, offs, maxes_d=maxes_d)
        return r.cpu().numpy()
    raise ValueError('unknown reward %r' % (f_rew,))


# ------------------------------------------------------------------------------------------ max-value sampling
class RFFTarget(object):
    """One posterior function sample f(x) = theta^T sqrt(2 v / F) cos(W x^T + b) (reference general_target,
    aq_library.py:138-139).  Calling it evaluates the sample on the device; `scalar`/`gradient` are the
    closed forms SLSQP (host, as in the reference :527) needs for one point."""

    def __init__(self, variance, nFeatures, W, theta, b):
        self.variance, self.nFeatures = float(variance), int(nFeatures)
        self.W, self.theta, self.b = W, theta, b
        self.scale = np.sqrt(2.0 * self.variance / self.nFeatures)
        self._dev = None

    def _device_params(self):
        if self._dev is None:
            self._dev = (L.to_device(self.W), L.to_device(self.b.reshape(-1)),
                         L.to_device((self.theta.reshape(-1) * self.scale)))
        return self._dev

    def eval_device(self, x_d, want_values=True):
        """x_d: (G, D) CUDA tensor.  Returns (values[G] or None, max value, argmax index)."""
        W_d, b_d, th_d = self._device_params()
        G, D = x_d.shape
        vals = torch.empty(G, dtype=torch.float64, device=x_d.device) if want_values else None
        mx = torch.empty(1, dtype=torch.float64, device=x_d.device)
        am = torch.empty(1, dtype=torch.int64, device=x_d.device)
        work = L.workspace(L.lib.ipp_rff_workspace(self.nFeatures, 1, G, 0), slot='rff')
        L.check(L.lib.ipp_rff_eval_argmax(L.ptr(W_d), L.ptr(b_d), L.ptr(th_d), self.nFeatures, 1, D, 0, L.ptr(x_d), G,
                                          L.ptr(vals), L.ptr(mx), L.ptr(am), L.ptr(work), L.stream_ptr()))
        return vals, mx, am

    def __call__(self, x):
        x = np.asarray(x, dtype=float)
        vals, _, _ = self.eval_device(L.to_device(x))
        return vals.cpu().numpy().reshape(-1, 1)

    # closed forms for the host optimiser (one point at a time)
    def scalar(self, x):
        d = self.W.shape[1]
        return np.dot(self.theta.T * self.scale, np.cos(np.dot(self.W, x.reshape(1, d).T) + self.b)).T

    def gradient(self, x):
        d = self.W.shape[1]
        return np.dot(self.theta.T * -self.scale, np.sin(np.dot(self.W, x.reshape((d, 1))) + self.b) * self.W)


def general_target(x, robot_model, nFeatures, W, theta, b):
    """reference aq_library.py:138-139 (device evaluation)."""
    return RFFTarget(robot_model.variance, nFeatures, W, theta, b)(x)


def _rff_theta(Z, zvals, noise_var, eps, nFeatures):
    """Posterior weight draw (reference aq_library.py:177-206): small dense LAPACK on the host, as in the
    reference (eig / inv / cholesky of an N x N or F x F matrix)."""
    N = Z.shape[1]
    if N < nFeatures:
        Sigma = np.dot(Z.T, Z) + noise_var * np.eye(N)
        mu = np.dot(np.dot(Z, np.linalg.inv(Sigma)), zvals)
        # the reference calls the general solver (np.linalg.eig) and keeps the real parts; Sigma is symmetric and the result
        # below is the matrix function U diag(R) U^T of it, which does not depend on the eigenvector basis: the symmetric
        # solver gives the same theta to round-off at a sixth of the time (1.9 ms -> 0.3 ms for N = 100)
        D, U = np.linalg.eigh(Sigma)
        D = np.reshape(D, (D.shape[0], 1))
        R = np.reciprocal(np.sqrt(D) * (np.sqrt(D) + np.sqrt(noise_var)))
        return eps - np.dot(Z, np.dot(U, R * (np.dot(U.T, np.dot(Z.T, eps))))) + mu
    Sigma = np.dot(Z, Z.T) / noise_var + np.eye(nFeatures)
    Sigma = np.linalg.inv(Sigma)
    mu = np.dot(np.dot(Sigma, Z), zvals) / noise_var
    return mu + np.dot(np.linalg.cholesky(Sigma), eps)


def rff_features_device(W, b, X_d, variance, nFeatures):
    """Z = sqrt(2 v / F) cos(W X^T + b) on the device (reference aq_library.py:171) -> CUDA tensor (F, ld) with ld even."""
    F, D = W.shape
    N = X_d.shape[0]
    ldz = L.round_up(N, 2)
    Z = torch.zeros((F, ldz), dtype=torch.float64, device=X_d.device)
    W_d, b_d = L.to_device(W), L.to_device(b.reshape(-1))     # keep both alive until the launch is enqueued
    L.check(L.lib.ipp_rff_features(L.ptr(W_d), L.ptr(b_d), F, D, L.ptr(X_d), N,
                                   float(np.sqrt(2 * variance / nFeatures)), L.ptr(Z), ldz, L.stream_ptr()))
    return Z


def rff_features(W, b, X_d, variance, nFeatures):
    """rff_features_device -> numpy (F, N)."""
    return np.ascontiguousarray(rff_features_device(W, b, X_d, variance, nFeatures)[:, :X_d.shape[0]].cpu().numpy())


def _gemm_nt(A, B, M, N, K, alpha=1.0, C=None, beta=0.0):
    """C[M][ldc] = alpha A[M][K] B[N][K]^T + beta C through the library's FP64 DMMA GEMM (row-major CUDA tensors)."""
    ldc = L.round_up(N, 2)
    if C is None:
        C = torch.empty((M, ldc), dtype=torch.float64, device=A.device)
    L.check(L.lib.ipp_dgemm_nt(M, N, K, float(alpha), L.ptr(A), A.stride(0), L.ptr(B), B.stride(0), float(beta),
                               L.ptr(C), C.stride(0), None, 0, L.stream_ptr()))
    return C


def rff_theta_batch_device(robot_model, W, b, eps, variance):
    """All S posterior weight draws of the N >= F branch (reference aq_library.py:195-200) for ONE feature set in one device
    call (ipp_rff_theta_draw with S draws): Sigma = (Z Z^T / s2 + I)^-1, m = Sigma Z z / s2, Theta[:, s] = m +
    chol(Sigma) eps[:, s].  The reference redoes the F x F inverse and Cholesky for every sample although only eps
    changes; here one factorisation of the reversed system serves all of them (chol(Sigma) = U^-T, see ipp_b200.h).
    eps: (F, S) numpy.  Returns Theta as a device tensor [S][F] (unscaled)."""
    F, S = eps.shape
    N = robot_model.xvals.shape[0]
    dev = robot_model._X_d.device
    z_d = getattr(robot_model, '_z_d', None)
    if z_d is None:
        z_d = L.to_device(np.asarray(robot_model.zvals, dtype=float).reshape(-1))
    W_d, b_d = L.to_device(W), L.to_device(b.reshape(-1))
    eps_d = L.to_device(np.ascontiguousarray(eps.T))           # [S][F]
    theta = torch.empty((S, F), dtype=torch.float64, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    work = L.workspace(L.lib.ipp_rff_theta_workspace(F, N, S), slot='rff_theta')
    L.check(L.lib.ipp_rff_theta_draw(L.ptr(W_d), L.ptr(b_d), F, int(W_d.shape[1]), L.ptr(robot_model._X_d), L.ptr(z_d), N,
                                     float(variance), float(robot_model.noise), L.ptr(eps_d), S, L.ptr(theta), L.ptr(info),
                                     L.ptr(work), L.stream_ptr()))
    if int(info.item()) != 0:
        raise np.linalg.LinAlgError('Z Z^T / noise + I is not positive definite')
    return theta


def rff_theta_ul(Z_d, N, zvals, noise_var, eps, z_d=None):
    """One posterior weight draw of the N >= F branch (reference aq_library.py:195-200):
    Sigma = (Z Z^T / s2 + I)^-1, m = Sigma Z z / s2, theta = m + chol(Sigma) eps -- without forming Sigma or factoring it.
    With the UL factorisation A = Z Z^T / s2 + I = U U^T (U upper triangular: the Cholesky factor of A with rows and
    columns reversed, reversed back), Sigma = U^-T U^-1 and U^-T is lower triangular with a positive diagonal, i.e.
    chol(Sigma) = U^-T exactly.  So m and chol(Sigma) eps are three triangular solves.  The N-dependent work (the F x F x N
    product and Z z) runs on the device; the F x F factorisation and the solves are host LAPACK, as in the reference.
    eps: (F, 1).  Returns theta (F, 1)."""
    import scipy.linalg
    F = eps.shape[0]
    A_d = _gemm_nt(Z_d, Z_d, F, F, N, alpha=1.0 / noise_var)
    zrow = torch.zeros((1, Z_d.stride(0)), dtype=torch.float64, device=Z_d.device)
    zrow[0, :N] = z_d.reshape(-1)[:N] if z_d is not None else L.to_device(np.asarray(zvals, dtype=float).reshape(-1))
    Zz_d = _gemm_nt(Z_d, zrow, F, 1, N)
    host = torch.cat([A_d[:, :F].reshape(-1), Zz_d[:, 0]]).cpu().numpy()          # one transfer, one synchronisation
    A = host[:F * F].reshape(F, F) + np.eye(F)
    rhs = host[F * F:] / noise_var
    U = np.linalg.cholesky(A[::-1, ::-1])[::-1, ::-1]                              # A = U U^T
    w = scipy.linalg.solve_triangular(U, rhs, lower=False, check_finite=False)
    m = scipy.linalg.solve_triangular(U.T, w, lower=True, check_finite=False)
    y = scipy.linalg.solve_triangular(U.T, np.asarray(eps, dtype=float).reshape(-1), lower=True, check_finite=False)
    return (m + y).reshape(F, 1)


def rff_theta_device(robot_model, W_d, b_d, eps_d, variance, nFeatures):
    """One posterior weight draw of the N >= F branch (reference aq_library.py:195-200) entirely on the device, nothing
    read back:  Sigma = (Z Z^T / s2 + I)^-1,  theta = Sigma Z z / s2 + chol(Sigma) eps  (ipp_rff_theta_draw: the system
    is factored with the features in reverse order, which makes chol(Sigma) a by-product of the factorisation of A --
    rff_theta_ul does the same with host LAPACK).
    W_d (F, D), b_d (F), eps_d (F): device tensors.  Nothing here synchronises with the host.
    Returns (theta [F], info [1] int32: != 0 when A is not positive definite), both on the device."""
    F = int(nFeatures)
    N = robot_model.xvals.shape[0]
    dev = robot_model._X_d.device
    z_d = getattr(robot_model, '_z_d', None)
    if z_d is None:
        z_d = L.to_device(np.asarray(robot_model.zvals, dtype=float).reshape(-1))
    theta = torch.empty(F, dtype=torch.float64, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    work = L.workspace(L.lib.ipp_rff_theta_workspace(F, N, 1), slot='rff_theta')
    W_d, b_d, eps_d = W_d.contiguous(), b_d.reshape(-1).contiguous(), eps_d.reshape(-1).contiguous()
    L.check(L.lib.ipp_rff_theta_draw(L.ptr(W_d), L.ptr(b_d), F, int(W_d.shape[1]), L.ptr(robot_model._X_d), L.ptr(z_d), N,
                                     float(variance), float(robot_model.noise), L.ptr(eps_d), 1, L.ptr(theta), L.ptr(info),
                                     L.ptr(work), L.stream_ptr()))
    return theta, info


def _grid_stage_device(target, guesses_d, dim, time, ax, gridSize=300):
    """Grid stage of global_maximization (reference aq_library.py:455-479) for an RFFTarget, enqueued only: the 300 x 300
    meshgrid (+ the data points as extra guesses when dim == 2) is assembled on the device from the 600 drawn coordinates
    (ax = [x1; x2] on the device; torch indexing = plumbing); row g = (x1[g % 300], x2[g // 300]) as np.meshgrid 'xy'
    ravels.  Returns (values, argmax)."""
    G0 = gridSize * gridSize
    cols = [ax[:gridSize].repeat(gridSize), ax[gridSize:].repeat_interleave(gridSize)]
    if dim != 2:
        cols.append(torch.full((G0,), float(time), dtype=torch.float64, device=ax.device))
    grid_d = torch.stack(cols, dim=1)
    if dim == 2:
        grid_d = torch.cat([grid_d, guesses_d[:, :2]], dim=0)
    vals, _, am = target.eval_device(grid_d.contiguous(), want_values=True)
    return vals, am


def global_maximization(target, target_vector_n, target_grad, target_vector_gradient_n, ranges, guesses, visualize,
                        filename, obstacles, f_rew, time=None, guesses_d=None, staged=None):
    """reference aq_library.py:444-553 without the plotting; the grid stage (:475-485) runs on the device
    when `target` is an RFFTarget.  `staged` = (x1, x2, values, argmax index, argmax index over the mesh points alone): the
    grid draws were already made (in the reference's order) and the grid stage already ran (sample_max_vals' pipelined form)."""
    gridSize = 300
    bb = ((ranges[1] - ranges[0]) * 0.10, (ranges[3] - ranges[2]) * 0.10)
    ranges = (ranges[0] + bb[0], ranges[1] - bb[0], ranges[2] + bb[1], ranges[3] - bb[1])
    dim = guesses.shape[1]
    am_mesh = None
    if staged is not None:
        x1, x2, vals, am_index, am_mesh = staged
    else:
        x1 = np.random.uniform(ranges[0], ranges[1], size=gridSize)
        x2 = np.random.uniform(ranges[2], ranges[3], size=gridSize)
    if isinstance(target, RFFTarget):
        G0 = gridSize * gridSize
        if staged is None:
            guesses_d = guesses_d if guesses_d is not None else L.to_device(guesses)
            vals, am = _grid_stage_device(target, guesses_d, dim, time, L.to_device(np.concatenate([x1, x2])), gridSize)
            am_index = int(am.item())

        def point(g):
            if g < G0:
                return np.array([x1[g % gridSize], x2[g // gridSize]] + ([float(time)] if dim != 2 else []))
            return np.asarray(guesses[g - G0, :], dtype=float)
        start = point(am_index)
        if start[0] < ranges[0] or start[0] > ranges[1] or start[1] < ranges[2] or start[1] > ranges[3]:
            start = point(am_mesh if am_mesh is not None else int(torch.argmax(vals[:G0]).item()))
    else:
        x1, x2 = np.meshgrid(x1, x2, sparse=False, indexing='xy')
        if dim == 2:
            Xgrid_sample = np.vstack([x1.ravel(), x2.ravel()]).T
            Xgrid = np.vstack([Xgrid_sample, guesses])
        else:
            Xgrid_sample = np.vstack([x1.ravel(), x2.ravel(), time * np.ones(len(x1.ravel()))]).T
            Xgrid = Xgrid_sample
        y = target(Xgrid)
        start = np.asarray(Xgrid[np.argmax(y), :])
        if start[0] < ranges[0] or start[0] > ranges[1] or start[1] < ranges[2] or start[1] > ranges[3]:
            start = np.asarray(Xgrid_sample[np.argmax(target(Xgrid_sample)), :])
    if dim == 2:
        bounds = ((ranges[0], ranges[1]), (ranges[2], ranges[3]))
    else:
        bounds = ((ranges[0], ranges[1]), (ranges[2], ranges[3]), (time, time))
    res = scipy.optimize.minimize(fun=target_vector_n, x0=start, method='SLSQP', jac=target_vector_gradient_n,
                                  bounds=bounds)
    if not res['success']:
        return 0, 0, 0, False
    return res['x'], -res['fun'], res['jac'], True


PIPELINE_MAX_VALUE_SAMPLES = True
_mv_buffers = {}


def _mv_stage(i, n_up, n_down):
    """Pinned staging of pipelined sample i on the current device: (upload buffer, download buffer), reused from call to
    call (pinning a fresh buffer costs more than the whole sample).  The upload buffer is only handed out again once the
    copy that last read it has run (_mv_guard's event)."""
    key = (L.device().index, i, n_up, n_down)
    st = _mv_buffers.get(key)
    if st is None:
        st = _mv_buffers[key] = {'up': torch.empty(n_up, dtype=torch.float64).pin_memory(),
                                 'down': torch.empty(n_down, dtype=torch.float64).pin_memory(), 'event': None, 'key': key}
    if st['event'] is not None:
        st['event'].synchronize()
    _mv_buffers['last', L.device().index, i] = st
    return st['up'], st['down']


def _mv_guard(i):
    st = _mv_buffers['last', L.device().index, i]
    st['event'] = torch.cuda.Event()
    st['event'].record()


def sample_max_vals(robot_model, t, nK=3, nFeatures=200, visualize=False, obstacles=None, f_rew='mes'):
    """reference aq_library.py:141-279.  Random draws come from the global np.random in the reference's
    order (W, b, eps, then the grid draws of each optimisation attempt).

    Pipelined form (N >= F, the usual case): the reference's samples only interact through the random stream, and the
    stream's order is known as long as every weight draw succeeds and every optimisation succeeds at its FIRST attempt.
    So all draws of all nK samples are made up front in exactly that order, the device work of every sample (features,
    F x F x N product, factorisation, weight draw, 300 x 300 grid stage, arg-max) is enqueued without a single
    synchronisation, and the host then polishes sample i with SLSQP while the device is still busy with i + 1.  If a
    sample does break the assumption (not positive definite, or SLSQP fails and the reference would redraw the grid), the
    stream is put back to where the reference's would be at that point and the rest runs one sample at a time."""
    _require_model(robot_model)
    if robot_model.xvals is None:
        return None, None, None
    d = robot_model.xvals.shape[1]
    samples = np.zeros((nK, 1))
    locs = np.zeros((nK, d))
    funcs = []
    delete_locs = []
    variance = float(np.asarray(robot_model.variance, dtype=float).reshape(-1)[0])
    N_obs = robot_model.xvals.shape[0]
    ls = robot_model.lengthscale if robot_model.dimension == 2 else robot_model.lengthscale[0]
    gridSize = 300
    rg = robot_model.ranges
    bb = ((rg[1] - rg[0]) * 0.10, (rg[3] - rg[2]) * 0.10)
    inner = (rg[0] + bb[0], rg[1] - bb[0], rg[2] + bb[1], rg[3] - bb[1])

    def draw_weights():
        W = np.random.normal(loc=0.0, scale=np.sqrt(1. / ls), size=(nFeatures, d))
        b = 2 * np.pi * np.random.uniform(low=0.0, high=1.0, size=(nFeatures, 1))
        eps = np.random.normal(loc=0.0, scale=1.0, size=(nFeatures, 1))
        return W, b, eps

    def make_target(W, b, theta, W_d, b_d, theta_d):
        target = RFFTarget(variance, nFeatures, W, theta, b)
        target._dev = (W_d, b_d, theta_d if theta_d is not None else L.to_device(theta.reshape(-1) * target.scale))
        return target

    def polish(i, target, staged=None, count=0):
        """The optimisation attempts of sample i (reference :237-262); `staged` serves the first one."""
        target_vector_n = lambda x, tg=target: -tg.scalar(x)
        target_vector_gradient_n = lambda x, tg=target: -np.asarray(tg.gradient(x).reshape(d, ))
        status = False
        while status is False and count < 5:
            maxima, max_val, _, status = global_maximization(target, target_vector_n, target.gradient,
                                                             target_vector_gradient_n, robot_model.ranges,
                                                             robot_model.xvals, visualize, 't%s.nK%s' % (t, i),
                                                             obstacles, f_rew=f_rew, time=t,
                                                             guesses_d=robot_model._X_d, staged=staged)
            staged = None
            count += 1
        if status is False:
            delete_locs.append(i)
            return
        samples[i] = np.array(max_val).reshape((1, 1))
        funcs.append(target)
        locs[i, :] = maxima.reshape((1, d))

    def one_sample(i):
        """Sample i start to finish, one synchronisation per stage (the reference's own order of events)."""
        W, b, eps = draw_weights()
        W_d, b_d = L.to_device(W), L.to_device(b.reshape(-1))           # uploaded once: features now, the sampled function later
        try:
            if N_obs >= nFeatures:      # (:195-200) on the device
                theta_d, info = rff_theta_device(robot_model, W_d, b_d, L.to_device(eps.reshape(-1)), variance, nFeatures)
                if int(info.item()) != 0:
                    raise np.linalg.LinAlgError('Z Z^T / noise + I is not positive definite')
                theta = theta_d.cpu().numpy().reshape(nFeatures, 1)
            else:                       # N x N eigendecomposition (:177-187): small dense LAPACK on the host
                Z_d = rff_features_device(W_d, b_d, robot_model._X_d, variance, nFeatures)
                Z = np.ascontiguousarray(Z_d[:, :N_obs].cpu().numpy())
                theta = _rff_theta(Z, robot_model.zvals, robot_model.noise, eps, nFeatures)
        except Exception:
            delete_locs.append(i)
            return
        polish(i, make_target(W, b, theta, W_d, b_d, None))

    first = 0
    if PIPELINE_MAX_VALUE_SAMPLES and N_obs >= nFeatures and nK > 1:
        scale = np.sqrt(2.0 * variance / nFeatures)
        staged = []
        for i in range(nK):             # every draw of every sample, in the reference's order; device work enqueued only
            W, b, eps = draw_weights()
            mid_state = np.random.get_state()
            x1 = np.random.uniform(inner[0], inner[1], size=gridSize)
            x2 = np.random.uniform(inner[2], inner[3], size=gridSize)
            end_state = np.random.get_state()
            # one pinned buffer up (W, b, eps, x1, x2), one down (theta, arg-max, info): no pageable copy, which would
            # wait for everything enqueued so far
            up, host = _mv_stage(i, nFeatures * (d + 2) + 2 * gridSize, nFeatures + 3)
            up_np = up.numpy()
            up_np[:nFeatures * d] = W.reshape(-1)
            up_np[nFeatures * d:nFeatures * (d + 1)] = b.reshape(-1)
            up_np[nFeatures * (d + 1):nFeatures * (d + 2)] = eps.reshape(-1)
            up_np[nFeatures * (d + 2):nFeatures * (d + 2) + gridSize] = x1
            up_np[nFeatures * (d + 2) + gridSize:] = x2
            dev_in = up.to(robot_model._X_d.device, non_blocking=True)
            _mv_guard(i)
            W_d = dev_in[:nFeatures * d].view(nFeatures, d)
            b_d = dev_in[nFeatures * d:nFeatures * (d + 1)]
            theta_d, info = rff_theta_device(robot_model, W_d, b_d, dev_in[nFeatures * (d + 1):nFeatures * (d + 2)],
                                             variance, nFeatures)
            target = make_target(W, b, None, W_d, b_d, theta_d * scale)
            vals, am = _grid_stage_device(target, robot_model._X_d, d, t, dev_in[nFeatures * (d + 2):], gridSize)
            am_mesh = torch.argmax(vals[:gridSize * gridSize]).reshape(1)        # (:481-485: the fall-back start point)
            host.copy_(torch.cat([theta_d, am.to(torch.float64), info.to(torch.float64), am_mesh.to(torch.float64)]),
                       non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            staged.append((target, x1, x2, vals, host, ev, mid_state, end_state))
        for i, (target, x1, x2, vals, host, ev, mid_state, end_state) in enumerate(staged):
            first = i + 1
            ev.synchronize()
            h = host.numpy()
            if int(h[nFeatures + 1]) != 0:          # the weight draw failed: the reference never drew this sample's grid
                delete_locs.append(i)
                np.random.set_state(mid_state)
                break
            target.theta = h[:nFeatures].reshape(nFeatures, 1).copy()
            before = len(delete_locs), len(funcs)
            target_vector_n = lambda x, tg=target: -tg.scalar(x)
            target_vector_gradient_n = lambda x, tg=target: -np.asarray(tg.gradient(x).reshape(d, ))
            maxima, max_val, _, status = global_maximization(target, target_vector_n, target.gradient,
                                                             target_vector_gradient_n, robot_model.ranges,
                                                             robot_model.xvals, visualize, 't%s.nK%s' % (t, i), obstacles,
                                                             f_rew=f_rew, time=t, guesses_d=robot_model._X_d,
                                                             staged=(x1, x2, vals, int(h[nFeatures]), int(h[nFeatures + 2])))
            if status is False:                     # the reference redraws the grid from here: later draws were speculative
                np.random.set_state(end_state)
                polish(i, target, count=1)
                break
            samples[i] = np.array(max_val).reshape((1, 1))
            funcs.append(target)
            locs[i, :] = maxima.reshape((1, d))
    for i in range(first, nK):
        one_sample(i)
    samples = np.delete(samples, delete_locs, axis=0)
    locs = np.delete(locs, delete_locs, axis=0)
    if len(delete_locs) == nK:
        samples[0] = np.max(robot_model.zvals) + 5.0 * np.sqrt(robot_model.noise)
        locs[0, :] = robot_model.xvals[np.argmax(robot_model.zvals)]
    return samples, locs, funcs


def _theta_rows(Theta, scale):
    """[S][F] device operand of the grid stage from Theta: (F, S) numpy or an [S][F] device tensor (unscaled)."""
    if isinstance(Theta, torch.Tensor):
        return (Theta * scale).contiguous(), int(Theta.shape[0])
    return L.to_device(np.ascontiguousarray(Theta.T) * scale), int(Theta.shape[1])


def rff_eval_argmax_mesh(W, b, Theta, variance, xs_d, ys_d, tval=0.0):
    """rff_eval_argmax_shared on the meshgrid of xs and ys (grid point g = j * nx + i is (xs[i], ys[j] [, tval]), i.e.
    np.meshgrid(xs, ys, indexing='xy') ravelled -- the grid of the reference's global_maximization, aq_library.py:455-466):
    the features factor as cos a_i cos c_j - sin a_i sin c_j (ipp_rff_eval_argmax_mesh).  Needs >= 8 samples."""
    F, D = W.shape
    nx, ny = int(xs_d.numel()), int(ys_d.numel())
    th_d, S = _theta_rows(Theta, np.sqrt(2.0 * variance / F))           # [S][F]
    mx = torch.empty(S, dtype=torch.float64, device=xs_d.device)
    am = torch.empty(S, dtype=torch.int64, device=xs_d.device)
    work = L.workspace(L.lib.ipp_rff_mesh_workspace(F, S, nx, ny), slot='rff')
    W_d, b_d = L.to_device(W), L.to_device(b.reshape(-1))     # keep both alive until the launch is enqueued
    L.check(L.lib.ipp_rff_eval_argmax_mesh(L.ptr(W_d), L.ptr(b_d), L.ptr(th_d), F, S, D, L.ptr(xs_d), nx, L.ptr(ys_d), ny,
                                           float(tval), None, L.ptr(mx), L.ptr(am), L.ptr(work), L.stream_ptr()))
    return mx, am


def shared_feature_draws(robot_model, nK, nFeatures, rng):
    """One feature set (W, b) and nK posterior weight vectors for it (reference aq_library.py:156-206 with the features
    shared by all samples): Theta is an [nK][F] device tensor when N >= F (one factorisation for all draws), else an
    (F, nK) numpy array from the host branch."""
    d = robot_model.xvals.shape[1]
    variance = float(np.asarray(robot_model.variance, dtype=float).reshape(-1)[0])
    ls = robot_model.lengthscale if robot_model.dimension == 2 else robot_model.lengthscale[0]
    W = rng.normal(0.0, np.sqrt(1. / ls), size=(nFeatures, d))
    b = 2 * np.pi * rng.uniform(0.0, 1.0, size=(nFeatures, 1))
    N = robot_model.xvals.shape[0]
    eps = np.hstack([rng.normal(size=(nFeatures, 1)) for _ in range(nK)])      # same draws, same order as one-by-one
    if N >= nFeatures:
        Theta = rff_theta_batch_device(robot_model, W, b, eps, variance)       # [S][F] on the device: stays there for the grid stage
    else:
        Z = rff_features(W, b, robot_model._X_d, variance, nFeatures)
        Theta = np.hstack([_rff_theta(Z, robot_model.zvals, robot_model.noise, eps[:, s:s + 1], nFeatures)
                           for s in range(nK)])                                # (F, nK)
    return W, b, Theta


def sample_max_vals_shared(robot_model, grid, nK=100, nFeatures=1000, rng=None):
    """Batched max-value sampling with ONE feature set shared by all nK posterior samples:
    phi(grid) is evaluated once per grid point and contracted against all nK weight vectors in the same
    kernel, followed by the per-sample arg-max.  Statistically the reference's sampler with a different
    RNG stream; used for BASELINE config 4 (F=1000, S=100, 1M-point grid).
    `grid`: (G, D) points, or a tuple (xs, ys) standing for their meshgrid (2-D beliefs; the factored form of the features).
    Returns (max_vals (nK,1), max_locs (nK,D), (W, b, Theta))."""
    _require_model(robot_model)
    rng = np.random.default_rng() if rng is None else rng
    variance = float(np.asarray(robot_model.variance, dtype=float).reshape(-1)[0])
    W, b, Theta = shared_feature_draws(robot_model, nK, nFeatures, rng)
    if isinstance(grid, tuple):                                  # (xs, ys): the meshgrid of two coordinate vectors, never built
        xs_d, ys_d = (g if isinstance(g, torch.Tensor) else L.to_device(g) for g in grid)
        mx, am = rff_eval_argmax_mesh(W, b, Theta, variance, xs_d, ys_d)
        nx = int(xs_d.numel())
        locs_d = torch.stack([xs_d[am % nx], ys_d[am // nx]], dim=1)
    else:
        grid_d = grid if isinstance(grid, torch.Tensor) else L.to_device(grid)
        mx, am = rff_eval_argmax_shared(W, b, Theta, variance, grid_d)
        locs_d = grid_d[am]
    if isinstance(Theta, torch.Tensor):
        Theta = np.ascontiguousarray(Theta.cpu().numpy().T)        # (F, nK), the layout the host-side branch returns
    return mx.cpu().numpy().reshape(-1, 1), locs_d.cpu().numpy(), (W, b, Theta)


def rff_eval_argmax_shared(W, b, Theta, variance, grid_d):
    """max_g / argmax_g of f_s(x_g) for all samples s with shared features; device tensors out."""
    F, D = W.shape
    G = grid_d.shape[0]
    th_d, S = _theta_rows(Theta, np.sqrt(2.0 * variance / F))           # [S][F]
    mx = torch.empty(S, dtype=torch.float64, device=grid_d.device)
    am = torch.empty(S, dtype=torch.int64, device=grid_d.device)
    work = L.workspace(L.lib.ipp_rff_workspace(F, S, G, 1), slot='rff')
    W_d, b_d = L.to_device(W), L.to_device(b.reshape(-1))     # keep both alive until the launch is enqueued
    L.check(L.lib.ipp_rff_eval_argmax(L.ptr(W_d), L.ptr(b_d), L.ptr(th_d), F, S, D, 1,
                                      L.ptr(grid_d), G, None, L.ptr(mx), L.ptr(am), L.ptr(work), L.stream_ptr()))
    return mx, am
