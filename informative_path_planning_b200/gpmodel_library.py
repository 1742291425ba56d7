"""Drop-in belief models backed by libipp_b200.so (sm_100a CUDA).

Mirrors the public interface of the reference's gpmodel_library.py:
  GPModel        (reference :25-185)   GPy-wrapping model: full refactor on every add_data (:124
                                       "if self.model == None or True"), Cholesky-form predict
  OnlineGPModel  (reference :187-466)  explicit (K + s2 I)^-1, block-inverse append, numpy-facing predict
Same constructor arguments, methods, attributes (.xvals .zvals .variance .lengthscale .noise .ranges
.dimension .kern .model) and error behaviour (ValueError on configuration, AssertionError on shapes,
numpy.linalg.LinAlgError when the factorisation fails even with GPy's jitter retries).

What differs is where the arithmetic runs: numpy in / numpy out, but K(.,.), the FP64 factorisation,
the block append and the posterior mean/variance are CUDA kernels reached through the C ABI
(include/ipp_b200.h).  The device state of a model is immutable once built (every append writes
new buffers), which makes copy.copy() the O(1) per-rollout fork mcts_library.py:694 relies on.

Extra, batched entry points used by the planners (not in the reference):
  predict_device(xq)                      device tensors in / out, no host round trip
  reward_paths(kind, points, offsets...)  GP predict + acquisition + per-path reduction in one call
"""
import copy
import ctypes

import numpy as np
import torch

from . import _lib as L

_TWO_PI_E = 2 * np.pi * np.e


class _RBFKern(object):
    """The `.kern` attribute: GPy.kern.RBF's K / Kdiag surface (aq_library.py:49,58,59 call it)."""

    def __init__(self, input_dim, variance, lengthscale, ARD):
        self.input_dim = input_dim
        self.ARD = ARD
        self.variance = variance
        self.lengthscale = lengthscale
        self._c = L.make_rbf(input_dim, variance, lengthscale)

    def K_device(self, X1_d, X2_d):
        n1, n2 = X1_d.shape[0], X2_d.shape[0]
        ld = L.round_up(n2, 2)
        out = torch.empty((n1, ld), dtype=torch.float64, device=X1_d.device)
        L.check(L.lib.ipp_rbf_cross(ctypes.byref(self._c), L.ptr(X1_d), n1, L.ptr(X2_d), n2, L.ptr(out), ld,
                                    L.stream_ptr()))
        return out[:, :n2]

    def K(self, X, X2=None):
        X1_d = L.to_device(X)
        X2_d = X1_d if X2 is None else L.to_device(X2)
        return self.K_device(X1_d, X2_d).cpu().numpy()

    def Kdiag(self, X):
        ret = np.empty(X.shape[0])
        ret[:] = self.variance
        return ret

    # GPy parameter-array view: kern[:] = [variance, lengthscale(s)]  (gpmodel_library.py:144, 180)
    def _params(self):
        ls = np.asarray(self.lengthscale, dtype=float).reshape(-1)
        n_ls = self.input_dim if self.ARD else 1
        return np.concatenate([[float(np.asarray(self.variance, dtype=float).reshape(-1)[0])], ls[:n_ls]])

    def __getitem__(self, key):
        return self._params()[key]

    def __setitem__(self, key, value):
        p = self._params()
        p[key] = np.asarray(value, dtype=float).reshape(-1) if np.ndim(value) else float(value)
        self.variance = float(p[0])
        self.lengthscale = tuple(float(v) for v in p[1:]) if self.ARD else float(p[1])
        self._c = L.make_rbf(self.input_dim, self.variance, self.lengthscale)


class _LegacyModelHandle(object):
    """Stands in for the GPy GPRegression object the reference keeps in `.model`: callers only
    test it against None (mcts_library.py:222,418,527,598) or call predict / posterior_samples_f."""

    def __init__(self, owner):
        self._owner = owner

    def predict(self, Xnew, full_cov=False, include_likelihood=True):
        return self._owner._predict_chol(Xnew, include_likelihood, full_cov)

    def posterior_samples_f(self, X, size=10, full_cov=True):
        m, v = self._owner._predict_chol(X, False, full_cov)
        if full_cov:
            return np.random.multivariate_normal(m.flatten(), v, size).T
        return np.random.multivariate_normal(m.flatten(), np.diag(v.flatten()), size).T


class GPModel(object):
    """Reference gpmodel_library.py:25-185."""

    def __init__(self, ranges, lengthscale, variance, noise=0.0001, dimension=2, kernel='rbf', period=None):
        self.noise = noise
        self.lengthscale = lengthscale
        self.variance = variance
        self.ranges = ranges
        self.xvals = None
        self.zvals = None
        if dimension == 2:
            self.dimension = dimension
            self.asymmetric = False
        elif dimension == 3:
            if len(lengthscale) < dimension:
                raise ValueError('Lengthscale vector must have same length as dimension.')
            self.dimension = dimension
            self.asymmetric = True
        else:
            raise ValueError('Environment must have dimension 2 or 3')
        if kernel == 'rbf':
            self.kern = _RBFKern(self.dimension, variance, lengthscale, self.asymmetric)
        elif kernel == 'rbf-period':
            raise ValueError("Kernel 'rbf-period' (GPy StdPeriodic + RBF) is outside the accelerated hot path")
        else:
            raise ValueError('Kernel type must by \'rbf\'')
        self.model = None
        # device state (immutable once built; replaced wholesale by add_data)
        self._X_d = None
        self._z_d = None
        self._alpha_d = None
        self._winv_d = None        # [N][ld] (K + (noise+1e-8) I)^-1
        self._linv_d = None        # [N][ld] L^-1 (GPy-form models only)
        self._ld = 0
        self._ig = None            # cached (binv, logdet_before) for info_gain
        self._i8 = None            # cached (planes, ld, scale) int8 slices of L^-1 for the sliced-INT8 predict
        self.predict_precision = 'fp64'   # 'i8x7' / 'i8x6' / 'i8x5': batched predicts run on the tcgen05 INT8 path (see predict_device)
        # Rollout forks may defer the "is the Schur block positive definite" read-back (a device->host
        # sync per append) to one check per rollout: see defer_checks() / check_pending().
        self._defer = False
        self._pending_info = []

    # ------------------------------------------------------------------ helpers
    @property
    def n_obs(self):
        return 0 if self.xvals is None else int(self.xvals.shape[0])

    def _noise_var(self):
        """Likelihood variance the regression actually uses (the legacy ipp_library model overrides it: GPy default 1.0)."""
        return float(self.noise)

    def _diag_add(self):
        return self._noise_var() + 1e-8

    def _factor(self, want_linv):
        """Full FP64 refresh on the device, with GPy jitchol's retry ladder (linalg.py: jitter =
        mean(diag) * 1e-6, x10 per retry, 5 retries, then LinAlgError)."""
        N = self.n_obs
        dev = L.device()
        self._X_d = L.to_device(self.xvals)
        self._z_d = L.to_device(self.zvals.reshape(-1))
        ld = L.round_up(N, 16)
        winv = torch.empty((N, ld), dtype=torch.float64, device=dev)
        linv = torch.empty((N, ld), dtype=torch.float64, device=dev) if want_linv else None
        alpha = torch.empty(N, dtype=torch.float64, device=dev)
        scal = torch.zeros(2, dtype=torch.float64, device=dev)      # [logdet, (info as int32 view)]
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        work = L.workspace(L.lib.ipp_gp_factor_workspace(N))
        diag_add = self._diag_add()
        jitter, tries = 0.0, 0
        while True:
            L.check(L.lib.ipp_gp_factor(ctypes.byref(self.kern._c), L.ptr(self._X_d), L.ptr(self._z_d), N,
                                        diag_add + jitter, L.ptr(winv), L.ptr(linv), ld, L.ptr(alpha), L.ptr(scal),
                                        L.ptr(info), L.ptr(work), L.stream_ptr()))
            if int(info.item()) == 0:
                break
            diag_mean = float(np.asarray(self.variance, dtype=float).reshape(-1)[0]) + diag_add
            if tries >= 5 or not np.isfinite(diag_mean) or diag_mean <= 0:
                raise np.linalg.LinAlgError('not positive definite, even with jitter.')
            jitter = diag_mean * 1e-6 if tries == 0 else jitter * 10
            tries += 1
        self._winv_d, self._linv_d, self._alpha_d, self._ld = winv, linv, alpha, ld
        self._logdet = scal
        self._ig = None
        self._i8 = None

    def _predict_core(self, xq_d, include_noise, mode, mat):
        M = xq_d.shape[0]
        dev = xq_d.device
        mean = torch.empty(M, dtype=torch.float64, device=dev)
        var = torch.empty(M, dtype=torch.float64, device=dev)
        N = self.n_obs
        work = L.workspace(L.lib.ipp_gp_predict_workspace(N, M))
        L.check(L.lib.ipp_gp_predict(ctypes.byref(self.kern._c), L.ptr(self._X_d), N, L.ptr(self._alpha_d),
                                     L.ptr(mat), self._ld, mode, self._noise_var() if include_noise else 0.0,
                                     L.ptr(xq_d), M, L.ptr(mean), L.ptr(var), L.ptr(work), L.stream_ptr()))
        return mean, var

    def _check_queries(self, xvals):
        assert xvals.shape[0] >= 1
        assert xvals.shape[1] == self.dimension

    def _predict_chol(self, xvals, include_noise, full_cov=False):
        xvals = np.asarray(xvals, dtype=float)
        if full_cov:
            return self._predict_full_cov(xvals, include_noise, self._noise_var(), diagonal_only=True)
        mean, var = self._predict_core(L.to_device(xvals), include_noise, L.VAR_CHOL, self._linv_d)
        return mean.cpu().numpy().reshape(-1, 1), var.cpu().numpy().reshape(-1, 1)

    def _predict_full_cov(self, xvals, include_noise, noise, diagonal_only):
        """K(q,q) - Kx^T W^-1 Kx on the device.  diagonal_only: GPy adds the likelihood variance on the
        diagonal; OnlineGPModel (reference :326-327) adds the scalar to EVERY entry."""
        xq_d = L.to_device(xvals)
        M, N = xq_d.shape[0], self.n_obs
        ldc = L.round_up(M, 2)
        mean = torch.empty(M, dtype=torch.float64, device=xq_d.device)
        cov = torch.empty((M, ldc), dtype=torch.float64, device=xq_d.device)
        work = L.workspace(L.lib.ipp_gp_predict_cov_workspace(N, M))
        add = noise if (include_noise and not diagonal_only) else 0.0
        L.check(L.lib.ipp_gp_predict_cov(ctypes.byref(self.kern._c), L.ptr(self._X_d), N, L.ptr(self._alpha_d),
                                         L.ptr(self._winv_d), self._ld, add, L.ptr(xq_d), M, L.ptr(mean), L.ptr(cov),
                                         ldc, L.ptr(work), L.stream_ptr()))
        c = cov[:, :M].cpu().numpy()
        if include_noise and diagonal_only:
            c = c + np.eye(M) * noise
        return mean.cpu().numpy().reshape(-1, 1), c

    # ------------------------------------------------------------------ reference API
    def predict_value(self, xvals, include_noise=True):
        self._check_queries(xvals)
        n_points = xvals.shape[0]
        if self.model is None:
            return np.zeros((n_points, 1)), np.ones((n_points, 1)) * self.variance
        return self.model.predict(xvals, full_cov=False, include_likelihood=include_noise)

    def add_data(self, xvals, zvals):
        xvals = np.asarray(xvals, dtype=float)
        zvals = np.asarray(zvals, dtype=float)
        self.xvals = xvals if self.xvals is None else np.vstack([self.xvals, xvals])
        self.zvals = zvals if self.zvals is None else np.vstack([self.zvals, zvals])
        self._factor(want_linv=True)                 # reference :124-125: a new GPRegression every time
        self.model = _LegacyModelHandle(self)

    def posterior_samples(self, xvals, size=10, full_cov=True):
        return self.model.posterior_samples_f(xvals, size, full_cov=full_cov)

    def _refresh_after_kernel_change(self):
        if self.xvals is not None:
            self._factor(want_linv=self._linv_d is not None)

    def load_kernel(self, kernel_file='kernel_model.npy'):
        """reference :134-147: kern[:] = [variance, lengthscale(s)] from a .npy file (the model's own .variance /
        .lengthscale attributes are left alone, as in the reference)."""
        import os
        if not os.path.isfile(kernel_file):
            raise ValueError('Failed to load kernel. Kernel parameter file not found.')
        self.kern[:] = np.load(kernel_file)
        self._refresh_after_kernel_change()

    def log_marginal_likelihood(self):
        """Exact GP log marginal likelihood of the current data under the current kernel, from the device factorisation:
        -0.5 z^T alpha - 0.5 logdet(K + (noise + 1e-8) I) - N/2 log(2 pi)."""
        N = self.n_obs
        quad = float(torch.dot(self._z_d, self._alpha_d).item())
        return -0.5 * quad - 0.5 * float(self._logdet[0].item()) - 0.5 * N * np.log(2.0 * np.pi)

    def train_kernel(self, xvals=None, zvals=None, kernel_file='kernel_model.npy'):
        """reference :149-185: optimise (variance, lengthscale) on the model's data with the noise fixed, save kern[:] to
        `kernel_file`, adopt the values.  Same objective as GPy's (exact log marginal likelihood; every evaluation is one
        device factorisation), L-BFGS-B on the log-parameters with 2 random restarts; GPy's own restart randomisation
        (un-vendored, unpinned) is not reproduced, so trained values agree with GPy's to optimiser tolerance, not bitwise."""
        import scipy.optimize
        if self.xvals is None or self.zvals is None:
            raise ValueError('Failed to train kernel. No training data provided.')
        want_linv = self._linv_d is not None

        def objective(logp):
            try:
                self.kern[:] = np.exp(np.clip(logp, -20.0, 20.0))
                self._factor(want_linv=want_linv)
                val = -self.log_marginal_likelihood()
            except np.linalg.LinAlgError:
                return 1e25
            return val if np.isfinite(val) else 1e25

        start = np.log(self.kern[:])
        best = None
        for x0 in [start] + [np.random.normal(size=start.shape) for _ in range(2)]:
            res = scipy.optimize.minimize(objective, x0, method='L-BFGS-B', options={'maxiter': 60})
            if best is None or res.fun < best.fun:
                best = res
        objective(best.x)                                      # leave the model factored at the optimum
        np.save(kernel_file, self.kern[:])
        self.lengthscale = self.kern.lengthscale
        self.variance = self.kern.variance

    # ------------------------------------------------------------------ deferred status checks
    def defer_checks(self, on=True):
        """Planner-internal: postpone the per-append LinAlgError read-back; call check_pending() later."""
        self._defer = bool(on)
        return self

    def check_pending(self):
        """Raise numpy.linalg.LinAlgError if any append since the last check hit a non-PD Schur block."""
        if self._pending_info:
            bad = int(torch.stack(self._pending_info).max().item())
            self._pending_info = []
            if bad != 0:
                raise np.linalg.LinAlgError('not positive definite, even with jitter.')

    # ------------------------------------------------------------------ copy semantics
    def __copy__(self):
        new = self.__class__.__new__(self.__class__)
        new.__dict__.update(self.__dict__)            # shares immutable device buffers: O(1) fork
        new._pending_info = list(self._pending_info)
        if isinstance(self.model, _LegacyModelHandle):
            new.model = _LegacyModelHandle(new)
        return new

    def __deepcopy__(self, memo):
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if isinstance(v, torch.Tensor):
                new.__dict__[k] = v.clone()
            elif isinstance(v, _LegacyModelHandle):
                new.__dict__[k] = _LegacyModelHandle(new)
            elif isinstance(v, _RBFKern):
                new.__dict__[k] = _RBFKern(v.input_dim, v.variance, v.lengthscale, v.ARD)
            elif k == '_ig':
                new.__dict__[k] = None
            else:
                new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    # ------------------------------------------------------------------ batched planner API
    def _var_operand(self):
        return L.VAR_CHOL, self._linv_d

    def _linv_for_tensor_modes(self):
        """(L^-1 [N][ld] float64 CUDA tensor, ld) of the current belief; a device refactorisation when the model only
        carries W^-1 (OnlineGPModel after init / appends)."""
        if self._linv_d is not None:
            return self._linv_d, self._ld
        N, dev = self.n_obs, L.device()
        ld = L.round_up(N, 16)
        winv = torch.empty((N, ld), dtype=torch.float64, device=dev)
        linv = torch.empty((N, ld), dtype=torch.float64, device=dev)
        alpha = torch.empty(N, dtype=torch.float64, device=dev)
        scal = torch.zeros(2, dtype=torch.float64, device=dev)
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        work = L.workspace(L.lib.ipp_gp_factor_workspace(N))
        L.check(L.lib.ipp_gp_factor(ctypes.byref(self.kern._c), L.ptr(self._X_d), L.ptr(self._z_d), N,
                                    self._diag_add(), L.ptr(winv), L.ptr(linv), ld, L.ptr(alpha), L.ptr(scal),
                                    L.ptr(info), L.ptr(work), L.stream_ptr()))
        if int(info.item()) != 0:
            raise np.linalg.LinAlgError('not positive definite (tensor-mode operand refresh)')
        return linv, ld

    def _i8_operand(self, ns):
        """(tiles, scale): ns signed 7-bit slices of L^-1 against one power-of-two scale, in the kernel's stage-image
        layout (ipp_i8_slice)."""
        if self._i8 is None:
            self._i8 = {}
        if ns not in self._i8:
            N, dev = self.n_obs, L.device()
            linv, lda = self._linv_for_tensor_modes()
            scale = 2.0 ** int(np.ceil(np.log2(float(linv[:, :N].abs().max().item()) * (1.0 + 1e-12))))
            tiles = torch.empty(int(L.lib.ipp_i8_slice_bytes(N, N, ns)), dtype=torch.int8, device=dev)
            L.check(L.lib.ipp_i8_slice(L.ptr(linv), N, N, lda, scale, ns, L.ptr(tiles), L.stream_ptr()))
            self._i8[ns] = (tiles, scale)
        return self._i8[ns]

    def _predict_i8(self, xq_d, include_noise, ns=6):
        M, N = xq_d.shape[0], self.n_obs
        tiles, scale = self._i8_operand(ns)
        mean = torch.empty(M, dtype=torch.float64, device=xq_d.device)
        var = torch.empty(M, dtype=torch.float64, device=xq_d.device)
        work = L.workspace(L.lib.ipp_gp_predict_i8_workspace(N, M), slot='i8')
        L.check(L.lib.ipp_gp_predict_i8(ctypes.byref(self.kern._c), L.ptr(self._X_d), N, L.ptr(self._alpha_d),
                                        L.ptr(tiles), scale, ns, self._noise_var() if include_noise else 0.0,
                                        L.ptr(xq_d), M, L.ptr(mean), L.ptr(var), L.ptr(work), L.stream_ptr()))
        return mean, var

    def predict_device(self, xq_d, include_noise=True, precision=None):
        """(M, D) float64 CUDA tensor -> (mean[M], var[M]) CUDA tensors; no host transfer.
        precision: 'fp64' (FP64 DMMA), 'i8x6' (tcgen05 kind::i8 on six 7-bit slices per operand, exact integer
        accumulation: FP64-grade, ~4e-8 relative variance error), 'i8x7' (seven slices, 28 products: ~3e-10, below the
        FP64 DMMA mode's own round-off against an extended-precision solve) or 'i8x5' (five slices, 15 products on wider
        tiles: ~2e-6, the fast split mode); default = the model's .predict_precision.  Small batches always use the FP64
        path (bandwidth-bound)."""
        if self.n_obs == 0:
            M = xq_d.shape[0]
            return (torch.zeros(M, dtype=torch.float64, device=xq_d.device),
                    torch.full((M,), float(self.variance), dtype=torch.float64, device=xq_d.device))
        precision = self.predict_precision if precision is None else precision
        if precision not in ('fp64', 'i8x7', 'i8x6', 'i8x5'):
            raise ValueError("precision must be 'fp64', 'i8x7', 'i8x6' or 'i8x5'")
        if precision in ('i8x7', 'i8x6', 'i8x5') and xq_d.shape[0] > 64:
            return self._predict_i8(xq_d.contiguous(), include_noise, ns=int(precision[-1]))
        mode, mat = self._var_operand()
        return self._predict_core(xq_d.contiguous(), include_noise, mode, mat)

    def reward_paths_device(self, kind, xq_d, offsets_d, params=None, maxes_d=None, want_points=False):
        """GP predict + acquisition + per-path reduction; everything stays on the device.
        offsets_d: int64 CUDA tensor [P+1].  Returns (reward[P], point_values[M] or None)."""
        mean, var = self.predict_device(xq_d, include_noise=True)
        M, P = xq_d.shape[0], offsets_d.shape[0] - 1
        reward = torch.empty(P, dtype=torch.float64, device=xq_d.device)
        need_points = want_points or kind == L.REWARD_MES
        points = torch.empty(M, dtype=torch.float64, device=xq_d.device) if need_points else None
        par = (ctypes.c_double * 1)(float(params[0]) if params is not None else 0.0)
        S = 0 if maxes_d is None else int(maxes_d.numel())
        L.check(L.lib.ipp_reward_paths(kind, L.ptr(mean), L.ptr(var), L.ptr(offsets_d), P, par, 1, L.ptr(maxes_d), S,
                                       L.ptr(reward), L.ptr(points), M, L.stream_ptr()))
        return reward, points

    def rollout_rewards(self, kind, levels, params=None, maxes_d=None):
        """Every reward of one tree-search rollout in ONE device call (ipp_rollout_rewards): `levels` is the
        rollout's list of (query points (k_l, D), simulated observations (k_l, 1), times the block is appended
        afterwards).  This belief is the base and is not modified.  Returns (rewards ndarray [L], info int);
        info != 0 means a block was not positive definite and the caller must replay through add_data.
        kind REWARD_INFO (info_gain, reference aq_library.py:34-80): the call runs on the information-gain belief
        ((I + vK)^-1 with the variance-squared kernel, unit noise: _info_gain_state) and scales by 2 pi e."""
        if kind == L.REWARD_INFO:
            kern_c, mat, ldm, _ = self._info_gain_state()
            noise_pred, diag_add, params = 0.0, 1.0, [_TWO_PI_E]
        else:
            kern_c, mat, ldm, noise_pred, diag_add = self.kern._c, self._winv_d, self._ld, self._noise_var(), self._diag_add()
        D = self.dimension
        L_ = len(levels)
        offsets = (ctypes.c_int32 * (L_ + 1))()
        reps = (ctypes.c_int32 * L_)()
        M = 0
        for i, (xq, _, r) in enumerate(levels):
            M += xq.shape[0]
            offsets[i + 1] = M
            reps[i] = r
        host = np.empty(M * (D + 1))
        host[:M * D] = np.concatenate([np.asarray(xq, dtype=float).reshape(-1) for xq, _, _ in levels])
        host[M * D:] = np.concatenate([np.asarray(z, dtype=float).reshape(-1) for _, z, _ in levels])
        buf = torch.from_numpy(host).to(L.device(), non_blocking=False)
        out = torch.empty(L_ + 1, dtype=torch.float64, device=buf.device)
        N = self.n_obs
        work = L.workspace(L.lib.ipp_rollout_workspace(N, M), slot='rollout')
        par = (ctypes.c_double * 1)(float(params[0]) if params is not None else 0.0)
        S = 0 if maxes_d is None else int(maxes_d.numel())
        base = buf.data_ptr()
        L.check(L.lib.ipp_rollout_rewards(ctypes.byref(kern_c), L.ptr(self._X_d), N, L.ptr(self._alpha_d),
                                          L.ptr(mat), ldm, noise_pred, diag_add, kind,
                                          par, 1, L.ptr(maxes_d), S, ctypes.c_void_p(base),
                                          ctypes.c_void_p(base + 8 * M * D), offsets, reps, L_,
                                          L.ptr(out), ctypes.c_void_p(out.data_ptr() + 8 * L_), L.ptr(work),
                                          L.stream_ptr()))
        res = out.cpu().numpy()
        return res[:L_], int(res[L_:].view(np.int32)[0])

    # multi-GPU replication by broadcasting the refreshed state (parallel.add_data_replicated)
    def state_tensors(self):
        return [t for t in (self._X_d, self._z_d, self._alpha_d, self._winv_d, self._linv_d) if t is not None]

    def attach_state_shell(self, xvals, zvals):
        """Adopt new observations on the host and allocate device buffers of the matching shape; the
        contents arrive by broadcast from the rank that ran the refresh."""
        xvals = np.asarray(xvals, dtype=float)
        zvals = np.asarray(zvals, dtype=float)
        self.xvals = xvals if self.xvals is None else np.vstack([self.xvals, xvals])
        self.zvals = zvals if self.zvals is None else np.vstack([self.zvals, zvals])
        N, dev = self.n_obs, L.device()
        self._ld = L.round_up(N, 16)
        self._X_d = torch.empty((N, self.dimension), dtype=torch.float64, device=dev)
        self._z_d = torch.empty(N, dtype=torch.float64, device=dev)
        self._alpha_d = torch.empty(N, dtype=torch.float64, device=dev)
        self._winv_d = torch.empty((N, self._ld), dtype=torch.float64, device=dev)
        # rank-independent choice (the source rank allocates L^-1 exactly for these models, see _factor's callers)
        if type(self) is GPModel or getattr(self, 'update_legacy', False):
            self._linv_d = torch.empty((N, self._ld), dtype=torch.float64, device=dev)
            self.model = _LegacyModelHandle(self)
        else:
            self._linv_d = None
        self._ig = None
        self._K = None
        self._i8 = None            # tensor-mode operands were cut for the previous N
        self._pending_info = []

    # info_gain support: (I + v K)^-1 and its log-determinant, cached per data version
    def _info_gain_state(self):
        if self._ig is None:
            N = self.n_obs
            dev = L.device()
            v = float(np.asarray(self.variance, dtype=float).reshape(-1)[0])
            kern2 = L.make_rbf(self.dimension, v * v, self.lengthscale)
            ld = L.round_up(N, 16)
            binv = torch.empty((N, ld), dtype=torch.float64, device=dev)
            scal = torch.zeros(2, dtype=torch.float64, device=dev)
            info = torch.zeros(1, dtype=torch.int32, device=dev)
            work = L.workspace(L.lib.ipp_gp_factor_workspace(N))
            L.check(L.lib.ipp_gp_factor(ctypes.byref(kern2), L.ptr(self._X_d), None, N, 1.0, L.ptr(binv), None, ld,
                                        None, L.ptr(scal), L.ptr(info), L.ptr(work), L.stream_ptr()))
            if int(info.item()) != 0:
                raise np.linalg.LinAlgError('I + v K is not positive definite')
            self._ig = (kern2, binv, ld, float(scal[0].item()))
        return self._ig


class OnlineGPModel(GPModel):
    """Reference gpmodel_library.py:187-466."""

    def __init__(self, ranges, lengthscale, variance, noise=0.0001, dimension=2, kernel='rbf', update_legacy=False):
        super(OnlineGPModel, self).__init__(ranges, lengthscale, variance, noise, dimension, kernel)
        self._K = None
        self.update_legacy = update_legacy
        self._prior_mean = 0.

    def init_model(self, xvals, zvals):                       # reference :208-228
        self.xvals = np.asarray(xvals, dtype=float)
        self.zvals = np.asarray(zvals, dtype=float)
        self._K = None
        self._factor(want_linv=self.update_legacy)

    def update_model(self, xvals, zvals, incremental=True):   # reference :230-279
        assert self.xvals is not None
        assert self.zvals is not None
        xvals = np.asarray(xvals, dtype=float)
        zvals = np.asarray(zvals, dtype=float)
        N, k = self.n_obs, xvals.shape[0]
        self.xvals = np.vstack([self.xvals, xvals])
        self.zvals = np.vstack([self.zvals, zvals])
        self._K = None
        if not incremental or self.update_legacy:
            self._factor(want_linv=self.update_legacy)
            return
        dev = L.device()
        X_all = torch.cat([self._X_d, L.to_device(xvals)], dim=0)
        z_all = torch.cat([self._z_d, L.to_device(zvals.reshape(-1))], dim=0)
        winv, ld = self._winv_d, self._ld
        done = 0
        while done < k:                                       # blocks of <= 32 points (one block for any path)
            kb = min(L.IPP_APPEND_MAX_K, k - done)
            n_old = N + done
            ld_new = L.round_up(n_old + kb, 16)
            w_new = torch.empty((n_old + kb, ld_new), dtype=torch.float64, device=dev)
            alpha = torch.empty(n_old + kb, dtype=torch.float64, device=dev)
            info = torch.zeros(1, dtype=torch.int32, device=dev)
            work = L.workspace(L.lib.ipp_gp_append_workspace(n_old, kb))
            L.check(L.lib.ipp_gp_append(ctypes.byref(self.kern._c), L.ptr(X_all), L.ptr(z_all), n_old, kb,
                                        self._diag_add(), L.ptr(winv), ld, L.ptr(w_new), ld_new, L.ptr(alpha),
                                        L.ptr(info), L.ptr(work), L.stream_ptr()))
            if self._defer:
                self._pending_info.append(info)
            elif int(info.item()) != 0:
                raise np.linalg.LinAlgError('not positive definite, even with jitter.')
            winv, ld = w_new, ld_new
            done += kb
        self._X_d, self._z_d, self._winv_d, self._alpha_d, self._ld = X_all, z_all, winv, alpha, ld
        self._ig = None
        self._i8 = None
        self._linv_d = None

    def add_data(self, xvals, zvals):                         # reference :281-301
        if self.xvals is None:
            assert self.zvals is None
            self.init_model(xvals, zvals)
        else:
            assert self.zvals is not None
            self.update_model(xvals, zvals)
        if self.update_legacy:
            self.model = _LegacyModelHandle(self)

    def predict_value(self, xvals, include_noise=True, full_cov=False):   # reference :303-328
        xvals = np.asarray(xvals, dtype=float)
        self._check_queries(xvals)
        n_points = xvals.shape[0]
        if self.xvals is None:
            return np.zeros((n_points, 1)), np.ones((n_points, 1)) * self.variance
        if full_cov:
            return self._predict_full_cov(xvals, include_noise, self._noise_var(), diagonal_only=False)
        mean, var = self.predict_device(L.to_device(xvals), include_noise)     # FP64 unless .predict_precision says otherwise
        return mean.cpu().numpy().reshape(-1, 1), var.cpu().numpy().reshape(-1, 1)

    def posterior_samples(self, xvals, size=10, full_cov=True):           # reference :331-364
        m, v = self.predict_value(xvals, include_noise=True, full_cov=full_cov)
        if not full_cov:
            return np.random.multivariate_normal(m.flatten(), np.diag(v.flatten()), size).T
        return np.random.multivariate_normal(m.flatten(), v, size).T

    def _var_operand(self):
        return L.VAR_WINV_SYM, self._winv_d

    # host views of the device state (reference properties :366-457)
    @property
    def K(self):
        if self._K is None:
            self._K = self.kern.K(self.xvals, self.xvals)
        return self._K

    @property
    def woodbury_inv(self):
        return None if self._winv_d is None else self._winv_d[:, :self.n_obs].cpu().numpy()

    @property
    def woodbury_vector(self):
        return None if self._alpha_d is None else self._alpha_d.cpu().numpy().reshape(-1, 1)
