"""CPU restatement of the reward / acquisition callables (L2 of SURVEY.md).  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/aq_library.py:
  info_gain :34-80, mean_UCB :82-114, hotspot_info_UCB :117-135, general_target :138-139,
  sample_max_vals :141-279, mves :281-337, naive :339-367, naive_value :369-413,
  global_maximization :444-553, exp_improvement :556-582.
Every quirk catalogued in SURVEY.md section 8a (Q1, Q2, Q6, Q7) is kept on purpose.
Random draws come from the global ``np.random`` in the reference's order, so a seeded run of
this file and of the reference consume the same stream.
"""
import numpy as np
import scipy.optimize
import scipy.special
import scipy.stats


def _queries(time, xvals, robot_model):
    data = np.array(xvals)
    x1 = data[:, 0]
    x2 = data[:, 1]
    if robot_model.dimension == 2:
        return data, np.vstack([x1, x2]).T
    return data, np.vstack([x1, x2, time * np.ones(len(x1))]).T


def ucb_beta(time):
    """aq_library.py:106-109."""
    delta = 0.9
    d = 20
    pit = np.pi ** 2 * (time + 1) ** 2 / 6.
    return 2 * np.log(d * pit / delta)


def info_gain(time, xvals, robot_model, param=None):
    _, queries = _queries(time, xvals, robot_model)
    xobs = robot_model.xvals
    if xobs is None:                                                     # :48-53
        Sigma_after = robot_model.kern.K(queries)
        sign_after, entropy_after = np.linalg.slogdet(np.eye(Sigma_after.shape[0]) + robot_model.variance * Sigma_after)
        return 0.5 * sign_after * entropy_after
    all_data = np.vstack([xobs, queries])
    Sigma_before = robot_model.kern.K(xobs)
    Sigma_total = robot_model.kern.K(all_data)
    # NB the reference unpacks slogdet as (entropy, sign) = (sign, logabsdet): the names are
    # swapped (:50,:62,:66) but only the product sign*logabsdet is used, so the value is the same.
    s_b, l_b = np.linalg.slogdet(np.eye(Sigma_before.shape[0]) + robot_model.variance * Sigma_before)
    s_a, l_a = np.linalg.slogdet(np.eye(Sigma_total.shape[0]) + robot_model.variance * Sigma_total)
    return 2 * np.pi * np.e * s_a * l_a - 2 * np.pi * np.e * s_b * l_b   # :70, entropy_const = 0


def mean_UCB(time, xvals, robot_model, param=None, FVECTOR=False):
    data, queries = _queries(time, xvals, robot_model)
    if robot_model.xvals is None:                                        # :93-97
        return np.ones((data.shape[0], 1)) if FVECTOR else 1.0
    mu, var = robot_model.predict_value(queries)
    beta_t = ucb_beta(time)
    if FVECTOR:
        return mu + np.sqrt(beta_t) * np.fabs(var)                       # :112  (variance, not std: Q1)
    return np.sum(mu) + np.sqrt(beta_t) * np.sum(np.fabs(var))           # :114


def hotspot_info_UCB(time, xvals, robot_model, param=None):
    _, queries = _queries(time, xvals, robot_model)
    mu, var = robot_model.predict_value(queries)
    beta_t = ucb_beta(time)
    return info_gain(time, xvals, robot_model) + 1.0 * np.sum(mu) + np.sqrt(beta_t) * np.sum(np.fabs(var))   # :135


def general_target(x, variance, nFeatures, W, theta, b):
    """aq_library.py:138-139 with robot_model.variance passed as a number."""
    return np.dot(theta.T * np.sqrt(2.0 * variance / nFeatures), np.cos(np.dot(W, x.T) + b)).T


def rff_theta(Z, zvals, noise_var, eps, nFeatures):
    """Posterior weight draw, aq_library.py:177-206.  Z is (F, N); eps ~ N(0, I_F) is (F,1)."""
    N = Z.shape[1]
    if N < nFeatures:                                                    # :177-187
        Sigma = np.dot(Z.T, Z) + noise_var * np.eye(N)
        mu = np.dot(np.dot(Z, np.linalg.inv(Sigma)), zvals)
        D, U = np.linalg.eig(Sigma)
        U = np.real(U)
        D = np.real(np.reshape(D, (D.shape[0], 1)))
        R = np.reciprocal(np.sqrt(D) * (np.sqrt(D) + np.sqrt(noise_var)))
        return eps - np.dot(Z, np.dot(U, R * (np.dot(U.T, np.dot(Z.T, eps))))) + mu
    Sigma = np.dot(Z, Z.T) / noise_var + np.eye(nFeatures)               # :197-200
    Sigma = np.linalg.inv(Sigma)
    mu = np.dot(np.dot(Sigma, Z), zvals) / noise_var
    return mu + np.dot(np.linalg.cholesky(Sigma), eps)


def rff_features(W, b, X, variance, nFeatures):
    """aq_library.py:171."""
    return np.sqrt(2 * variance / nFeatures) * np.cos(np.dot(W, X.T) + b)


def grid_stage(target, ranges, guesses, time=None):
    """Grid part of global_maximization, aq_library.py:449-485.  Returns (start, shrunk ranges)."""
    gridSize = 300
    bb = ((ranges[1] - ranges[0]) * 0.10, (ranges[3] - ranges[2]) * 0.10)
    ranges = (ranges[0] + bb[0], ranges[1] - bb[0], ranges[2] + bb[1], ranges[3] - bb[1])
    dim = guesses.shape[1]
    x1 = np.random.uniform(ranges[0], ranges[1], size=gridSize)
    x2 = np.random.uniform(ranges[2], ranges[3], size=gridSize)
    x1, x2 = np.meshgrid(x1, x2, sparse=False, indexing='xy')
    if dim == 2:
        Xgrid_sample = np.vstack([x1.ravel(), x2.ravel()]).T
        Xgrid = np.vstack([Xgrid_sample, guesses])
    else:
        Xgrid_sample = np.vstack([x1.ravel(), x2.ravel(), time * np.ones(len(x1.ravel()))]).T
        Xgrid = Xgrid_sample
    y = target(Xgrid)
    start = np.asarray(Xgrid[np.argmax(y), :])
    if start[0] < ranges[0] or start[0] > ranges[1] or start[1] < ranges[2] or start[1] > ranges[3]:   # :482-485
        y = target(Xgrid_sample)
        start = np.asarray(Xgrid_sample[np.argmax(y), :])
    return start, ranges


def global_maximization(target, target_vector_n, target_vector_gradient_n, ranges, guesses, time=None):
    """aq_library.py:444-553 without the plotting."""
    start, ranges = grid_stage(target, ranges, guesses, time)
    dim = guesses.shape[1]
    if dim == 2:
        bounds = ((ranges[0], ranges[1]), (ranges[2], ranges[3]))
    else:
        bounds = ((ranges[0], ranges[1]), (ranges[2], ranges[3]), (time, time))
    res = scipy.optimize.minimize(fun=target_vector_n, x0=start, method='SLSQP', jac=target_vector_gradient_n,
                                  bounds=bounds)
    if not res['success']:
        return 0, 0, 0, False
    return res['x'], -res['fun'], res['jac'], True


def sample_max_vals(robot_model, t, nK=3, nFeatures=200, visualize=False, obstacles=None, f_rew='mes'):
    """aq_library.py:141-279."""
    if robot_model.xvals is None:
        return None, None, None
    d = robot_model.xvals.shape[1]
    samples = np.zeros((nK, 1))
    locs = np.zeros((nK, d))
    funcs = []
    delete_locs = []
    variance = robot_model.variance
    for i in range(nK):
        ls = robot_model.lengthscale if robot_model.dimension == 2 else robot_model.lengthscale[0]
        W = np.random.normal(loc=0.0, scale=np.sqrt(1. / ls), size=(nFeatures, d))        # :163 (Q7)
        b = 2 * np.pi * np.random.uniform(low=0.0, high=1.0, size=(nFeatures, 1))         # :167
        Z = rff_features(W, b, robot_model.xvals, variance, nFeatures)                    # :171
        eps = np.random.normal(loc=0.0, scale=1.0, size=(nFeatures, 1))                   # :174
        try:
            theta = rff_theta(Z, robot_model.zvals, robot_model.noise, eps, nFeatures)
        except Exception:
            delete_locs.append(i)
            continue

        def target(x, W=W, theta=theta, b=b):
            return general_target(x, variance, nFeatures, W, theta, b)

        def target_vector_n(x, target=target):
            return -target(x.reshape(1, d))

        def target_gradient(x, W=W, theta=theta, b=b):                                    # :226
            return np.dot(theta.T * -np.sqrt(2.0 * variance / nFeatures), np.sin(np.dot(W, x.reshape((d, 1))) + b) * W)

        def target_vector_gradient_n(x, tg=target_gradient):
            return -np.asarray(tg(x).reshape(d, ))

        status = False
        count = 0
        while status is False and count < 5:                                              # :233-245
            maxima, max_val, _, status = global_maximization(target, target_vector_n, target_vector_gradient_n,
                                                             robot_model.ranges, robot_model.xvals, time=t)
            count += 1
        if status is False:
            delete_locs.append(i)
            continue
        samples[i] = np.array(max_val).reshape((1, 1))
        funcs.append(target)
        locs[i, :] = maxima.reshape((1, d))
    samples = np.delete(samples, delete_locs, axis=0)
    locs = np.delete(locs, delete_locs, axis=0)
    if len(delete_locs) == nK:
        # :273-275 -- after np.delete the arrays are empty, so the reference's samples[0] = ...
        # raises IndexError; kept as is.
        samples[0] = np.max(robot_model.zvals) + 5.0 * np.sqrt(robot_model.noise)
        locs[0, :] = robot_model.xvals[np.argmax(robot_model.zvals)]
    return samples, locs, funcs


def mves_utility(maxes_i, mean, var):
    """aq_library.py:316-319 for one max value (gamma divides by the VARIANCE: Q2)."""
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        gamma = (maxes_i - mean) / var
        pdfgamma = scipy.stats.norm.pdf(gamma)
        cdfgamma = scipy.stats.norm.cdf(gamma)
        return gamma * pdfgamma / (2.0 * cdfgamma) - np.log(cdfgamma)


def mves(time, xvals, robot_model, param, FVECTOR=False):
    maxes = param[0]
    if maxes is None:                                                    # :287-291
        return np.ones((xvals.shape[0], 1)) if FVECTOR else 1.0
    data, queries = _queries(time, xvals, robot_model)
    f = np.zeros((data.shape[0], 1)) if FVECTOR else 0
    for i in range(maxes.shape[0]):
        mean, var = robot_model.predict_value(queries)                   # :313 (recomputed per max)
        utility = mves_utility(maxes[i], mean, var)
        if FVECTOR:
            f = f + utility
        else:
            f = f + sum(utility)                                         # python sum over rows -> shape (1,)
        f = f + sum(utility)                                             # :329 the second, unconditional add
    f = f / maxes.shape[0]
    return f if FVECTOR else f[0]


def naive(time, xvals, robot_model, param, FVECTOR=False):
    _, max_locs, _ = param[0]
    if max_locs is None:
        return np.zeros((xvals.shape[0], 1)) if FVECTOR else 0.0
    data = np.array(xvals)
    x1 = data[:, 0]
    x2 = data[:, 1]
    f = np.zeros((data.shape[0], 1))
    for i in range(max_locs.shape[0]):
        dist = np.sqrt(np.square(x1 - max_locs[i][0]) + np.square(x2 - max_locs[i][1]))
        f += (dist <= param[1]).astype(float).reshape(f.shape)
    f = f / max_locs.shape[0]
    return f if FVECTOR else np.sum(f)


def naive_value(time, xvals, robot_model, param, FVECTOR=False):
    max_vals, _, funcs = param[0]
    if max_vals is None or funcs is None:
        return np.zeros((xvals.shape[0], 1)) if FVECTOR else 0.0
    data, queries = _queries(time, xvals, robot_model)
    f = np.zeros((data.shape[0], 1))
    for i in range(max_vals.shape[0]):
        mean = funcs[i](queries)
        diff = np.fabs(mean - max_vals[i][0])
        f += (diff <= param[1]).astype(float).reshape(f.shape)
    f = f / float(max_vals.shape[0])
    return f if FVECTOR else np.sum(f)


def exp_improvement(time, xvals, robot_model, param=None):
    _, queries = _queries(time, xvals, robot_model)
    mu, var = robot_model.predict_value(queries)
    eta = 0.5 if param is None else sum(param) / len(param)              # :569-572
    x = np.sum([m - eta for m in mu])                                    # :575-576
    z = x / np.sum(np.fabs(var))                                         # :577 (variance, not std: Q1)
    big_phi = 0.5 * (1 + scipy.special.erf(z / np.sqrt(2)))
    small_phi = 1 / np.sqrt(2 * np.pi) * np.exp(-z ** 2 / 2)
    return x * big_phi + np.sum(np.fabs(var)) * small_phi                # :580


REWARDS = {
    'mean': mean_UCB, 'mes': mves, 'exp_improve': exp_improvement, 'info_gain': info_gain,
    'hotspot_info': hotspot_info_UCB, 'naive': naive, 'naive_value': naive_value,
}
