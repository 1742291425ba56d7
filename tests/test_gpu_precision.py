"""The ill-conditioned regime in the suite: every predict mode against an EXTENDED-PRECISION solve of the same
posterior variance -- Cholesky in float64 + three rounds of iterative refinement with long-double residuals, good to
~1e-15 -- at N in {1024, 4096} and noise in {0.5 (BASELINE config 2), 1e-4 (config 1)}.

Bounds (measured on a B200 by tools/check_precision_modes.py, profiles/r2_precision_modes.json, asserted with a margin):
relative error where the posterior variance is of the order of the noise or larger (noise 0.5); at noise 1e-4 the
variance near the data is ~1e-4 against a prior of 100, so the error is bounded against the prior scale and,
separately, relative to the true variance for the queries that are not on top of the data."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')

RANGES = (0., 10., 0., 10.)


def make_case(N, noise, nq=16):
    rng = np.random.default_rng(N + int(1e6 * noise))
    X = rng.uniform(0, 10, (N, 2))
    z = 10 * np.sin(0.7 * X[:, :1]) * np.cos(0.5 * X[:, 1:]) + rng.normal(0, np.sqrt(noise), (N, 1))
    return X, z, rng.uniform(0, 10, (nq, 2))


def extended_precision_variance(X, q, noise):
    import scipy.linalg as sl

    def Kf(A, B):
        return 100.0 * np.exp(-0.5 * ((A[:, None, :] - B[None, :, :]) ** 2).sum(-1))
    N = X.shape[0]
    Ky = Kf(X, X) + (noise + 1e-8) * np.eye(N)
    c = sl.cho_factor(Ky, lower=True)
    Ky_l = Ky.astype(np.longdouble)
    kq = Kf(X, q)
    x = sl.cho_solve(c, kq).astype(np.longdouble)
    for _ in range(3):
        r = kq.astype(np.longdouble) - Ky_l @ x
        x = x + sl.cho_solve(c, r.astype(np.float64)).astype(np.longdouble)
    return (np.longdouble(100.0) - np.einsum('ij,ij->j', kq.astype(np.longdouble), x) + np.longdouble(noise)).astype(np.float64)


# mode -> (relative bound at noise 0.5, bound on |error| / prior variance at noise 1e-4); measured (N = 1024 / 4096):
#   noise 0.5 , relative:          fp64 3.0e-10 / 1.1e-9   i8x7 7.9e-11 / 1.5e-10   i8x6 1.4e-8 / 1.8e-8   i8x5 2.6e-6 / 2.2e-6
#   noise 1e-4, |err| / prior:     fp64 4.6e-9  / 4.4e-8   i8x7 4.3e-11 / 2.3e-11   i8x6 4.5e-9 / 3.5e-9   i8x5 3.8e-7 / 3.9e-7
# (at noise 1e-4 the true variance near the data is ~1e-4: the FP64 explicit-inverse form is off by up to 4 % of it at
# N = 4096, the seven-slice Cholesky form by 2e-5 -- the reference's own arithmetic is the former)
BOUNDS = {'fp64': (1e-8, 2.5e-7), 'i8x7': (1e-9, 2.5e-10), 'i8x6': (1e-7, 2.5e-8), 'i8x5': (1.5e-5, 2e-6)}


@pytest.mark.gpu
@pytest.mark.parametrize('N', [1024, 4096])
@pytest.mark.parametrize('noise', [0.5, 1e-4])
def test_predict_modes_against_extended_precision(N, noise):
    import informative_path_planning_b200 as pkg
    X, z, q = make_case(N, noise)
    truth = extended_precision_variance(X, q, noise)
    d = pkg.OnlineGPModel(RANGES, 1.0, 100.0, noise=noise)
    d.add_data(X, z)
    big = np.vstack([q, np.random.default_rng(1).uniform(0, 10, (4096, 2))])       # > 64 queries: the tensor modes engage
    errs = {}
    for prec, (rel_05, abs_1e4) in BOUNDS.items():
        d.predict_precision = prec
        var = d.predict_value(big)[1][:q.shape[0], 0]
        err = np.abs(var - truth)
        errs[prec] = float(np.max(err / truth))
        if noise == 0.5:
            assert errs[prec] <= rel_05, (prec, N, errs[prec])
        else:
            assert float(np.max(err)) / 100.0 <= abs_1e4, (prec, N, float(np.max(err)) / 100.0)
    # the ranking DESIGN.md states: seven slices beat the FP64 arithmetic path, six are FP64-grade, five are the fast mode
    assert errs['i8x7'] <= errs['fp64'] and errs['i8x7'] <= errs['i8x6'] <= errs['i8x5']
    if noise == 0.5:
        assert errs['i8x6'] <= 1e-6 and errs['i8x5'] <= 1e-4
    else:
        assert errs['i8x7'] <= errs['fp64'] / 10          # Cholesky form + exact accumulation vs the cancelling inverse form
