"""GPU: whole missions (BASELINE configs 1 and 2 in miniature) -- the device-backed planner and the CPU oracle
planner are stepped side by side on the same seeds: same chosen action at EVERY step while the belief grows by
block appends, and the same posterior at the end.  Reference loops: Robot.choose_trajectory / collect_observations
(robot_library.py:155-208, 325) and Robot.planner -> cMCTS.choose_trajectory (robot_library.py:256-337)."""
import contextlib
import io
import random

import numpy as np
import pytest

from conftest import RANGES

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu


def field(x):
    """A fixed smooth world with one dominant hotspot (stands in for Environment.sample_value)."""
    return (25.0 * np.exp(-0.5 * ((x[:, 0] - 7.0) ** 2 + (x[:, 1] - 6.5) ** 2) / 1.5 ** 2)
            + 8.0 * np.sin(0.8 * x[:, 0]) * np.cos(0.6 * x[:, 1]))[:, None]


@pytest.fixture(scope='module')
def pkg():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import informative_path_planning_b200 as p
    return p


def _models(pkg, noise):
    from oracle.gpmodel_oracle import OnlineGPModelOracle
    d = pkg.OnlineGPModel(RANGES, 1.0, 100.0, noise=noise)
    o = OnlineGPModelOracle(RANGES, 1.0, 100.0, noise=noise)
    return d, o


def _observe(rng, noise, path):
    xobs = np.array(path)[:, :2]
    return xobs, field(xobs) + rng.normal(0, np.sqrt(noise), (xobs.shape[0], 1))


def _truth_ucb(t, paths, X, z, noise):
    """UCB reward of every path from a from-scratch Cholesky solve of (K + (noise + 1e-8) I) -- the arbiter when the
    reference's recursive explicit inverse (and therefore the oracle) has drifted."""
    import scipy.linalg as sl
    from oracle import aq_oracle

    def K(A, B):
        return 100.0 * np.exp(-0.5 * ((A[:, None, :] - B[None, :, :]) ** 2).sum(-1))
    c = sl.cho_factor(K(X, X) + (noise + 1e-8) * np.eye(len(X)), lower=True)
    out = {}
    for k, path in paths.items():
        q = np.array(path)[:, :2]
        kx = K(X, q)
        mu = kx.T @ sl.cho_solve(c, z)
        var = 100.0 - np.einsum('ij,ij->j', kx, sl.cho_solve(c, kx)) + noise
        out[k] = float(mu.sum() + np.sqrt(aq_oracle.ucb_beta(t)) * np.abs(var).sum())
    return out


@pytest.mark.parametrize('noise', [0.5, 1e-4])
def test_myopic_mission_same_actions_every_step(pkg, noise):
    """Config 1 in miniature: UCB reward, Dubins fan, 40 steps (the reference runs 150), N growing by block appends.
    noise 0.5: device == oracle (reference arithmetic) at every step, values to 1e-6.
    noise 1e-4 (config 1's own value, myopic_experiments.py:69): the reference's recursive explicit inverse drifts
    (error 1e-3 by step 4, O(10) by step 15) and then dies with LinAlgError around step 20 in this world; the
    device path (symmetric U U^T append) stays within 1e-3 of a from-scratch solve for the whole mission.  So the
    oracle is compared while it is still within 1e-2 of the truth, the truth at every step."""
    from informative_path_planning_b200 import mcts_library as mc
    from informative_path_planning_b200.paths_library import Dubins_Path_Generator
    from oracle import aq_oracle
    d, o = _models(pkg, noise)
    rng = np.random.default_rng(3)
    x0 = rng.uniform(1, 9, (5, 2))
    z0 = field(x0) + rng.normal(0, np.sqrt(noise), (5, 1))
    d.add_data(x0, z0); o.add_data(x0, z0)
    pg = Dubins_Path_Generator(15, 1.5, 0.05, 0.5, RANGES)
    pose = (5.0, 5.0, 0.0)
    trail, oracle_alive, oracle_steps = [], True, 0
    for t in range(40):
        paths, true_paths = pg.get_path_set(pose)
        ks = sorted(paths)
        np.random.seed(1000 + t)
        best = mc.choose_myopic(d, pg, pose, 'mean', t)
        assert best is not None
        dev = np.array([best[4][k] for k in ks])
        truth = _truth_ucb(t, paths, d.xvals, d.zvals, noise)
        tru = np.array([truth[k] for k in ks])
        np.testing.assert_allclose(dev, tru, rtol=1e-6, atol=1e-6 if noise == 0.5 else 5e-3)
        key = ks[int(np.argmax(tru))]
        assert np.array_equal(np.array(best[0]), np.array(paths[key])), t           # the action the exact posterior picks
        if oracle_alive:
            ora = np.array([float(aq_oracle.mean_UCB(t, paths[k], o)) for k in ks])    # the reference's per-path loop
            if np.abs(ora - tru).max() < 1e-2:
                np.random.seed(1000 + t)
                okey = np.random.choice([k for k in ks if ora[ks.index(k)] == ora.max()])
                assert okey == key, t
                np.testing.assert_allclose(dev, ora, rtol=1e-6, atol=1e-6 if noise == 0.5 else 2e-2)
                oracle_steps += 1
            elif noise == 0.5:
                raise AssertionError('oracle drifted at noise 0.5, step %d' % t)
        xobs, zobs = _observe(rng, noise, paths[key])
        d.add_data(xobs, zobs)
        if oracle_alive:
            try:
                o.add_data(xobs, zobs)
            except np.linalg.LinAlgError:
                assert noise < 0.5
                oracle_alive = False
        pose = true_paths[key][-1]
        trail.append(pose)
    assert d.n_obs > 80 and oracle_steps >= (40 if noise == 0.5 else 3)
    assert len({(round(p[0], 6), round(p[1], 6)) for p in trail}) > 20           # the robot actually moved
    if noise == 0.5:
        q = rng.uniform(0, 10, (300, 2))
        np.testing.assert_allclose(d.predict_value(q)[0], o.predict_value(q)[0], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(d.predict_value(q)[1], o.predict_value(q)[1], rtol=1e-6)


@pytest.mark.parametrize('f_rew,tree_type', [('mean', 'dpw'), ('exp_improve', 'dpw'), ('mean', 'belief')])
def test_nonmyopic_mission_same_actions_every_step(pkg, f_rew, tree_type):
    """Config 2 in miniature: cMCTS per step (budget 40, horizon 3), 8 steps, noise 0.5; rollouts go through the
    one-call-per-rollout device path, the oracle through the reference's per-level update_model."""
    from informative_path_planning_b200 import mcts_library as mc
    from informative_path_planning_b200.paths_library import Dubins_Path_Generator
    from oracle import aq_oracle, mcts_oracle
    aq = pkg.aq_library
    noise = 0.5
    d, o = _models(pkg, noise)
    rng = np.random.default_rng(4)
    x0 = rng.uniform(1, 9, (12, 2))
    z0 = field(x0) + rng.normal(0, np.sqrt(noise), (12, 1))
    d.add_data(x0, z0); o.add_data(x0, z0)
    pg = Dubins_Path_Generator(9, 1.5, 0.05, 0.5, RANGES)
    fn_d = {'mean': aq.mean_UCB, 'exp_improve': aq.exp_improvement}[f_rew]
    fn_o = {'mean': aq_oracle.mean_UCB, 'exp_improve': aq_oracle.exp_improvement}[f_rew]
    pose = (4.0, 4.0, 0.5)
    for t in range(8):
        cur_max = float(np.max(d.zvals))
        res = {}
        for tag, model, fn, cls in (('dev', d, fn_d, mc.cMCTS), ('ora', o, fn_o, mcts_oracle.cMCTSOracle)):
            np.random.seed(50 + t); random.seed(60 + t)
            with contextlib.redirect_stdout(io.StringIO()), np.errstate(all='ignore'):
                planner = cls(40, model, pose, 3, pg, fn, f_rew, t, aq_param=cur_max if f_rew == 'exp_improve' else None,
                              tree_type=tree_type)
                out = planner.choose_trajectory(t=t)
            res[tag] = (out, [c.nqueries for c in planner.tree.root.children])
        assert res['dev'][1] == res['ora'][1], (t, res['dev'][1], res['ora'][1])
        np.testing.assert_array_equal(np.array(res['dev'][0][0]), np.array(res['ora'][0][0]))
        np.testing.assert_allclose(res['dev'][0][2], res['ora'][0][2], rtol=1e-6)
        action, dense = res['dev'][0][0], res['dev'][0][1]
        xobs, zobs = _observe(rng, noise, action)
        d.add_data(xobs, zobs); o.add_data(xobs, zobs)
        pose = dense[-1]
    q = rng.uniform(0, 10, (200, 2))
    np.testing.assert_allclose(d.predict_value(q)[0], o.predict_value(q)[0], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(d.predict_value(q)[1], o.predict_value(q)[1], rtol=1e-6)


def test_environment_world_matches_oracle_built_world(pkg):
    """envmodel_library.Environment (reference :27-260): the same seed gives the same Gaussian world whether the GP
    underneath is the device GPModel or the oracle's GPy restatement (prior draw at one grid point, then one joint
    posterior sample of the remaining NUM_PTS^2 - 1 points through np.random.multivariate_normal), the maximum sits
    inside the 5 % border, and sample_value adds ONE scalar noise draw to the whole batch (Q9)."""
    from informative_path_planning_b200.envmodel_library import Environment
    from oracle.gpmodel_oracle import GPModelOracle
    ranges = (0., 10., 0., 10.)
    dev = Environment(ranges, 12, 100.0, 1.0, noise=0.5, seed=3, dim=2)
    ora = Environment(ranges, 12, 100.0, 1.0, noise=0.5, seed=3, dim=2, gp_cls=GPModelOracle)
    assert dev.GP.xvals.shape == (144, 2) and dev.GP.zvals.shape == (144, 1)
    np.testing.assert_allclose(dev.GP.xvals, ora.GP.xvals)
    # the joint draw goes through an SVD of the 143 x 143 posterior covariance: continuous, not bit-stable
    np.testing.assert_allclose(dev.GP.zvals, ora.GP.zvals, rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(dev.max_loc, ora.max_loc)
    assert 0.5 <= dev.max_loc[0] <= 9.5 and 0.5 <= dev.max_loc[1] <= 9.5
    q = np.random.default_rng(0).uniform(0, 10, (7, 2))
    np.random.seed(9)
    a = dev.sample_value(q)
    np.random.seed(9)
    b = ora.sample_value(q)
    np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-4)
    mean = dev.GP.predict_value(q, include_noise=False)[0]
    assert np.ptp(a - mean) < 1e-9                                   # one scalar noise value for all seven points


def test_environment_world_matches_reference_worlds(pkg):
    """The device-backed Environment against the worlds the REFERENCE's own Environment source drew
    (tests/golden/environment.npz; envmodel_library.py:27-260 executed by tests/golden/make_golden.py: seeds 0-2 at
    NUM_PTS 12, seed 0 at NUM_PTS 20 -- the 399-dimensional multivariate_normal draw and the regenerate-until-inside
    loop).  Tolerance: the joint draw is L ~ SVD(posterior covariance) applied to the same normals, so a 1e-10 change
    of the covariance moves the sample continuously, not bitwise: 1e-5 relative + 1e-4 absolute on values of order 10."""
    from conftest import load_golden
    from informative_path_planning_b200.envmodel_library import Environment
    g = load_golden('environment')
    ranges = (0., 10., 0., 10.)
    for seed, npts in g['cases']:
        tag = 's%d_n%d_' % (seed, npts)
        e = Environment(ranges, int(npts), 100.0, 1.0, noise=0.5, seed=int(seed), dim=2)
        np.testing.assert_array_equal(e.GP.xvals, g[tag + 'X'])
        np.testing.assert_allclose(e.GP.zvals, g[tag + 'z'], rtol=1e-5, atol=1e-4)
        np.testing.assert_array_equal(np.asarray(e.max_loc, dtype=float), g[tag + 'max_loc'])
        np.testing.assert_allclose(np.asarray(e.max_val, dtype=float), g[tag + 'max_val'], rtol=1e-5, atol=1e-4)
        np.random.seed(100 + int(seed))
        np.testing.assert_allclose(e.sample_value(g['q']), g[tag + 'sample'], rtol=1e-5, atol=1e-4)
        assert np.random.random() == g[tag + 'after'][0]          # same number of regenerations, same draws consumed


def test_predict_max_matches_reference_loop(pkg):
    """Robot.predict_max (robot_library.py:222-254): arg-max of the posterior mean over the 100 x 100 grid."""
    from informative_path_planning_b200 import mcts_library as mc
    from oracle.gpmodel_oracle import OnlineGPModelOracle
    d, o = _models(pkg, 0.5)
    assert mc.predict_max(d, RANGES)[1] == 0.
    rng = np.random.default_rng(8)
    x = rng.uniform(0, 10, (60, 2))
    z = field(x) + rng.normal(0, 0.7, (60, 1))
    d.add_data(x, z); o.add_data(x, z)
    loc, val = mc.predict_max(d, RANGES)
    x1, x2 = np.meshgrid(np.linspace(0, 10, 100), np.linspace(0, 10, 100), sparse=False, indexing='xy')
    data = np.vstack([x1.ravel(), x2.ravel()]).T
    obs, _ = o.predict_value(data)
    np.testing.assert_allclose(loc, data[np.argmax(obs), :])
    np.testing.assert_allclose(val, np.max(obs), rtol=1e-9)


# ------------------------------------------------------------------ metrics caller (evaluation_library.Evaluation)
def _evaluation_case(pkg, g):
    import types
    wm = pkg.OnlineGPModel(RANGES, 1.0, 100.0, noise=1e-4, dimension=2)
    wm.add_data(g['Xw'], g['zw'])
    w = types.SimpleNamespace(dim=2, GP=wm, x1min=0., x1max=10., x2min=0., x2max=10.)
    m = pkg.OnlineGPModel(RANGES, 1.0, 100.0, noise=1e-4, dimension=2)
    m.add_data(g['X'], g['z'])
    offs = g['offsets']
    paths = {int(k): [tuple(p) for p in g['pts'][offs[i]:offs[i + 1]]] for i, k in enumerate(g['keys'])}
    return w, m, paths, [tuple(p) for p in g['sel']]


@pytest.mark.parametrize('rf', ['mean', 'hotspot_info', 'info_gain'])
def test_evaluation_update_metrics_golden(pkg, rf):
    """One Evaluation.update_metrics step against the reference source's own output (tests/golden/evaluation.npz):
    both regret sweeps are one batched device call each, the MSE / hotspot probes consume np.random identically."""
    from conftest import load_golden
    from informative_path_planning_b200.evaluation_library import Evaluation
    g = load_golden('evaluation')
    w, m, paths, sel = _evaluation_case(pkg, g)
    e = Evaluation(w, rf)
    np.random.seed(5)
    e.update_metrics(3, m, paths, sel, value=1.25, max_loc=g['max_loc'], max_val=float(g['max_val']),
                     params=[7.0, (1.0, 2.0), None, None], dist=4.5)
    for name, want in zip(g['names'], g['metrics_' + rf]):
        np.testing.assert_allclose(e.metrics[str(name)][3], want, rtol=1e-6, atol=1e-7, err_msg=str(name))
    np.testing.assert_allclose(np.array(e.inst_regret(3, paths, sel, m), dtype=float), g['regret_' + rf], rtol=1e-6)
    assert e.metrics['star_obs_0'][3] == -1. and e.metrics['distance_traveled'][3] == 4.5
    assert e.metrics['robot_location_a'][3] == sel[0][2] and e.metrics['aquisition_function'][3] == 1.25


def test_evaluation_batched_sweep_equals_per_path_calls(pkg, tmp_path):
    """reward_all == [f_rew(p) for p in paths] for every GP-backed evaluation reward, with and without robot
    observations; the CSV writer lays rows out as the reference's plot_metrics.  ('naive' / 'naive_value' only ever
    get the -1 sentinels in update_metrics; their f_rew indexes a 1-D max_loc as 2-D in the reference and raises.)"""
    from conftest import load_golden
    from informative_path_planning_b200.evaluation_library import Evaluation
    g = load_golden('evaluation')
    w, m, paths, sel = _evaluation_case(pkg, g)
    empty = pkg.OnlineGPModel(RANGES, 1.0, 100.0, noise=1e-4, dimension=2)
    plist = [paths[k] for k in sorted(paths)]
    for rf in ('mean', 'mes', 'exp_improve', 'hotspot_info', 'info_gain'):
        e = Evaluation(w, rf)
        for model in (m, empty):
            if rf == 'hotspot_info' and model is empty:
                continue
            batched = e.reward_all(3, plist, model)
            single = np.array([e.f_rew(time=3, xvals=np.array(p), robot_model=model) for p in plist], dtype=float)
            np.testing.assert_allclose(batched, single, rtol=1e-9, atol=1e-9, err_msg=rf)
    with pytest.raises(ValueError):
        Evaluation(w, 'no_such_reward')
    e = Evaluation(w, 'mean')
    for t in range(3):
        np.random.seed(t)
        e.update_metrics(t, m, paths, sel, value=0.5 * t, max_loc=g['max_loc'], max_val=float(g['max_val']),
                         params=[7.0, (1.0, 2.0), None if t == 0 else np.array([[1.0], [2.0], [3.0]]),     # MES max-value samples: (nK, 1)
                                 [(0, 1), (2, 3), (4, 5)]], dist=float(t))
    out = e.plot_metrics(str(tmp_path))
    rows = np.loadtxt(out + '/metrics.csv')
    assert rows.shape == (21, 3)
    np.testing.assert_allclose(rows[0], [0, 1, 2])
    np.testing.assert_allclose(rows[2], np.cumsum([0.0, 0.5, 1.0]))
    np.testing.assert_allclose(rows[10], np.cumsum([e.metrics['instant_regret'][t] for t in range(3)]))
    assert np.loadtxt(out + '/stars.csv').shape == (9, 3)


# ------------------------------------------------------------------ whole missions of the reference's Robot
@pytest.mark.parametrize('name', ['myopic_mean_fan', 'myopic_mean_dubins_blocks', 'myopic_mes_fan', 'nonmyopic_mean_dpw',
                                  'nonmyopic_mean_belief_dubins', 'nonmyopic_mean_dpw_dubins',
                                  'nonmyopic_mes_dpw_dubins_blocks', 'nonmyopic_ei_dpw_dubins_channel',
                                  'myopic_info_gain_fan', 'myopic_ei_dubins', 'myopic_naive_dubins',
                                  'nonmyopic_naive_dpw_dubins', 'nonmyopic_naive_value_dpw_dubins',
                                  'nonmyopic_info_gain_dpw_dubins', 'nonmyopic_hotspot_dpw_dubins'])
def test_robot_mission_reproduces_reference_run(pkg, name, tmp_path, monkeypatch):
    """robot_library.Robot.planner + evaluation_library.Evaluation on the device against the SAME seeded mission run by
    the reference's own Robot / Evaluation source (tests/golden/missions.npz): the pose after every step, every
    observation (so the random streams stay in lock-step), and the per-step metrics.  The dpw + Dubins missions run
    through the native tree search (csrc/tree.cu), which therefore reproduces the reference's own tree code as well."""
    import os
    import sys
    from conftest import GOLDEN, load_golden
    if GOLDEN not in sys.path:
        sys.path.insert(0, GOLDEN)
    import mission_spec as ms
    from informative_path_planning_b200 import obstacles as ob
    from informative_path_planning_b200.evaluation_library import Evaluation
    from informative_path_planning_b200.robot_library import Robot
    from informative_path_planning_b200 import mcts_library as mc
    g = load_golden('missions')
    spec = ms.MISSIONS[name]
    native_before = mc.cMCTS.native_searches
    monkeypatch.chdir(tmp_path)                                   # planner writes ./figures/<f_rew>/robot_model.csv
    e = Evaluation(ms.mission_world(pkg.OnlineGPModel), spec[0])
    np.random.seed(11)
    random.seed(12)
    r = Robot(**ms.mission_kwargs(name, e, ms.mission_obstacles(ob, spec[4]), ms.mission_sample_world))
    locs = []
    orig = r.collect_observations

    def spy(xobs):
        orig(xobs)
        locs.append(np.array(xobs[-1], dtype=float))
    r.collect_observations = spy
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all='ignore'):
        r.planner(T=spec[5])
    np.testing.assert_allclose(np.array(locs), g[name + '_locs'], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(np.array([float(v) for v in r.loc]), g[name + '_end'], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(r.GP.xvals, g[name + '_X'], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(r.GP.zvals, g[name + '_z'], rtol=1e-9, atol=1e-9)
    got = np.array([[float(e.metrics[n][t]) for n in ms.MISSION_METRICS] for t in range(spec[5])])
    np.testing.assert_allclose(got, g[name + '_metrics'], rtol=2e-5, atol=1e-5)
    assert os.path.exists(os.path.join('figures', spec[0], 'robot_model.csv'))
    # every tree search of these missions runs natively (dpw and belief trees, both generators, all rewards: hotspot_info
    # as two device passes per rollout, UCB on the regression belief + information gain on (I + vK)^-1)
    native_expected = spec[5] if spec[1] else 0
    assert mc.cMCTS.native_searches - native_before == native_expected
    x1, x2, mean_grid, var_grid = r.visualize_world_model(screen=False)
    assert mean_grid.shape == (100, 100) and np.isfinite(mean_grid).all() and (var_grid > 0).all()
    if spec[0] == 'mean':
        _, _, reward_grid = r.visualize_reward(t=3)
        assert reward_grid.shape == (100, 100) and np.isfinite(reward_grid).all()
