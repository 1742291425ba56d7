#!/usr/bin/env python
"""Generate golden input/output vectors by RUNNING THE REFERENCE'S OWN SOURCE (see ref_loader.py).

Run in the build container only:   python tests/golden/make_golden.py
Writes tests/golden/*.npz (small, committed).  The GPU box never runs this file.

Every case records its inputs (so the oracle and the CUDA path are fed the identical arrays)
and the outputs the reference code produced.
"""
import contextlib
import io
import os
import random
import sys

import numpy as np

SRC = os.path.dirname(os.path.abspath(__file__))
HERE = os.environ.get('IPP_GOLDEN_OUT') or SRC          # IPP_GOLDEN_OUT: write elsewhere (reproducibility checks)
sys.path.insert(0, SRC)
sys.path.insert(0, os.path.dirname(os.path.dirname(SRC)))
import ref_loader  # noqa: E402


class _DubinsPath(object):
    """The two calls the reference makes on the absent `dubins` package (paths_library.py:151-152), served by
    oracle/dubins_oracle.py -- the restated library, pinned by tests/test_dubins.py; no product code is involved."""

    def __init__(self, q0, q1, rho):
        from oracle import dubins_oracle
        self._do, self.q0, self.rho = dubins_oracle, q0, rho
        self.word, self.params = dubins_oracle.shortest_path(q0, q1, rho)

    def sample_many(self, step):
        return self._do.sample_many(self.q0, self.word, self.params, self.rho, step)

RANGES = (0., 10., 0., 10.)


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def world(rng, n, noise):
    """Synthetic 'Gaussian environment' observations: uniform X, z = smooth field + noise."""
    X = rng.uniform(0, 10, (n, 2))
    z = 10.0 * np.sin(0.7 * X[:, :1]) * np.cos(0.5 * X[:, 1:2]) + 3.0 * np.cos(1.3 * X[:, :1] + 0.4 * X[:, 1:2])
    z = z + rng.normal(0, np.sqrt(noise), (n, 1))
    return X, z


def case_online_gp(gp, name, n0, appends, k, noise, nq, seed):
    rng = np.random.default_rng(seed)
    X, z = world(rng, n0 + appends * k, noise)
    m = gp.OnlineGPModel(RANGES, 1.0, 100.0, noise=noise, dimension=2)
    m.add_data(X[:n0], z[:n0])
    for a in range(appends):
        s = n0 + a * k
        m.add_data(X[s:s + k], z[s:s + k])
    q = rng.uniform(0, 10, (nq, 2))
    q[0] = X[0]                                  # a query exactly on a datum
    q[1] = [55.0, -40.0]                         # far away: prior
    mu, var = m.predict_value(q)
    mu_nn, var_nn = m.predict_value(q, include_noise=False)
    mu_fc, cov_fc = m.predict_value(q[:8], include_noise=True, full_cov=True)
    extra = {'winv': m.woodbury_inv} if X.shape[0] <= 100 else {}       # keep fixtures small
    np.savez_compressed(os.path.join(HERE, name + '.npz'), X=X, z=z, n0=n0, appends=appends, k=k, noise=noise, q=q,
                        mu=mu, var=var, mu_nn=mu_nn, var_nn=var_nn, mu_fc=mu_fc, cov_fc=cov_fc,
                        alpha=m.woodbury_vector, **extra)
    return m


def case_duplicates(gp):
    """Q4: the trees append the same (xobs, zobs) twice (mcts_library.py:426 & :430)."""
    rng = np.random.default_rng(7)
    X, z = world(rng, 30, 0.5)
    m = gp.OnlineGPModel(RANGES, 1.0, 100.0, noise=0.5, dimension=2)
    m.add_data(X[:22], z[:22])
    for s in (22, 26):
        m.add_data(X[s:s + 4], z[s:s + 4])
        m.add_data(X[s:s + 4], z[s:s + 4])
    q = rng.uniform(0, 10, (16, 2))
    mu, var = m.predict_value(q)
    np.savez_compressed(os.path.join(HERE, 'online_gp_duplicates.npz'), X=X, z=z, q=q, mu=mu, var=var,
                        winv=m.woodbury_inv, alpha=m.woodbury_vector)


def case_gpy_model(gp):
    rng = np.random.default_rng(11)
    X, z = world(rng, 48, 0.5)
    m = gp.GPModel(RANGES, 1.0, 100.0, noise=0.5, dimension=2)
    m.add_data(X[:40], z[:40])
    m.add_data(X[40:], z[40:])
    q = rng.uniform(0, 10, (32, 2))
    mu, var = m.predict_value(q)
    mu_nn, var_nn = m.predict_value(q, include_noise=False)
    e = gp.GPModel(RANGES, 1.0, 100.0)
    mu0, var0 = e.predict_value(q)
    np.savez_compressed(os.path.join(HERE, 'gpmodel_gpy.npz'), X=X, z=z, q=q, mu=mu, var=var, mu_nn=mu_nn,
                        var_nn=var_nn, mu0=mu0, var0=var0)


def case_rewards(gp, aq):
    rng = np.random.default_rng(3)
    out = {}
    for tag, noise, n in (('n05', 0.5, 60), ('n1e4', 1e-4, 45)):
        X, z = world(rng, n, noise)
        m = gp.OnlineGPModel(RANGES, 1.0, 100.0, noise=noise, dimension=2)
        m.add_data(X, z)
        P, k = 12, 5
        paths = rng.uniform(0.5, 9.5, (P, k, 3))        # (x, y, heading) tuples like the path generators emit
        maxes = np.array([[22.0], [27.5], [31.0]])
        out[tag + '_X'], out[tag + '_z'], out[tag + '_paths'], out[tag + '_maxes'] = X, z, paths, maxes
        out[tag + '_noise'] = noise
        t = 17
        out[tag + '_t'] = t
        out[tag + '_ucb'] = np.array([aq.mean_UCB(time=t, xvals=p, robot_model=m) for p in paths])
        out[tag + '_ucb_f'] = aq.mean_UCB(time=t, xvals=paths.reshape(-1, 3), robot_model=m, FVECTOR=True)
        out[tag + '_ei'] = np.array([aq.exp_improvement(time=t, xvals=p, robot_model=m, param=[float(np.max(z))])
                                     for p in paths])
        out[tag + '_ei_default'] = np.array([aq.exp_improvement(time=t, xvals=p, robot_model=m) for p in paths])
        with np.errstate(all='ignore'):
            out[tag + '_mes'] = np.array([aq.mves(time=t, xvals=p, robot_model=m, param=(maxes, None, None))
                                          for p in paths])
            out[tag + '_mes_f'] = aq.mves(time=t, xvals=paths.reshape(-1, 3), robot_model=m,
                                          param=(maxes, None, None), FVECTOR=True)
        out[tag + '_ig'] = np.array([aq.info_gain(time=t, xvals=p, robot_model=m) for p in paths])
        out[tag + '_hot'] = np.array([aq.hotspot_info_UCB(time=t, xvals=p, robot_model=m) for p in paths])
        max_locs = rng.uniform(2, 8, (3, 2))
        out[tag + '_max_locs'] = max_locs
        out[tag + '_naive'] = np.array([aq.naive(time=t, xvals=p, robot_model=m, param=((maxes, max_locs, None), 1.5))
                                        for p in paths])
    # no-data branches
    e = gp.OnlineGPModel(RANGES, 1.0, 100.0, noise=0.5, dimension=2)
    p = out['n05_paths'][0]
    out['empty_ucb'] = np.array(aq.mean_UCB(time=0, xvals=p, robot_model=e))
    out['empty_ucb_f'] = aq.mean_UCB(time=0, xvals=p, robot_model=e, FVECTOR=True)
    out['empty_ig'] = np.array(aq.info_gain(time=0, xvals=p, robot_model=e))
    out['empty_mes'] = np.array(aq.mves(time=0, xvals=p, robot_model=e, param=(None, None, None)))
    out['empty_ei'] = np.array(aq.exp_improvement(time=0, xvals=p, robot_model=e, param=[-1000]))
    np.savez_compressed(os.path.join(HERE, 'rewards.npz'), **out)


def case_max_vals(gp, aq):
    out = {}
    for tag, n, F, nK, seed in (('lt', 50, 200, 3, 5), ('ge', 120, 60, 2, 9)):
        rng = np.random.default_rng(seed)
        X, z = world(rng, n, 0.5)
        m = gp.OnlineGPModel(RANGES, 1.0, 100.0, noise=0.5, dimension=2)
        m.add_data(X, z)
        np.random.seed(1000 + seed)
        with quiet():
            samples, locs, funcs = aq.sample_max_vals(m, t=0, nK=nK, nFeatures=F)
        probe = rng.uniform(1, 9, (64, 2))
        fvals = np.hstack([f(probe) for f in funcs])
        out.update({tag + '_X': X, tag + '_z': z, tag + '_F': F, tag + '_nK': nK, tag + '_seed': 1000 + seed,
                    tag + '_samples': samples, tag + '_locs': locs, tag + '_probe': probe, tag + '_fvals': fvals})
    np.savez_compressed(os.path.join(HERE, 'max_vals.npz'), **out)


def case_mcts(gp, aq, mc, pl):
    """Action selection of the reference's own Tree / BeliefTree on fixed seeds (straight-line
    Path_Generator: paths_library.py:18-117, the one generator that needs no `dubins`)."""
    out = {}
    rng = np.random.default_rng(21)
    X, z = world(rng, 40, 0.5)
    out['X'], out['z'] = X, z
    for tag, f_rew, tree_type, budget in (('dpw_mean', 'mean', 'dpw', 60), ('dpw_mes', 'mes', 'dpw', 50),
                                          ('belief_mean', 'mean', 'belief', 40),
                                          ('dpw_ei', 'exp_improve', 'dpw', 40)):
        m = gp.OnlineGPModel(RANGES, 1.0, 100.0, noise=0.5, dimension=2)
        m.add_data(X, z)
        pg = pl.Path_Generator(6, 1.5, 0.05, 0.5, RANGES)
        fn = {'mean': aq.mean_UCB, 'mes': aq.mves, 'exp_improve': aq.exp_improvement}[f_rew]
        aq_param = float(np.max(z)) if f_rew == 'exp_improve' else None
        np.random.seed(77)
        random.seed(78)
        with quiet(), np.errstate(all='ignore'):
            mcts = mc.cMCTS(budget, m, (5.0, 5.0, 0.3), 3, pg, fn, f_rew, 4, aq_param=aq_param, tree_type=tree_type)
            res = mcts.choose_trajectory(t=4)
        action, dense, best_val, paths, all_vals, max_locs, max_val, target = res
        out[tag + '_action'] = np.array(action)
        out[tag + '_best_val'] = np.array(best_val)
        out[tag + '_nq'] = np.array([c.nqueries for c in mcts.tree.root.children])
        out[tag + '_avg'] = np.array([c.reward / float(c.nqueries) for c in mcts.tree.root.children])
        out[tag + '_budget'] = budget
        if max_val is not None:
            out[tag + '_max_val'] = max_val
    np.savez_compressed(os.path.join(HERE, 'mcts.npz'), **out)


def case_myopic(gp, aq, pl):
    """Robot.choose_trajectory's inner loop (robot_library.py:173-205) for f_rew='mean'."""
    rng = np.random.default_rng(31)
    X, z = world(rng, 36, 1e-4)
    m = gp.OnlineGPModel(RANGES, 1.0, 100.0, noise=1e-4, dimension=2)
    m.add_data(X, z)
    pg = pl.Path_Generator(15, 1.5, 0.05, 0.5, RANGES)
    paths, true_paths = pg.get_path_set((4.0, 6.0, 1.1))
    keys = sorted(paths.keys())
    vals = np.array([aq.mean_UCB(time=9, xvals=paths[k], robot_model=m, param=None) for k in keys])
    flat = np.array([pt for k in keys for pt in paths[k]])
    offs = np.cumsum([0] + [len(paths[k]) for k in keys])
    np.savez_compressed(os.path.join(HERE, 'myopic.npz'), X=X, z=z, keys=np.array(keys), vals=vals, pts=flat,
                        offsets=offs, best=np.array(keys[int(np.argmax(vals))]))


def case_evaluation(gp, aq, pl, ev):
    """Evaluation.update_metrics (evaluation_library.py:262-314) for three evaluation rewards on one planning step:
    a 10 x 10 gridded world model, a 36-observation robot belief, the 15-path frontier set of one pose."""
    import types
    rng = np.random.default_rng(77)
    g = np.linspace(0.5, 9.5, 10)
    gx, gy = np.meshgrid(g, g)
    Xw = np.vstack([gx.ravel(), gy.ravel()]).T
    zw = 10.0 * np.sin(0.7 * Xw[:, :1]) * np.cos(0.5 * Xw[:, 1:2]) + 3.0 * np.cos(1.3 * Xw[:, :1] + 0.4 * Xw[:, 1:2])
    wm = gp.OnlineGPModel(RANGES, 1.0, 100.0, noise=1e-4, dimension=2)
    wm.add_data(Xw, zw)
    w = types.SimpleNamespace(dim=2, GP=wm, x1min=0., x1max=10., x2min=0., x2max=10.)
    X, z = world(rng, 36, 1e-4)
    m = gp.OnlineGPModel(RANGES, 1.0, 100.0, noise=1e-4, dimension=2)
    m.add_data(X, z)
    pg = pl.Path_Generator(15, 1.5, 0.05, 0.5, RANGES)
    paths, true_paths = pg.get_path_set((4.0, 6.0, 1.1))
    keys = sorted(paths.keys())
    sel = paths[keys[11]]
    max_loc = np.array([6.1, 2.2, 0.0])
    max_val = 11.5
    out = dict(Xw=Xw, zw=zw, X=X, z=z, keys=np.array(keys), sel=np.array(sel), max_loc=max_loc, max_val=max_val,
               pts=np.array([pt for k in keys for pt in paths[k]]),
               offsets=np.cumsum([0] + [len(paths[k]) for k in keys]))
    names = ['simple_regret', 'sample_regret_loc', 'sample_regret_val', 'max_loc_error', 'max_val_error',
             'instant_regret', 'max_val_regret', 'mes_reward_robot', 'mes_reward_omni', 'info_gain_reward', 'MSE',
             'hotspot_error']
    for rf in ('mean', 'hotspot_info', 'info_gain'):
        with quiet():
            e = ev.Evaluation(w, rf)
        np.random.seed(5)
        e.update_metrics(3, m, paths, sel, value=1.25, max_loc=max_loc, max_val=max_val,
                         params=[7.0, (1.0, 2.0), None, None], dist=4.5)
        out['metrics_' + rf] = np.array([float(e.metrics[n][3]) for n in names])
        r, vs, vm = e.inst_regret(3, paths, sel, m)
        out['regret_' + rf] = np.array([r, vs, vm], dtype=float)
    out['names'] = np.array(names)
    np.savez_compressed(os.path.join(HERE, 'evaluation.npz'), **out)


def case_obstacles(ob, pl):
    """obstacles.py in_obstacle predicates (:58-75, :145-163, :189-193) on random points and buffers -- including points
    placed exactly on the (buffered) edges -- and Dubins_Path_Generator.buffered_paths (paths_library.py:147-175)
    in a block world.  `dubins` is absent: the reference loop runs on the restated library (oracle/dubins_oracle.py)
    through an adapter with the dubins package's two calls."""
    rng = np.random.default_rng(9)
    ext = [0., 10., 0., 10.]
    with quiet():
        worlds = {'block': ob.BlockWorld(ext, num_blocks=3, dim_blocks=(2., 1.5), centers=[(3., 3.), (7., 6.5), (2.5, 8.)]),
                  'trap_left': ob.BugTrap(ext, (3., 3.), 2., 0.5, 2., 'left'),
                  'trap_right': ob.BugTrap(ext, (7., 6.), 1.5, 0.25, 3., 'right'),
                  'channel': ob.ChannelWorld(ext, (5., 4.), 3., 2.)}
    pts = rng.uniform(-1, 11, (3000, 2))
    edges = np.array([1.5, 2.0, 2.25, 2.5, 3.0, 3.5, 3.75, 4.0, 4.5, 5.0, 5.25, 5.5, 5.75, 6.0, 6.5, 6.75, 7.0, 7.25, 7.5,
                      8.0, 8.75, 0.0, 10.0])
    grid = np.array([(a, b) for a in edges for b in edges])
    pts = np.vstack([pts, grid])
    buffs = np.array([0.0, 0.1, 0.25, 0.5, 1.5])
    out = dict(pts=pts, buffs=buffs)
    for name, w in worlds.items():
        out[name] = np.array([[bool(w.in_obstacle((p[0], p[1]), buff=b)) for p in pts] for b in buffs])

    import types
    saved = pl.dubins
    pl.dubins = types.SimpleNamespace(shortest_path=_DubinsPath)
    try:
        poses = [(1.2, 1.0, 0.3), (4.6, 2.4, 1.9), (5.2, 5.1, 2.6), (8.6, 8.2, 4.0), (2.2, 6.0, 1.2), (6.0, 3.3, 5.5)]
        out['poses'] = np.array(poses)
        for name in ('block', 'trap_left', 'channel'):
            with quiet():
                pg = pl.Dubins_Path_Generator(15, 1.5, 0.05, 0.5, ext, worlds[name])
            for pi, pose in enumerate(poses):
                with quiet():
                    paths, true_paths = pg.get_path_set(pose)
                ns = np.zeros(len(pg.goals), dtype=np.int64)
                nd = np.zeros(len(pg.goals), dtype=np.int64)
                flat = []
                for i in range(len(pg.goals)):
                    if i in paths:
                        ns[i], nd[i] = len(paths[i]), len(true_paths[i])
                        flat.extend(paths[i])
                tag = 'dubins_%s_%d_' % (name, pi)
                out[tag + 'ns'], out[tag + 'nd'] = ns, nd
                out[tag + 'pts'] = np.array(flat).reshape(-1, 3)
    finally:
        pl.dubins = saved
    np.savez_compressed(os.path.join(HERE, 'obstacles.npz'), **out)


from mission_spec import (MISSION_METRICS, MISSIONS, mission_kwargs, mission_obstacles,  # noqa: E402
                          mission_sample_world, mission_world)


def case_missions(gp, pl, ob, ev, rb):
    """Robot.planner (robot_library.py:256-337) of the reference source, metrics by its Evaluation: whole seeded
    missions -- myopic and tree search, straight fan and Dubins (the absent `dubins` package replaced by the closed-form
    solver adapter), free and obstacle worlds.  Recorded: the pose after every step, the final belief data, the
    per-step metrics."""
    import types
    saved = pl.dubins
    pl.dubins = types.SimpleNamespace(shortest_path=_DubinsPath)
    out = {}
    try:
        for name, spec in MISSIONS.items():
            with quiet():
                world_obs = mission_obstacles(ob, spec[4])
                e = ev.Evaluation(mission_world(gp.OnlineGPModel), spec[0])
                np.random.seed(11)
                random.seed(12)
                r = rb.Robot(**mission_kwargs(name, e, world_obs, mission_sample_world))
                locs = []
                orig = r.collect_observations

                def spy(xobs, orig=orig, r=r, locs=locs):
                    orig(xobs)
                    locs.append(np.array(xobs[-1], dtype=float))
                r.collect_observations = spy
                with np.errstate(all='ignore'):
                    r.planner(T=spec[5])
            out[name + '_locs'] = np.array(locs)
            out[name + '_end'] = np.array([float(v) for v in r.loc])
            out[name + '_X'] = r.GP.xvals
            out[name + '_z'] = r.GP.zvals
            out[name + '_metrics'] = np.array([[float(e.metrics[n][t]) for n in MISSION_METRICS] for t in range(spec[5])])
    finally:
        pl.dubins = saved
    np.savez_compressed(os.path.join(HERE, 'missions.npz'), **out)


def case_environment(en):
    """Environment.__init__ / sample_value of the reference source (envmodel_library.py:27-260): the prior draw at the
    first grid point, the joint NUM_PTS^2 - 1 posterior sample, the regenerate-until-the-maximum-is-inside loop (seeds
    advance inside it), and one scalar noise draw per sample_value batch (Q9)."""
    out = {}
    cases = [(0, 12), (1, 12), (2, 12), (0, 20)]
    out['cases'] = np.array(cases)
    q = np.random.default_rng(21).uniform(0, 10, (9, 2))
    out['q'] = q
    for seed, npts in cases:
        with quiet():
            e = en.Environment(RANGES, npts, 100.0, 1.0, noise=0.5, visualize=False, seed=seed, dim=2)
        tag = 's%d_n%d_' % (seed, npts)
        out[tag + 'X'], out[tag + 'z'] = e.GP.xvals, e.GP.zvals
        out[tag + 'max_loc'], out[tag + 'max_val'] = np.asarray(e.max_loc, dtype=float), np.asarray(e.max_val, dtype=float)
        np.random.seed(100 + seed)
        out[tag + 'sample'] = e.sample_value(q)
        out[tag + 'after'] = np.array([np.random.random()])         # the random stream position after the call
    np.savez_compressed(os.path.join(HERE, 'environment.npz'), **out)


def main():
    import tempfile
    os.chdir(tempfile.mkdtemp())          # the reference's visualize=True branches mkdir ./figures
    gp = ref_loader.load('gpmodel_library')
    aq = ref_loader.load('aq_library')
    mc = ref_loader.load('mcts_library')
    pl = ref_loader.load('paths_library')
    case_online_gp(gp, 'online_gp_n05', 40, 2, 5, 0.5, 64, 0)
    case_online_gp(gp, 'online_gp_n1e4', 48, 3, 4, 1e-4, 64, 1)
    case_online_gp(gp, 'online_gp_mid', 300, 4, 4, 0.5, 256, 2)
    case_duplicates(gp)
    case_gpy_model(gp)
    case_rewards(gp, aq)
    case_max_vals(gp, aq)
    case_mcts(gp, aq, mc, pl)
    case_myopic(gp, aq, pl)
    case_evaluation(gp, aq, pl, ref_loader.load('evaluation_library'))
    case_obstacles(ref_loader.load('obstacles'), pl)
    case_environment(ref_loader.load('envmodel_library'))
    case_missions(gp, pl, ref_loader.load('obstacles'), ref_loader.load('evaluation_library'), ref_loader.load('robot_library'))
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == '__main__':
    main()
