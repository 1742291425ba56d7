"""Independent pin of the Dubins shortest-path solver behind the candidate-path generators.

The reference's `dubins.shortest_path(...).sample_many(step)` (paths_library.py:150-152) is an un-vendored C library;
oracle/dubins_oracle.py restates its published algorithm.  Nothing here trusts a closed form on its own word:
  * every word's pieces are re-integrated (rigid rotations about the circle centres; RK4 of the unicycle ODE on a
    subset) and must end on the goal pose to 1e-9;
  * the chosen length must be the minimum over a GEOMETRIC brute force of all six word families (tangent lines /
    tangent circles; oracle/dubins_oracle.geometric_candidates), on 12 000 random (pose, goal, rho) triples incl. the
    rho = 0.05 of BASELINE config 1;
  * exact ties go to the earlier word of the library's enum order LSL, LSR, RSL, RSR, RLR, LRL;
  * the package's Python solver, its native host generator (ipp_dubins_path_set) and -- on a GPU -- the device batch
    (ipp_dubins_paths_device) are compared with the oracle's solver + sample_many + the reference's trimming loop.
"""
import ctypes
import math

import numpy as np
import pytest

from oracle import dubins_oracle as do

EXT = [0., 10., 0., 10.]


def _triples(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    rhos = [0.05, 0.3, 1.0, 2.5]
    for i in range(n):
        rho = rhos[i % len(rhos)]
        q0 = (rng.uniform(0, 10), rng.uniform(0, 10), rng.uniform(-math.pi, math.pi))
        if i % 3 == 0:        # goals a few turning radii away: CCC words and overlapping circles are exercised
            r = rho * rng.uniform(0.05, 6.0)
            a = rng.uniform(-math.pi, math.pi)
            q1 = (q0[0] + r * math.cos(a), q0[1] + r * math.sin(a), rng.uniform(-math.pi, math.pi))
        elif i % 3 == 1:      # the planner's own shape: goal `horizon` away inside the +-2.35 rad fan, heading = bearing
            a = q0[2] + rng.uniform(-2.35, 2.35)
            h = rng.choice([1.5, 3.0, 5.0])
            q1 = (q0[0] + h * math.cos(a), q0[1] + h * math.sin(a), a)
        else:
            q1 = (rng.uniform(0, 10), rng.uniform(0, 10), rng.uniform(-math.pi, math.pi))
        out.append((q0, q1, rho))
    return out


def test_oracle_words_end_on_the_goal_and_the_choice_is_the_geometric_minimum():
    n_ccc = 0
    for q0, q1, rho in _triples(12000, 0):
        alpha, beta, d = do.normalise(q0, q1, rho)
        q0n = (0.0, 0.0, q0[2])
        q1n = ((q1[0] - q0[0]) / rho, (q1[1] - q0[1]) / rho, q1[2])
        scale = 1.0 + abs(q1n[0]) + abs(q1n[1])
        # the library's frame is rotated by theta: integrate in the rotated frame (0, 0, alpha) -> (d, 0, beta)
        for w in do.words(alpha, beta, d):
            if w is None:
                continue
            end = do.integrate_rigid((0.0, 0.0, alpha), w[0], w[1:])
            assert do.pose_error(end, (d, 0.0, beta)) <= 1e-9 * scale, (q0, q1, rho, w)
        word, prm = do.shortest_path(q0, q1, rho)
        end = do.integrate_rigid(q0n, word, prm)
        assert do.pose_error(end, q1n) <= 1e-9 * scale
        brute = do.brute_force_shortest(q0, q1, rho)
        assert brute, (q0, q1, rho)
        assert abs(sum(prm) - brute[0][0]) <= 1e-9 * (1.0 + brute[0][0]), (q0, q1, rho, word, prm, brute[:2])
        n_ccc += word in ('RLR', 'LRL')
    assert n_ccc > 50            # the three-arc families were actually chosen somewhere


def test_oracle_words_by_rk4_integration():
    for q0, q1, rho in _triples(300, 1):
        word, prm = do.shortest_path(q0, q1, rho)
        q1n = ((q1[0] - q0[0]) / rho, (q1[1] - q0[1]) / rho, q1[2])
        end = do.integrate_rk4((0.0, 0.0, q0[2]), word, prm, steps=400)
        assert do.pose_error(end, q1n) <= 1e-8 * (1.0 + abs(q1n[0]) + abs(q1n[1]))


def test_tie_order_is_the_librarys_enum_order():
    """Exact ties (cost equal bit for bit) go to the first word of LSL, LSR, RSL, RSR, RLR, LRL."""
    from informative_path_planning_b200.paths_library import dubins_shortest_path
    cases = [((0., 0., 0.), (4., 0., 0.), 1.0),            # straight ahead: all four CSC words are the same segment
             ((1., 1., math.pi / 2), (1., 6., math.pi / 2), 0.5),
             ((2., 2., 0.3), (2. + 3 * math.cos(0.3), 2. + 3 * math.sin(0.3), 0.3), 0.05)]
    for q0, q1, rho in cases:
        alpha, beta, d = do.normalise(q0, q1, rho)
        ws = [w for w in do.words(alpha, beta, d) if w is not None]
        best = min(sum(w[1:]) for w in ws)
        first = next(w for w in ws if sum(w[1:]) == best)
        assert do.shortest_path(q0, q1, rho)[0] == first[0]
        assert dubins_shortest_path(q0, q1, rho)[0] == first[0]
    # mirror-symmetric goal behind the vehicle: LSL/RSR (or LRL/RLR) tie mathematically; whichever the rounding prefers,
    # package and oracle must agree on it
    for q0, q1, rho in [((5., 5., 0.), (3., 5., math.pi), 1.0), ((5., 5., 0.), (5.2, 5., math.pi), 1.0)]:
        assert dubins_shortest_path(q0, q1, rho)[0] == do.shortest_path(q0, q1, rho)[0]


def test_package_solver_is_the_oracle_solver():
    from informative_path_planning_b200.paths_library import dubins_sample_many, dubins_shortest_path
    for q0, q1, rho in _triples(3000, 2):
        w_o, p_o = do.shortest_path(q0, q1, rho)
        w_p, p_p = dubins_shortest_path(q0, q1, rho)
        assert w_o == w_p, (q0, q1, rho)
        assert tuple(p_o) == tuple(p_p), (q0, q1, rho)       # same formulas, same operation order: bit for bit
    for q0, q1, rho in _triples(400, 3):
        w, p = do.shortest_path(q0, q1, rho)
        step = 0.05
        qs, _ = do.sample_many(q0, w, p, rho, step)
        got = dubins_sample_many(q0, w, p, rho, step)
        assert got.shape[0] == len(qs), (q0, q1, rho)
        np.testing.assert_allclose(got, np.array(qs).reshape(-1, 3), rtol=0, atol=1e-12)


def _reference_trimming(q0, goal, rho, ss, ext):
    """Dubins_Path_Generator.buffered_paths (paths_library.py:147-175) for one goal in a free world, on the oracle."""
    w, p = do.shortest_path(q0, goal, rho)
    fconfig, _ = do.sample_many(q0, w, p, rho, ss / 10.0)
    ftemp = []
    for c in fconfig:
        if c[0] > ext[0] and c[0] < ext[1] and c[1] > ext[2] and c[1] < ext[3]:
            ftemp.append(c)
        else:
            break
    ttemp = ftemp[0::10]
    for m, c in enumerate(ttemp):
        if c[0] <= ext[0] + 3 * rho or c[0] >= ext[1] - 3 * rho or c[1] <= ext[2] + 3 * rho or c[1] >= ext[3] - 3 * rho:
            ttemp = ttemp[0:m - 1]
    if len(ttemp) < 2:
        return None, None
    return ttemp, ftemp[0:ftemp.index(ttemp[-1]) + 1]


def _fan_goals(pose, horizon, rho, ext, fs=15):
    goals = []
    for a in np.linspace(-2.35, 2.35, fs):
        x = horizon * np.cos(pose[2] + a) + pose[0]
        y = horizon * np.sin(pose[2] + a) + pose[1]
        if x > ext[1] - 3 * rho or x < ext[0] + 3 * rho or y > ext[3] - 3 * rho or y < ext[2] + 3 * rho:
            continue
        goals.append((float(x), float(y), float(pose[2] + a)))
    return goals


def test_native_host_path_sets_against_the_oracle_generator():
    """ipp_dubins_path_set (C, host) == oracle solver + sample_many + the reference's trimming loop, for rho = 0.05
    (config 1) and a wide turning circle, poses in the open and against the walls."""
    from informative_path_planning_b200 import _lib as L
    rng = np.random.default_rng(4)
    vp = ctypes.c_void_p
    for rho, horizon, ss in ((0.05, 1.5, 0.5), (0.5, 3.0, 0.5), (1.0, 2.0, 0.25)):
        for trial in range(60):
            pose = (rng.uniform(0.2, 9.8), rng.uniform(0.2, 9.8), rng.uniform(-math.pi, math.pi))
            goals = _fan_goals(pose, horizon, rho, EXT)
            if not goals:
                continue
            G = len(goals)
            max_dense = int((horizon + 8.0 * math.pi * rho) / (ss / 10)) + 16
            max_samp = max_dense // 10 + 2
            samp, dense = np.empty((G, max_samp, 3)), np.empty((G, max_dense, 3))
            ns, nd = np.zeros(G, dtype=np.int32), np.zeros(G, dtype=np.int32)
            pose_a, goal_a, ext_a = np.array(pose), np.array(goals), np.array(EXT)
            L.check(L.lib.ipp_dubins_path_set(pose_a.ctypes.data_as(vp), goal_a.ctypes.data_as(vp), G, rho, ss / 10, 10,
                                              ext_a.ctypes.data_as(vp), 3 * rho, None, 0, samp.ctypes.data_as(vp),
                                              ns.ctypes.data_as(vp), max_samp, dense.ctypes.data_as(vp),
                                              nd.ctypes.data_as(vp), max_dense))
            for i, goal in enumerate(goals):
                tt, ft = _reference_trimming(pose, goal, rho, ss, EXT)
                if tt is None:
                    assert ns[i] < 2, (pose, goal, rho)
                    continue
                assert ns[i] == len(tt) and nd[i] == len(ft), (pose, goal, rho, ns[i], len(tt), nd[i], len(ft))
                np.testing.assert_allclose(samp[i, :ns[i]], np.array(tt), rtol=0, atol=1e-12)
                np.testing.assert_allclose(dense[i, :nd[i]], np.array(ft), rtol=0, atol=1e-12)


@pytest.mark.gpu
def test_device_path_batches_against_the_oracle_generator():
    """ipp_dubins_paths_device: one thread per (pose, goal) == the oracle generator, 4 000 pairs, three radii."""
    from informative_path_planning_b200 import _lib as L
    from informative_path_planning_b200.paths_library import dubins_paths_device
    rng = np.random.default_rng(5)
    for rho, horizon, ss in ((0.05, 1.5, 0.5), (0.3, 3.0, 0.5), (1.0, 2.0, 0.25)):
        P = 1400
        poses = np.column_stack([rng.uniform(0.2, 9.8, P), rng.uniform(0.2, 9.8, P), rng.uniform(-math.pi, math.pi, P)])
        ang = poses[:, 2] + rng.uniform(-2.35, 2.35, P)
        goals = np.column_stack([poses[:, 0] + horizon * np.cos(ang), poses[:, 1] + horizon * np.sin(ang), ang])
        max_samp = int((horizon + 8.0 * math.pi * rho) / ss) + 4
        samp_d, cnt_d, _ = dubins_paths_device(L.to_device(poses), L.to_device(goals), rho, ss, EXT, max_samples=max_samp)
        samp, cnt = samp_d.cpu().numpy(), cnt_d.cpu().numpy()
        for i in range(P):
            tt, _ = _reference_trimming(tuple(poses[i]), tuple(goals[i]), rho, ss, EXT)
            if tt is None:
                assert cnt[i] < 2, i
                continue
            assert cnt[i] == len(tt), (i, rho, cnt[i], len(tt))
            np.testing.assert_allclose(samp[i, :cnt[i]], np.array(tt), rtol=0, atol=1e-12)
