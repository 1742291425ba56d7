"""CPU: the native DPW tree search (csrc/tree.cu) against the Python Tree, without a GPU.

With the geometric `naive` reward (reference aq_library.py:339-367) a rollout needs no device work at all, so the whole
native search -- walk, progressive widening, Dubins candidate sets, prior draws, both Mersenne Twister streams,
back-propagation -- runs on the host and can be compared here with mcts_library.Tree driving the CPU oracle belief
through the reference's per-level sequence (reward, then add_data once or twice).  The GPU suite repeats the comparison
for the device-scored rewards."""
import contextlib
import ctypes
import io
import random

import numpy as np
import pytest

from conftest import RANGES


def _world(kind):
    from informative_path_planning_b200 import obstacles as ob
    ext = list(RANGES)
    return {'free': None, 'block': ob.BlockWorld(ext, num_blocks=2, dim_blocks=(1.5, 1.5), centers=[(6.5, 5.0), (4.0, 7.0)]),
            'trap': ob.BugTrap(ext, (3., 3.), 2., 0.5, 2., 'left')}[kind]


@pytest.mark.parametrize('tree_type', ['dpw', 'belief'])
@pytest.mark.parametrize('world_kind,pose,budget,depth,horizon,step', [
    ('free', (5.2, 4.6, 0.7), 250, 5, 1.5, 0.5), ('block', (5.0, 5.0, 0.0), 200, 4, 1.5, 0.5),
    ('trap', (8.9, 1.2, 2.2), 300, 7, 1.5, 0.5), ('free', (0.6, 9.4, -2.0), 120, 5, 3.0, 0.25),
    ('fan', (5.2, 4.6, 0.7), 250, 5, 1.5, 0.5), ('fan', (9.3, 0.8, 2.9), 200, 6, 1.7, 0.4)])
def test_native_search_equals_python_tree_on_the_host(world_kind, pose, budget, depth, horizon, step, tree_type):
    from informative_path_planning_b200 import _lib as L
    from informative_path_planning_b200 import mcts_library as mc
    from informative_path_planning_b200.obstacles import box_array
    from informative_path_planning_b200.paths_library import Dubins_Path_Generator, Path_Generator
    from oracle import aq_oracle
    from oracle.gpmodel_oracle import OnlineGPModelOracle
    rng = np.random.default_rng(7)
    X = rng.uniform(0, 10, (25, 2))
    z = 5 * np.sin(X[:, :1]) + rng.normal(0, 0.7, (25, 1))
    belief = OnlineGPModelOracle(RANGES, 1.0, 100.0, noise=0.5, dimension=2)
    belief.add_data(X, z)
    max_locs = np.array([[6.1, 5.2], [4.0, 6.4], [7.7, 2.1]])
    radius = 1.5
    param = ((np.array([[20.], [21.], [19.]]), max_locs, None), radius)
    fan = world_kind == 'fan'                                       # the straight fan of Path_Generator instead of Dubins curves
    world = None if fan else _world(world_kind)

    # ---- Python tree (reference per-level sequence on the CPU oracle belief)
    pg = (Path_Generator if fan else Dubins_Path_Generator)(15, horizon, 0.05, step, RANGES, world)
    np.random.seed(31); random.seed(32)
    np.random.standard_normal(1)                                   # leave a cached Gaussian behind
    planner = mc.cMCTS(budget, belief, pose, depth, pg, aq_oracle.naive, 'naive', 4, aq_param=(3, radius), tree_type=tree_type)
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all='ignore'):
        kids = planner.search(4, param)
    assert not planner.native_search_used
    want = ([c.nqueries for c in kids], [float(c.reward) for c in kids], random.random(), np.random.standard_normal(3))

    # ---- native search through the C ABI, host only
    vp = ctypes.c_void_p
    boxes, nb = box_array(pg.obstacle_world.boxes())
    angles = np.ascontiguousarray(np.linspace(-2.35, 2.35, 15))
    scale = mc._prior_scale_table(100.0)
    locs = np.ascontiguousarray(max_locs)
    prm = L.IppTreeParams()
    prm.budget, prm.max_depth, prm.kind = budget, depth, L.REWARD_NAIVE
    prm.frontier_size, prm.keep_every, prm.n_obstacles = 15, 10, nb
    prm.c, prm.reward_param, prm.time = float(planner.c), radius, 4.0
    prm.horizon, prm.turning_radius, prm.dense_step, prm.border = horizon, 0.05, float(step / 10), float(3 * 0.05)
    prm.generator, prm.sample_step = (1 if fan else 0), float(step)
    log_table = np.array([0.0] + [np.log(float(n)) for n in range(1, budget + 2)])
    if tree_type == 'belief':
        prm.tree_type, prm.log_table = 1, log_table.ctypes.data
    prm.extent[:] = list(RANGES)
    prm.angles, prm.prior_scale = angles.ctypes.data, scale.ctypes.data
    prm.obstacles = ctypes.cast(boxes, vp).value if nb else None
    prm.naive_locs, prm.n_naive = locs.ctypes.data, 3
    kern = L.make_rbf(2, 100.0, 1.0)
    np.random.seed(31); random.seed(32)
    np.random.standard_normal(1)
    py_state, py_extra = L.py_rng_state()
    np_state = L.np_rng_state()
    n_root = np.zeros(1, dtype=np.int32)
    goal = np.zeros(16, dtype=np.int32)
    visits = np.zeros(16, dtype=np.int64)
    reward = np.zeros(16)
    evals = np.zeros(1, dtype=np.int64)
    status = np.zeros(1, dtype=np.int32)
    pose_a = np.array(pose, dtype=np.float64)
    L.check(L.lib.ipp_tree_search_dpw(ctypes.byref(prm), pose_a.ctypes.data_as(vp), ctypes.byref(kern), None, 0, None, None, 0,
                                      0.5, 0.5 + 1e-8, None, 0, ctypes.byref(py_state), ctypes.byref(np_state), None, None, None,
                                      n_root.ctypes.data_as(vp), goal.ctypes.data_as(vp), visits.ctypes.data_as(vp),
                                      reward.ctypes.data_as(vp), evals.ctypes.data_as(vp), status.ctypes.data_as(vp), None))
    assert status[0] == 0
    L.set_py_rng_state(py_state, py_extra)
    L.set_np_rng_state(np_state)
    n = int(n_root[0])
    got = (visits[:n].tolist(), reward[:n].tolist(), random.random(), np.random.standard_normal(3))
    assert got[0] == want[0] and sum(got[0]) == budget
    assert got[1] == want[1]                                       # np.sum's pairwise order restated: bit-identical
    assert got[2] == want[2] and np.array_equal(got[3], want[3])   # both random streams left where the Python search leaves them
    assert evals[0] == planner.tree.reward_evals
    paths, _ = pg.get_path_set(pose)
    assert goal[:n].tolist() == list(paths.keys())
    assert max(got[0]) > min(got[0])                               # a non-trivial search: the visits concentrated somewhere


def test_native_search_rejects_bad_arguments():
    """Error behaviour of the C ABI: status code + ipp_last_error(), nothing touched."""
    from informative_path_planning_b200 import _lib as L
    from informative_path_planning_b200 import mcts_library as mc
    vp = ctypes.c_void_p
    angles = np.ascontiguousarray(np.linspace(-2.35, 2.35, 15))
    scale = mc._prior_scale_table(100.0)
    locs = np.array([[6.1, 5.2]])

    def params(**kw):
        prm = L.IppTreeParams()
        prm.budget, prm.max_depth, prm.kind = 10, 3, L.REWARD_NAIVE
        prm.frontier_size, prm.keep_every, prm.n_obstacles = 15, 10, 0
        prm.c, prm.reward_param = 1.0, 1.5
        prm.horizon, prm.turning_radius, prm.dense_step, prm.border = 1.5, 0.05, 0.05, 0.15
        prm.extent[:] = list(RANGES)
        prm.angles, prm.prior_scale = angles.ctypes.data, scale.ctypes.data
        prm.naive_locs, prm.n_naive = locs.ctypes.data, 1
        for k, v in kw.items():
            setattr(prm, k, v)
        return prm

    def call(prm, kern=None):
        kern = L.make_rbf(2, 100.0, 1.0) if kern is None else kern
        random.seed(1); np.random.seed(1)
        py_state, _ = L.py_rng_state()
        np_state = L.np_rng_state()
        before = (bytes(py_state.key), py_state.pos, bytes(np_state.key), np_state.pos)
        outs = [np.zeros(16, dtype=dt) for dt in (np.int32, np.int32, np.int64, np.float64, np.int64, np.int32)]
        pose = np.array([5.0, 5.0, 0.3])
        rc = L.lib.ipp_tree_search_dpw(ctypes.byref(prm), pose.ctypes.data_as(vp), ctypes.byref(kern), None, 0, None, None, 0,
                                       0.5, 0.5, None, 0, ctypes.byref(py_state), ctypes.byref(np_state), None, None, None,
                                       *[o.ctypes.data_as(vp) for o in outs], None)
        after = (bytes(py_state.key), py_state.pos, bytes(np_state.key), np_state.pos)
        return rc, L.lib.ipp_last_error().decode(), before == after, outs

    rc, msg, untouched, outs = call(params())
    assert rc == 0 and not untouched and int(outs[2][:outs[0][0]].sum()) == 10            # the good call works, host only
    for bad in (params(kind=7), params(n_naive=0), params(generator=1, sample_step=0.0), params(tree_type=1),
                params(frontier_size=0), params(n_obstacles=3), params(turning_radius=0.0)):
        rc, msg, untouched, _ = call(bad)
        assert rc != 0 and 'ipp_tree_search_dpw' in msg and untouched
    rc, msg, untouched, _ = call(params(kind=L.REWARD_UCB))                                # device reward without a belief
    assert rc != 0 and untouched
    with pytest.raises(L.IppError):
        L.check(rc)


def test_native_search_steps_aside_without_consuming_the_streams():
    """A rollout the one-pass scoring cannot take (here: 36 sample points on one path, the block limit is 32) makes the
    call report status 1 and leave both random streams exactly as they were, so the caller's own search starts from the
    same state."""
    from informative_path_planning_b200 import _lib as L
    from informative_path_planning_b200 import mcts_library as mc
    vp = ctypes.c_void_p
    angles = np.ascontiguousarray(np.linspace(-2.35, 2.35, 9))
    scale = mc._prior_scale_table(100.0)
    locs = np.array([[6.1, 5.2], [3.0, 4.0]])
    prm = L.IppTreeParams()
    prm.budget, prm.max_depth, prm.kind = 20, 3, L.REWARD_NAIVE
    prm.frontier_size, prm.keep_every, prm.n_obstacles = 9, 10, 0
    prm.c, prm.reward_param = 1.0, 1.5
    prm.horizon, prm.turning_radius, prm.dense_step, prm.border = 9.0, 0.05, 0.025, 0.15      # 9.0 / 0.25 = 36 samples
    prm.extent[:] = [0., 40., 0., 40.]
    prm.angles, prm.prior_scale = angles.ctypes.data, scale.ctypes.data
    prm.naive_locs, prm.n_naive = locs.ctypes.data, 2
    kern = L.make_rbf(2, 100.0, 1.0)
    random.seed(3); np.random.seed(4)
    py_state, _ = L.py_rng_state()
    np_state = L.np_rng_state()
    before = (bytes(py_state.key), py_state.pos, bytes(np_state.key), np_state.pos, np_state.has_gauss)
    outs = [np.zeros(16, dtype=dt) for dt in (np.int32, np.int32, np.int64, np.float64, np.int64, np.int32)]
    pose = np.array([20.0, 20.0, 0.3])
    L.check(L.lib.ipp_tree_search_dpw(ctypes.byref(prm), pose.ctypes.data_as(vp), ctypes.byref(kern), None, 0, None, None, 0,
                                      0.5, 0.5, None, 0, ctypes.byref(py_state), ctypes.byref(np_state), None, None, None,
                                      *[o.ctypes.data_as(vp) for o in outs], None))
    assert outs[5][0] == 1                                            # status: not done here
    assert before == (bytes(py_state.key), py_state.pos, bytes(np_state.key), np_state.pos, np_state.has_gauss)
