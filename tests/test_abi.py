"""CPU: the C-ABI library loads and exports every symbol include/ipp_b200.h declares; no compute calls."""
import ctypes
import os
import re

from conftest import ROOT

HEADER = os.path.join(ROOT, 'include', 'ipp_b200.h')
LIBRARY = os.path.join(ROOT, 'informative_path_planning_b200', 'csrc', 'libipp_b200.so')


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ipp_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(LIBRARY):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(LIBRARY)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), name


def test_python_binding_covers_the_header():
    from informative_path_planning_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert _lib.lib.ipp_version() == 100
    # workspace queries are pure host arithmetic: safe without a device
    assert _lib.lib.ipp_gp_predict_workspace(4096, 1 << 20) > 4096 * 32768
    assert _lib.lib.ipp_gp_factor_workspace(1024) >= 4 * 1024 * 1024


def test_struct_layout_matches_header():
    from informative_path_planning_b200 import _lib
    from informative_path_planning_b200.obstacles import IppObstacleBox
    assert ctypes.sizeof(_lib.IppRbf) == 40          # int32 + pad, double, double[3]
    sizes = (ctypes.c_int32 * 4)()
    assert _lib.lib.ipp_abi_struct_sizes(sizes) == 0
    assert list(sizes) == [ctypes.sizeof(_lib.IppRbf), ctypes.sizeof(IppObstacleBox), ctypes.sizeof(_lib.IppMtState),
                           ctypes.sizeof(_lib.IppTreeParams)]
    # the staging size is the header's own expression, evaluated with the binding's constants
    import re
    hdr = open(os.path.join(ROOT, 'include', 'ipp_b200.h')).read()
    expr = re.search(r'#define IPP_TREE_STAGE_DOUBLES \((.*)\)\s*$', hdr, re.M).group(1)
    assert _lib.IPP_TREE_STAGE_DOUBLES == eval(expr, {'IPP_ROLLOUT_MAX_POINTS': _lib.IPP_ROLLOUT_MAX_POINTS,
                                                      'IPP_MAX_DIM': _lib.IPP_MAX_DIM,
                                                      'IPP_ROLLOUT_MAX_LEVELS': _lib.IPP_ROLLOUT_MAX_LEVELS})
    k = _lib.make_rbf(2, 100.0, 1.0)
    assert (k.dim, k.variance, k.inv_lengthscale[0], k.inv_lengthscale[1], k.inv_lengthscale[2]) == (2, 100.0, 1.0, 1.0, 0.0)
    k3 = _lib.make_rbf(3, 4.0, (2.0, 4.0, 8.0))
    assert list(k3.inv_lengthscale) == [0.5, 0.25, 0.125]


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'informative_path_planning_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src.replace('np.random', ''), (f, 'product code must not reference oracle/')


def test_reference_ctor_errors_without_gpu():
    import pytest
    from informative_path_planning_b200.gpmodel_library import OnlineGPModel
    with pytest.raises(ValueError):
        OnlineGPModel((0, 10, 0, 10), 1.0, 100.0, dimension=4)
    with pytest.raises(ValueError):
        OnlineGPModel((0, 10, 0, 10), 1.0, 100.0, kernel='matern')
    with pytest.raises(ValueError):
        OnlineGPModel((0, 10, 0, 10), (1.0, 1.0), 100.0, dimension=3)
    m = OnlineGPModel((0, 10, 0, 10), 1.0, 100.0, noise=0.5)
    assert m.xvals is None and m.zvals is None and m.model is None
    import numpy as np
    mu, var = m.predict_value(np.zeros((3, 2)))          # no data: prior, no device needed
    assert mu.shape == (3, 1) and np.all(var == 100.0)
    with pytest.raises(AssertionError):
        m.predict_value(np.zeros((3, 3)))


def test_paths_library_matches_oracle_generator():
    import numpy as np
    from informative_path_planning_b200.paths_library import Path_Generator, Dubins_Path_Generator, dubins_shortest_path
    from oracle.mcts_oracle import PathGeneratorOracle
    a = Path_Generator(15, 1.5, 0.05, 0.5, (0, 10, 0, 10)).get_path_set((4.0, 6.0, 1.1))[0]
    b = PathGeneratorOracle(15, 1.5, 0.05, 0.5, (0, 10, 0, 10)).get_path_set((4.0, 6.0, 1.1))[0]
    assert sorted(a) == sorted(b)
    for k in a:
        np.testing.assert_allclose(np.array(a[k]), np.array(b[k]))
    # Dubins: straight-ahead goal is a straight line of the right length; every path ends near its goal
    word, (t, p, q) = dubins_shortest_path((0, 0, 0), (4, 0, 0), 1.0)
    assert abs(t + p + q - 4.0) < 1e-12
    # native (C, host) generator == Python generator, including the border-trimming quirk near the walls
    for pose in ((5.0, 5.0, 0.3), (9.6, 9.5, 0.3), (0.4, 5.0, 3.0), (5.0, 0.2, -1.4), (9.9, 0.1, 2.0)):
        a_gen = Dubins_Path_Generator(15, 1.5, 0.05, 0.5, (0, 10, 0, 10))
        b_gen = Dubins_Path_Generator(15, 1.5, 0.05, 0.5, (0, 10, 0, 10))
        b_gen.use_native = False
        (pa, ta), (pb, tb) = a_gen.get_path_set(pose), b_gen.get_path_set(pose)
        assert sorted(pa) == sorted(pb), pose
        for k in pa:
            np.testing.assert_allclose(np.array(pa[k]), np.array(pb[k]), rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(np.array(ta[k]), np.array(tb[k]), rtol=1e-12, atol=1e-12)
    g = Dubins_Path_Generator(15, 1.5, 0.05, 0.5, (0, 10, 0, 10))
    paths, true_paths = g.get_path_set((5.0, 5.0, 0.3))
    assert len(paths) >= 10
    for k, pts in paths.items():
        goal = g.goals[k]
        end = true_paths[k][-1]
        assert np.hypot(end[0] - goal[0], end[1] - goal[1]) < 0.5 + 1e-9      # samples every ss; the end point itself is excluded
        steps = np.hypot(np.diff([c[0] for c in pts]), np.diff([c[1] for c in pts]))
        assert np.all(steps < 0.5 + 1e-6)


def test_prior_draw_is_numpys_multivariate_normal_bit_for_bit():
    """The planner's fast prior sampler must consume np.random exactly like the reference's
    np.random.multivariate_normal(zeros(k), eye(k) * variance) call (mcts_library.py:418-422)."""
    import numpy as np
    from informative_path_planning_b200 import mcts_library as mc
    for k in (1, 3, 4, 7, 20):
        for var in (100.0, 2.5):
            np.random.seed(5)
            want = [np.random.multivariate_normal(mean=np.zeros((k,)), cov=np.eye(k) * var) for _ in range(4)]
            after_want = np.random.random()
            np.random.seed(5)
            got = [mc._prior_draw(k, var) for _ in range(4)]
            after_got = np.random.random()
            assert all(np.array_equal(a, b) for a, b in zip(want, got))
            assert after_want == after_got


def test_lazy_dense_paths_equal_eager_ones():
    import numpy as np
    from informative_path_planning_b200.paths_library import Dubins_Path_Generator
    for pose in ((5.0, 5.0, 0.3), (9.6, 9.5, 0.3), (0.4, 5.0, 3.0)):
        g = Dubins_Path_Generator(15, 1.5, 0.05, 0.5, (0, 10, 0, 10))
        pa, ta = g.get_path_set(pose)
        pb, tb = g.get_path_set_lazy(pose)
        assert pa == pb and sorted(ta) == sorted(tb)
        for k in ta:
            assert ta[k] == tb[k].tolist() == list(tb[k]) and ta[k][-1] == tb[k][-1] and len(ta[k]) == len(tb[k])
            assert ta[k][2:5] == tb[k][2:5]


def _golden_worlds():
    from informative_path_planning_b200 import obstacles as ob
    ext = [0., 10., 0., 10.]
    return {'block': ob.BlockWorld(ext, num_blocks=3, dim_blocks=(2., 1.5), centers=[(3., 3.), (7., 6.5), (2.5, 8.)]),
            'trap_left': ob.BugTrap(ext, (3., 3.), 2., 0.5, 2., 'left'),
            'trap_right': ob.BugTrap(ext, (7., 6.), 1.5, 0.25, 3., 'right'),
            'channel': ob.ChannelWorld(ext, (5., 4.), 3., 2.)}


def test_obstacle_worlds_match_reference_predicates():
    """in_obstacle of every world, and the box-with-hole list the native generators evaluate, against the reference
    source's own answers (tests/golden/obstacles.npz: random points + points exactly on the buffered edges)."""
    import numpy as np
    from conftest import load_golden
    from informative_path_planning_b200 import obstacles as ob
    g = load_golden('obstacles')
    for name, w in _golden_worlds().items():
        boxes = w.boxes()
        for bi, buff in enumerate(g['buffs']):
            got = np.array([w.in_obstacle((p[0], p[1]), buff=float(buff)) for p in g['pts']])
            via_boxes = np.array([ob.in_boxes(boxes, p, float(buff)) for p in g['pts']])
            assert np.array_equal(got, g[name][bi]), (name, buff)
            assert np.array_equal(via_boxes, g[name][bi]), (name, buff)
    assert ctypes.sizeof(ob.IppObstacleBox) == 104
    free = ob.FreeWorld()
    assert free.boxes() == [] and not free.in_obstacle((1, 1)) and free.get_obstacles() == []
    np.random.seed(3)
    a = ob.BlockWorld([0., 10., 0., 10.], num_blocks=2)
    np.random.seed(3)
    want = [tuple(10 * np.random.random_sample(2)) for _ in range(2)]
    np.testing.assert_allclose(np.array(a.get_centers()), np.array(want))


def test_native_dubins_path_sets_in_obstacle_worlds():
    """ipp_dubins_path_set with an obstacle list == the reference's trimming loop (golden: the reference source run on
    the closed-form solver) == this package's Python loop; lazy dense paths agree as well."""
    import numpy as np
    from conftest import load_golden
    from informative_path_planning_b200.paths_library import Dubins_Path_Generator
    g = load_golden('obstacles')
    worlds = _golden_worlds()
    for name in ('block', 'trap_left', 'channel'):
        for pi, pose in enumerate(g['poses']):
            pose = tuple(float(v) for v in pose)
            nat = Dubins_Path_Generator(15, 1.5, 0.05, 0.5, [0., 10., 0., 10.], worlds[name])
            py = Dubins_Path_Generator(15, 1.5, 0.05, 0.5, [0., 10., 0., 10.], worlds[name])
            py.use_native = False
            (pa, ta), (pb, tb) = nat.get_path_set(pose), py.get_path_set(pose)
            tag = 'dubins_%s_%d_' % (name, pi)
            ns, nd, flat = g[tag + 'ns'], g[tag + 'nd'], g[tag + 'pts']
            assert sorted(pa) == sorted(pb) == [i for i in range(len(ns)) if ns[i] > 0], (name, pose)
            off = 0
            for k in sorted(pa):
                assert len(pa[k]) == ns[k] and len(ta[k]) == len(tb[k]) == nd[k], (name, pose, k)
                np.testing.assert_allclose(np.array(pa[k]), flat[off:off + ns[k]], rtol=1e-12, atol=1e-12)
                np.testing.assert_allclose(np.array(pa[k]), np.array(pb[k]), rtol=1e-12, atol=1e-12)
                np.testing.assert_allclose(np.array(ta[k]), np.array(tb[k]), rtol=1e-12, atol=1e-12)
                off += ns[k]
            pl_, tl_ = nat.get_path_set_lazy(pose)
            assert pl_ == pa and all(tl_[k].tolist() == ta[k] for k in ta)

    class Duck(object):                          # a world without boxes() keeps working through the Python loop
        def in_obstacle(self, point, buff=0.1):
            return worlds['block'].in_obstacle(point, buff)
    duck = Dubins_Path_Generator(15, 1.5, 0.05, 0.5, [0., 10., 0., 10.], Duck())
    ref = Dubins_Path_Generator(15, 1.5, 0.05, 0.5, [0., 10., 0., 10.], worlds['block'])
    assert duck.get_path_set((4.6, 2.4, 1.9))[0].keys() == ref.get_path_set((4.6, 2.4, 1.9))[0].keys()


def test_native_random_streams_continue_pythons_and_numpys():
    """The tree search's generators (csrc/tree.cu) against the interpreters': random.choice and
    np.random.standard_normal continued bit for bit from an arbitrary state (cached Gaussian included), and the state
    handed back continues the interpreter's stream."""
    import random
    import numpy as np
    from informative_path_planning_b200 import _lib as L
    vp = ctypes.c_void_p
    random.seed(123)
    [random.random() for _ in range(7)]
    st, extra = L.py_rng_state()
    ns = np.random.default_rng(0).integers(1, 70, 6000).astype(np.int32)
    ns[:40] = np.arange(1, 41)
    out = np.zeros(ns.size, dtype=np.int32)
    L.check(L.lib.ipp_rng_py_choices(ctypes.byref(st), ns.ctypes.data_as(vp), int(ns.size), out.ctypes.data_as(vp)))
    want = np.array([random.choice(range(int(n))) for n in ns])
    assert np.array_equal(out, want)
    nxt = [random.random(), random.choice([3, 4, 5])]
    L.set_py_rng_state(st, extra)
    assert [random.random(), random.choice([3, 4, 5])] == nxt
    for warm in (0, 3):                                   # 3: an odd number of draws leaves a cached second variate
        np.random.seed(77)
        np.random.standard_normal(warm)
        st = L.np_rng_state()
        assert st.has_gauss == warm % 2
        got = np.zeros(10001)
        L.check(L.lib.ipp_rng_np_normals(ctypes.byref(st), 10001, got.ctypes.data_as(vp)))
        assert np.array_equal(got, np.random.standard_normal(10001))
        nxt = np.random.standard_normal(5)
        L.set_np_rng_state(st)
        assert np.array_equal(np.random.standard_normal(5), nxt)


def test_native_tree_path_sets_equal_the_generators():
    """ipp_tree_path_set (frontier rule + Dubins sets inside the native tree search) == Dubins_Path_Generator
    .get_path_set, bit for bit, in free and obstacle worlds, at interior poses and against the walls."""
    import numpy as np
    from informative_path_planning_b200 import _lib as L
    from informative_path_planning_b200 import mcts_library as mc
    from informative_path_planning_b200.obstacles import box_array
    from informative_path_planning_b200.paths_library import Dubins_Path_Generator
    vp = ctypes.c_void_p
    rng = np.random.default_rng(4)
    worlds = _golden_worlds()
    for world in (None, worlds['block'], worlds['trap_left'], worlds['channel']):
        pg = Dubins_Path_Generator(15, 1.5, 0.05, 0.5, [0., 10., 0., 10.], world)
        boxes, nb = box_array(pg.obstacle_world.boxes())
        angles = np.ascontiguousarray(np.linspace(-2.35, 2.35, 15))
        scale = mc._prior_scale_table(100.0)
        prm = L.IppTreeParams()
        prm.budget, prm.max_depth, prm.kind, prm.frontier_size, prm.keep_every, prm.n_obstacles = 1, 3, 0, 15, 10, nb
        prm.horizon, prm.turning_radius, prm.dense_step, prm.border = 1.5, 0.05, float(0.5 / 10), float(3 * 0.05)
        prm.extent[:] = [0., 10., 0., 10.]
        prm.angles, prm.prior_scale = angles.ctypes.data, scale.ctypes.data
        prm.obstacles = ctypes.cast(boxes, vp).value if nb else None
        for _ in range(40):
            pose = (float(rng.uniform(0.2, 9.8)), float(rng.uniform(0.2, 9.8)), float(rng.uniform(-4, 4)))
            paths, _ = pg.get_path_set(pose)
            n = np.zeros(1, dtype=np.int32)
            goal = np.zeros(16, dtype=np.int32)
            cnt = np.zeros(16, dtype=np.int32)
            samp = np.zeros((16, 16, 3))
            pose_a = np.array(pose)
            L.check(L.lib.ipp_tree_path_set(ctypes.byref(prm), pose_a.ctypes.data_as(vp), n.ctypes.data_as(vp),
                                            goal.ctypes.data_as(vp), cnt.ctypes.data_as(vp), samp.ctypes.data_as(vp), 16))
            assert list(paths.keys()) == goal[:n[0]].tolist(), pose
            for i, key in enumerate(paths):
                assert np.array_equal(np.array(paths[key]), samp[i, :cnt[i]]), (pose, key)


def test_stale_library_is_refused(monkeypatch):
    """The built library carries a stamp of the sources it was built from (csrc/Makefile); _lib refuses to load one whose
    stamp does not match the sources beside it."""
    import pytest
    from informative_path_planning_b200 import _lib as L
    monkeypatch.delenv('IPP_B200_LIB', raising=False)
    L._check_source_stamp()                                   # the library in the tree is current
    assert open(L.LIB_PATH + '.srchash').read().strip() == L.source_hash()
    monkeypatch.setattr(L, 'source_hash', lambda: 'edited-after-the-build')
    with pytest.raises(ImportError, match='was not built from the current'):
        L._check_source_stamp()


def test_new_entry_points_reject_bad_arguments_before_touching_the_device():
    """Argument validation of the entry points added in round 2 happens on the host: null pointers and impossible shapes
    come back as IPP_E_INVALID with a message, on a machine without a GPU too (no launch is attempted)."""
    import ctypes
    from informative_path_planning_b200 import _lib as L
    null = ctypes.c_void_p(0)
    one = ctypes.c_void_p(16)                    # never dereferenced: validation fails first
    rc = L.lib.ipp_rff_theta_draw(null, one, 200, 2, one, one, 300, 100.0, 0.5, one, 1, one, one, one, null)
    assert rc != 0 and b'null pointer' in L.lib.ipp_last_error()
    rc = L.lib.ipp_rff_theta_draw(one, one, 200, 4, one, one, 300, 100.0, 0.5, one, 1, one, one, one, null)
    assert rc != 0 and b'bad arguments' in L.lib.ipp_last_error()
    rc = L.lib.ipp_rff_theta_draw(one, one, 200, 2, one, one, 300, 100.0, 0.0, one, 1, one, one, one, null)      # noise 0
    assert rc != 0
    rc = L.lib.ipp_rff_eval_argmax_mesh(one, one, one, 100, 3, 2, one, 10, one, 10, 0.0, null, one, one, one, null)   # S < 8
    assert rc != 0 and b'at least 8 samples' in L.lib.ipp_last_error()
    rc = L.lib.ipp_rff_eval_argmax_mesh(one, one, one, 100, 16, 2, one, 0, one, 10, 0.0, null, one, one, one, null)
    assert rc != 0 and b'bad mesh' in L.lib.ipp_last_error()
    assert L.lib.ipp_rff_theta_workspace(200, 300, 1) > 0 and L.lib.ipp_rff_theta_workspace(200, 300, 100) > L.lib.ipp_rff_theta_workspace(200, 300, 1)
    assert L.lib.ipp_rff_mesh_workspace(1000, 100, 1024, 1024) > L.lib.ipp_rff_workspace(1000, 100, 1 << 20, 1)
