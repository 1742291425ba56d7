"""Candidate-path generators (reference paths_library.py), restated for Python 3 and without the
un-vendored `dubins` C library.

Path_Generator         reference :18-139   frontier fan of straight segments
Dubins_Path_Generator  reference :141-189  shortest Dubins curves to the frontier goals, sampled densely
                                           (step ss/10), every 10th point kept, trimmed at the border

The Dubins solver below is the classical six-word closed form (LSL, LSR, RSL, RSR, RLR, LRL; Shkel & Lumelsky
2001), the algorithm the `dubins` package wraps; shortest word wins.  Obstacle worlds are the reference's
FreeWorld contract: any object with in_obstacle(point, buff) -> bool; the worlds of this package's obstacles.py also
expose boxes(), which keeps them on the native / device generators.
"""
import math

import numpy as np

from .obstacles import FreeWorld, box_array  # noqa: F401  (FreeWorld re-exported: reference paths_library uses obs.FreeWorld)


def py2_round(x):
    """int(round(x)) as the reference's Python 2 computes it (C round(): halves away from zero; Python 3 and numpy round
    halves to even) -- the sample count of a straight path differs when distance / sample_step lands on x.5."""
    x = float(x)
    f = np.floor(abs(x))
    r = f + 1.0 if abs(x) - f >= 0.5 else f
    return int(-r if x < 0 else r)


class LazyPath(object):
    """A dense path that is only materialised on demand.  The tree search only ever looks at dense_path[-1] (the pose
    of the next belief node) for all but the chosen action, so the ~30-100 dense configurations per candidate path per
    expanded node are neither computed nor turned into tuples unless somebody asks: `end` is the last configuration,
    `factory()` produces the whole list of (x, y, heading) tuples."""

    __slots__ = ('end', 'factory', '_items')

    def __init__(self, end, factory):
        self.end = end
        self.factory = factory
        self._items = None

    def tolist(self):
        if self._items is None:
            self._items = self.factory()
        return self._items

    def __len__(self):
        return len(self.tolist())

    def __getitem__(self, i):
        if i == -1:
            return self.end
        return self.tolist()[i]

    def __iter__(self):
        return iter(self.tolist())


class Path_Generator(object):
    def __init__(self, frontier_size, horizon_length, turning_radius, sample_step, extent, obstacle_world=None):
        self.fs = frontier_size
        self.hl = horizon_length
        self.tr = turning_radius
        self.ss = sample_step
        self.extent = extent
        self.goals = []
        self.samples = {}
        self.cp = (0, 0, 0)
        self.obstacle_world = FreeWorld() if obstacle_world is None else obstacle_world

    def generate_frontier_points(self):
        """reference :44-83 (plain-float arithmetic: this runs once per expanded tree node; for a 15-goal fan the scalar
        loop beats numpy's per-call overhead -- 27 vs 50 us)"""
        goals = []
        angles = self.__dict__.get('_angles')
        if angles is None or len(angles) != self.fs:
            angles = self._angles = [float(a) for a in np.linspace(-2.35, 2.35, self.fs)]
        cx, cy, ch = float(self.cp[0]), float(self.cp[1]), float(self.cp[2])
        hl, tr = self.hl, self.tr
        x_lo, x_hi = self.extent[0] + 3 * tr, self.extent[1] - 3 * tr
        y_lo, y_hi = self.extent[2] + 3 * tr, self.extent[3] - 3 * tr
        cos, sin, sqrt = math.cos, math.sin, math.sqrt
        for a in angles:
            p = ch + a
            x = hl * cos(p) + cx
            y = hl * sin(p) + cy
            if sqrt((cx - x) * (cx - x) + (cy - y) * (cy - y)) <= tr:
                continue
            if x > x_hi or x < x_lo:
                continue
            if y > y_hi or y < y_lo:
                continue
            goals.append((x, y, p))
        goals.append(self.cp)
        self.goals = goals
        return self.goals

    def make_sample_paths(self):
        """reference :85-105"""
        cp = np.array(self.cp)
        coords = {}
        for i, goal in enumerate(self.goals):
            g = np.array(goal)
            distance = np.sqrt((cp[0] - g[0]) ** 2 + (cp[1] - g[1]) ** 2)
            for j in range(py2_round(distance / self.ss)):
                step = (j + 1) * self.ss
                coords.setdefault(i, []).append((cp[0] + step * np.cos(g[2]), cp[1] + step * np.sin(g[2]), g[2]))
        self.samples = coords
        return self.samples, self.samples

    def get_path_set(self, current_pose):
        self.cp = current_pose
        self.generate_frontier_points()
        return self.make_sample_paths()

    def path_cost(self, path, loc=None):
        """reference :119-132"""
        dist = 0
        if loc is None:
            for i in range(len(path) - 1):
                dist += np.sqrt((path[i][0] - path[i + 1][0]) ** 2 + (path[i][1] - path[i + 1][1]) ** 2)
            return dist
        for coord in path:
            dist += np.sqrt((coord[0] - loc[0]) ** 2 + (coord[1] - loc[1]) ** 2)
        return dist / len(path)

    def get_frontier_points(self):
        return self.goals

    def get_sample_points(self):
        return self.samples


# ------------------------------------------------------------------------------------------ Dubins curves
def _mod2pi(t):
    return t - 2.0 * math.pi * math.floor(t / (2.0 * math.pi))


def _dubins_words(alpha, beta, d):
    """The six words in the order of the `dubins` C library's DubinsPathType enum (LSL, LSR, RSL, RSR, RLR, LRL): the
    order is the tie-break, because the library keeps a word only when it is strictly shorter than the best so far."""
    sa, sb, ca, cb, cab = math.sin(alpha), math.sin(beta), math.cos(alpha), math.cos(beta), math.cos(alpha - beta)
    d_sq = d * d
    out = []
    p2 = 2 + d_sq - (2 * cab) + (2 * d * (sa - sb))                      # LSL
    if p2 >= 0:
        tmp = math.atan2(cb - ca, d + sa - sb)
        out.append(('LSL', _mod2pi(tmp - alpha), math.sqrt(p2), _mod2pi(beta - tmp)))
    p2 = -2 + d_sq + (2 * cab) + (2 * d * (sa + sb))                     # LSR
    if p2 >= 0:
        p = math.sqrt(p2)
        tmp = math.atan2(-ca - cb, d + sa + sb) - math.atan2(-2.0, p)
        out.append(('LSR', _mod2pi(tmp - alpha), p, _mod2pi(tmp - _mod2pi(beta))))
    p2 = -2 + d_sq + (2 * cab) - (2 * d * (sa + sb))                     # RSL
    if p2 >= 0:
        p = math.sqrt(p2)
        tmp = math.atan2(ca + cb, d - sa - sb) - math.atan2(2.0, p)
        out.append(('RSL', _mod2pi(alpha - tmp), p, _mod2pi(beta - tmp)))
    p2 = 2 + d_sq - (2 * cab) + (2 * d * (sb - sa))                      # RSR
    if p2 >= 0:
        tmp = math.atan2(ca - cb, d - sa + sb)
        out.append(('RSR', _mod2pi(alpha - tmp), math.sqrt(p2), _mod2pi(tmp - beta)))
    tmp = (6.0 - d_sq + 2 * cab + 2 * d * (sa - sb)) / 8.0               # RLR
    if abs(tmp) <= 1:
        phi = math.atan2(ca - cb, d - sa + sb)
        p = _mod2pi(2 * math.pi - math.acos(tmp))
        t = _mod2pi(alpha - phi + _mod2pi(p / 2.0))
        out.append(('RLR', t, p, _mod2pi(alpha - beta - t + _mod2pi(p))))
    tmp = (6.0 - d_sq + 2 * cab + 2 * d * (sb - sa)) / 8.0               # LRL
    if abs(tmp) <= 1:
        phi = math.atan2(ca - cb, d + sa - sb)
        p = _mod2pi(2 * math.pi - math.acos(tmp))
        t = _mod2pi(-alpha - phi + p / 2.0)
        out.append(('LRL', t, p, _mod2pi(_mod2pi(beta) - alpha - t + _mod2pi(p))))
    return out


def dubins_shortest_path(q0, q1, rho):
    """Shortest Dubins word between poses q0 and q1 = (x, y, heading); returns (word, (t, p, q)) in units of rho.
    Restates the `dubins` library's dubins_shortest_path (the call at reference paths_library.py:151): first word of the
    enum order with the strictly smallest cost."""
    dx, dy = q1[0] - q0[0], q1[1] - q0[1]
    D = math.sqrt(dx * dx + dy * dy)
    d = D / rho
    theta = _mod2pi(math.atan2(dy, dx)) if d > 0 else 0.0
    alpha, beta = _mod2pi(q0[2] - theta), _mod2pi(q1[2] - theta)
    best, best_cost = None, float('inf')
    for w in _dubins_words(alpha, beta, d):
        cost = w[1] + w[2] + w[3]
        if cost < best_cost:
            best, best_cost = w, cost
    return best[0], best[1:]


def _segment_vec(x, y, th, length, kind):
    """The library's dubins_segment, vectorised over `length`: offsets first, start configuration added last."""
    st, ct = np.sin(th), np.cos(th)
    if kind == 'L':
        return (np.sin(th + length) - st) + x, (-np.cos(th + length) + ct) + y, length + th
    if kind == 'R':
        return (-np.sin(th - length) + st) + x, (np.cos(th - length) - ct) + y, -length + th
    return ct * length + x, st * length + y, th + 0.0 * length


def dubins_sample_many(q0, word, params, rho, step):
    """Configurations every `step` along the path (the `dubins` package's sample_many, reference paths_library.py:152),
    vectorised over the arc-length samples: returns an (n, 3) array of (x, y, heading).  Like the library: x = 0;
    while x < length: sample(x); x += step -- the accumulated rounding decides whether the end point is sampled -- and
    each sample picks its piece by `t' < p1`, `t' < p1 + p2`."""
    total = (params[0] + params[1] + params[2]) * rho
    n = int(math.ceil(total / step)) + 2 if total > 0 else 0
    s = np.concatenate([[0.0], np.cumsum(np.full(max(n - 1, 0), float(step)))])[:n]
    s = s[s < total]
    tp = s / rho
    p1, p2 = params[0], params[1]
    z = np.zeros(1)
    x1, y1, th1 = _segment_vec(z, z, z + q0[2], z + p1, word[0])
    x2, y2, th2 = _segment_vec(x1, y1, th1, z + p2, word[1])
    first, second = tp < p1, tp < (p1 + p2)
    xa, ya, tha = _segment_vec(0.0, 0.0, q0[2], tp, word[0])
    xb, yb, thb = _segment_vec(x1[0], y1[0], th1[0], tp - p1, word[1])
    xc, yc, thc = _segment_vec(x2[0], y2[0], th2[0], tp - p1 - p2, word[2])
    x = np.where(first, xa, np.where(second, xb, xc))
    y = np.where(first, ya, np.where(second, yb, yc))
    th = np.where(first, tha, np.where(second, thb, thc))
    out = np.empty((s.shape[0], 3))
    out[:, 0] = x * rho + q0[0]
    out[:, 1] = y * rho + q0[1]
    out[:, 2] = th - 2.0 * math.pi * np.floor(th / (2.0 * math.pi))
    return out


class Dubins_Path_Generator(Path_Generator):
    """reference :141-189.  The whole path set of a pose is produced by one native call (ipp_dubins_path_set, the
    counterpart of the `dubins` C library the reference uses) for every obstacle world that can state itself as a
    box list (obstacles.py: FreeWorld, BlockWorld, BugTrap, ChannelWorld); any other duck-typed world keeps the
    Python loop below, which is also the cross-check of the native code (tests/test_abi.py)."""

    use_native = True

    def _obstacle_boxes(self):
        """(ctypes box array or None, count) of the obstacle world, or None when it cannot be stated as boxes."""
        cached = self.__dict__.get('_boxes')
        if cached is not None and cached[0] is self.obstacle_world:
            return cached[1]
        fn = getattr(self.obstacle_world, 'boxes', None)
        out = box_array(fn()) if callable(fn) else None
        self._boxes = (self.obstacle_world, out)
        return out

    def buffered_paths(self, lazy=False):
        if self.use_native and len(self.goals) > 0 and self._obstacle_boxes() is not None:
            return self._buffered_paths_native(lazy)
        return self._buffered_paths_python()

    def get_path_set_lazy(self, current_pose):
        """get_path_set for the tree search: dense paths come back as LazyPath views (same values)."""
        self.cp = current_pose
        self.generate_frontier_points()
        coords, true_coords = self.buffered_paths(lazy=True)
        self.samples = coords
        return coords, true_coords

    def _native_buffers(self, G):
        """Reusable host buffers + their ctypes pointers for ipp_dubins_path_set (allocated once per generator and goal
        count: the wrapper used to cost 4x the native call in allocations and pointer casts)."""
        import ctypes
        nb = self.__dict__.get('_nb')
        if nb is None or nb['G'] < G:
            # longest shortest-Dubins path to a goal hl away: hl + a few turning circles
            max_dense = int((self.hl + 8.0 * math.pi * self.tr) / (self.ss / 10)) + 16
            max_samp = max_dense // 10 + 2
            nb = {'G': G, 'max_dense': max_dense, 'max_samp': max_samp,
                  'goals': np.empty((G, 3)), 'pose': np.empty(3), 'ext': np.asarray(self.extent, dtype=np.float64).copy(),
                  'samp': np.empty((G, max_samp, 3)), 'dense': np.empty((G, max_dense, 3)),
                  'ns': np.zeros(G, dtype=np.int32), 'nd': np.zeros(G, dtype=np.int32)}
            vp = ctypes.c_void_p
            nb['ptr'] = {k: nb[k].ctypes.data_as(vp) for k in ('goals', 'pose', 'ext', 'samp', 'dense', 'ns', 'nd')}
            self._nb = nb
        return nb

    def _buffered_paths_native(self, lazy=False):
        from . import _lib as L
        G = len(self.goals)
        nb = self._native_buffers(G)
        nb['goals'][:G] = self.goals
        nb['pose'][:] = self.cp
        P = nb['ptr']
        boxes, n_boxes = self._obstacle_boxes()
        L.check(L.lib.ipp_dubins_path_set(P['pose'], P['goals'], G, float(self.tr), float(self.ss / 10), 10, P['ext'],
                                          float(3 * self.tr), boxes, n_boxes, P['samp'], P['ns'], nb['max_samp'],
                                          None if lazy else P['dense'], P['nd'], nb['max_dense']))
        ns = nb['ns'][:G].tolist()
        samp = nb['samp'].tolist()
        sampling_path, true_path = {}, {}
        if lazy:
            pose, goals = self.cp, list(self.goals)
            for i in range(G):
                if ns[i] >= 2:
                    pts = [tuple(c) for c in samp[i][:ns[i]]]
                    sampling_path[i] = pts
                    true_path[i] = LazyPath(pts[-1], lambda pose=pose, goal=goals[i]: self._dense_for(pose, goal))
            return sampling_path, true_path
        nd, dense = nb['nd'][:G].tolist(), nb['dense']
        for i in range(G):
            if ns[i] >= 2:
                sampling_path[i] = [tuple(c) for c in samp[i][:ns[i]]]
                true_path[i] = [tuple(c) for c in dense[i, :nd[i]].tolist()]
        return sampling_path, true_path

    def _dense_for(self, pose, goal):
        """Dense path of ONE (pose, goal) pair on demand (LazyPath.factory): own small buffers, generator state untouched."""
        import ctypes
        from . import _lib as L
        max_dense = int((self.hl + 8.0 * math.pi * self.tr) / (self.ss / 10)) + 16
        max_samp = max_dense // 10 + 2
        samp, dense = np.empty((1, max_samp, 3)), np.empty((1, max_dense, 3))
        ns, nd = np.zeros(1, dtype=np.int32), np.zeros(1, dtype=np.int32)
        pose_a, goal_a = np.asarray(pose, dtype=np.float64).copy(), np.asarray(goal, dtype=np.float64).reshape(1, 3).copy()
        ext = np.asarray(self.extent, dtype=np.float64).copy()
        vp = ctypes.c_void_p
        boxes, n_boxes = self._obstacle_boxes()
        L.check(L.lib.ipp_dubins_path_set(pose_a.ctypes.data_as(vp), goal_a.ctypes.data_as(vp), 1, float(self.tr),
                                          float(self.ss / 10), 10, ext.ctypes.data_as(vp), float(3 * self.tr),
                                          boxes, n_boxes, samp.ctypes.data_as(vp), ns.ctypes.data_as(vp), max_samp,
                                          dense.ctypes.data_as(vp), nd.ctypes.data_as(vp), max_dense))
        return [tuple(c) for c in dense[0, :int(nd[0])].tolist()]

    def _buffered_paths_python(self):
        sampling_path, true_path = {}, {}
        ext, tr = self.extent, self.tr
        free = isinstance(self.obstacle_world, FreeWorld)
        for i, goal in enumerate(self.goals):
            word, params = dubins_shortest_path(self.cp, goal, tr)
            fconfig = dubins_sample_many(self.cp, word, params, tr, self.ss / 10)
            inside = (fconfig[:, 0] > ext[0]) & (fconfig[:, 0] < ext[1]) & (fconfig[:, 1] > ext[2]) & (fconfig[:, 1] < ext[3])
            if not free:
                inside &= ~np.array([self.obstacle_world.in_obstacle((c[0], c[1]), buff=0.0) for c in fconfig], dtype=bool)
            stop = int(np.argmin(inside)) if not inside.all() else fconfig.shape[0]      # first violation ends the path
            ftemp = [tuple(c) for c in fconfig[:stop].tolist()]
            try:
                ttemp = ftemp[0::10]
                for m, c in enumerate(ttemp):
                    if (c[0] <= ext[0] + 3 * tr or c[0] >= ext[1] - 3 * tr or c[1] <= ext[2] + 3 * tr or
                            c[1] >= ext[3] - 3 * tr or self.obstacle_world.in_obstacle((c[0], c[1]), buff=3 * tr)):
                        ttemp = ttemp[0:m - 1]
                if len(ttemp) >= 2:
                    sampling_path[i] = ttemp
                    true_path[i] = ftemp[0:ftemp.index(ttemp[-1]) + 1]
            except Exception:
                pass
        return sampling_path, true_path

    def make_sample_paths(self):
        coords, true_coords = self.buffered_paths()
        self.samples = coords
        return coords, true_coords


def dubins_paths_device(poses, goals, turning_radius, sample_step, extent, max_samples=32, obstacle_world=None):
    """Batched candidate-path generation on the GPU (BASELINE config 5): one Dubins path per (pose, goal) row with the
    reference generator's sampling (dense step ss/10, every 10th kept) and border / obstacle trimming
    (paths_library.py:147-175); obstacle_world: None or a world with boxes() (obstacles.py, at most 16 boxes).
    poses, goals: (P, 3) arrays or CUDA tensors.  Returns CUDA tensors (samples (P, max_samples, 3), counts (P,) int32,
    end_poses (P, 3)); counts[p] == 0 means the path was dropped (fewer than 2 samples survive)."""
    import ctypes
    import torch
    from . import _lib as L
    poses_d, goals_d = L.to_device(poses), L.to_device(goals)
    P = poses_d.shape[0]
    samp = torch.empty((P, max_samples, 3), dtype=torch.float64, device=poses_d.device)
    counts = torch.empty(P, dtype=torch.int32, device=poses_d.device)
    end = torch.empty((P, 3), dtype=torch.float64, device=poses_d.device)
    ext = (ctypes.c_double * 4)(*[float(e) for e in extent])
    if obstacle_world is not None and not callable(getattr(obstacle_world, 'boxes', None)):
        raise TypeError('dubins_paths_device needs an obstacle world with boxes() (informative_path_planning_b200.obstacles)')
    boxes, n_boxes = box_array(obstacle_world.boxes()) if obstacle_world is not None else (None, 0)
    L.check(L.lib.ipp_dubins_paths_device(L.ptr(poses_d), L.ptr(goals_d), P, float(turning_radius), float(sample_step) / 10,
                                          10, ext, float(3 * turning_radius), boxes, n_boxes, L.ptr(samp), L.ptr(counts), int(max_samples),
                                          L.ptr(end), L.stream_ptr()))
    return samp, counts, end


def ragged_points(samp, counts):
    """(samples, counts) of dubins_paths_device -> (points (M, 2) of the kept paths, offsets (P_kept + 1,) int64,
    indices of the kept paths): the layout reward_paths_device takes.  torch indexing = plumbing, no numerics."""
    import torch
    keep = counts > 0
    c = counts[keep].to(torch.int64)
    mask = torch.arange(samp.shape[1], device=samp.device)[None, :] < c[:, None]
    pts = samp[keep][mask][:, :2].contiguous()
    offsets = torch.zeros(c.numel() + 1, dtype=torch.int64, device=samp.device)
    offsets[1:] = torch.cumsum(c, 0)
    return pts, offsets, torch.nonzero(keep).reshape(-1)
