"""Candidate-path generators (reference paths_library.py), restated for Python 3 and without the
un-vendored `dubins` C library.

Path_Generator         reference :18-139   frontier fan of straight segments
Dubins_Path_Generator  reference :141-189  shortest Dubins curves to the frontier goals, sampled densely
                                           (step ss/10), every 10th point kept, trimmed at the border

The Dubins solver below is the classical six-word closed form (LSL, LSR, RSL, RSR, RLR, LRL; Shkel & Lumelsky
2001), the algorithm the `dubins` package wraps; shortest word wins.  Obstacle worlds are the reference's
FreeWorld contract: any object with in_obstacle(point, buff) -> bool.
"""
import math

import numpy as np


class FreeWorld(object):
    """reference obstacles.py:16-24"""

    def __init__(self):
        self.obstacles = []

    def in_obstacle(self, point, buff=0.1):
        return False

    def get_obstacles(self):
        return []


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
        """reference :44-83"""
        goals = []
        for a in np.linspace(-2.35, 2.35, self.fs):
            x = self.hl * np.cos(self.cp[2] + a) + self.cp[0]
            y = self.hl * np.sin(self.cp[2] + a) + self.cp[1]
            p = self.cp[2] + a
            if np.linalg.norm([self.cp[0] - x, self.cp[1] - y]) <= self.tr:
                continue
            if x > self.extent[1] - 3 * self.tr or x < self.extent[0] + 3 * self.tr:
                continue
            if y > self.extent[3] - 3 * self.tr or y < self.extent[2] + 3 * self.tr:
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
            for j in range(int(round(distance / self.ss))):
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
    sa, sb, ca, cb, cab = math.sin(alpha), math.sin(beta), math.cos(alpha), math.cos(beta), math.cos(alpha - beta)
    out = []
    # LSL
    p2 = 2 + d * d - 2 * cab + 2 * d * (sa - sb)
    if p2 >= 0:
        tmp = math.atan2(cb - ca, d + sa - sb)
        out.append(('LSL', _mod2pi(-alpha + tmp), math.sqrt(p2), _mod2pi(beta - tmp)))
    # RSR
    p2 = 2 + d * d - 2 * cab + 2 * d * (sb - sa)
    if p2 >= 0:
        tmp = math.atan2(ca - cb, d - sa + sb)
        out.append(('RSR', _mod2pi(alpha - tmp), math.sqrt(p2), _mod2pi(-beta + tmp)))
    # LSR
    p2 = -2 + d * d + 2 * cab + 2 * d * (sa + sb)
    if p2 >= 0:
        p = math.sqrt(p2)
        tmp = math.atan2(-ca - cb, d + sa + sb) - math.atan2(-2.0, p)
        out.append(('LSR', _mod2pi(-alpha + tmp), p, _mod2pi(-_mod2pi(beta) + tmp)))
    # RSL
    p2 = d * d - 2 + 2 * cab - 2 * d * (sa + sb)
    if p2 >= 0:
        p = math.sqrt(p2)
        tmp = math.atan2(ca + cb, d - sa - sb) - math.atan2(2.0, p)
        out.append(('RSL', _mod2pi(alpha - tmp), p, _mod2pi(beta - tmp)))
    # RLR
    tmp = (6.0 - d * d + 2 * cab + 2 * d * (sa - sb)) / 8.0
    if abs(tmp) <= 1:
        p = _mod2pi(2 * math.pi - math.acos(tmp))
        t = _mod2pi(alpha - math.atan2(ca - cb, d - sa + sb) + p / 2.0)
        out.append(('RLR', t, p, _mod2pi(alpha - beta - t + p)))
    # LRL
    tmp = (6.0 - d * d + 2 * cab + 2 * d * (sb - sa)) / 8.0
    if abs(tmp) <= 1:
        p = _mod2pi(2 * math.pi - math.acos(tmp))
        t = _mod2pi(-alpha - math.atan2(ca - cb, d + sa - sb) + p / 2.0)
        out.append(('LRL', t, p, _mod2pi(_mod2pi(beta) - alpha - t + p)))
    return out


def dubins_shortest_path(q0, q1, rho):
    """Shortest Dubins word between poses q0 and q1 = (x, y, heading); returns (word, (t, p, q)) in units of rho."""
    dx, dy = q1[0] - q0[0], q1[1] - q0[1]
    D = math.hypot(dx, dy)
    d = D / rho
    theta = _mod2pi(math.atan2(dy, dx)) if D > 0 else 0.0
    alpha, beta = _mod2pi(q0[2] - theta), _mod2pi(q1[2] - theta)
    best = min(_dubins_words(alpha, beta, d), key=lambda w: w[1] + w[2] + w[3])
    return best[0], best[1:]


def _segment(q, length, kind):
    x, y, th = q
    if kind == 'L':
        return (x + math.sin(th + length) - math.sin(th), y - math.cos(th + length) + math.cos(th), th + length)
    if kind == 'R':
        return (x - math.sin(th - length) + math.sin(th), y + math.cos(th - length) - math.cos(th), th - length)
    return (x + math.cos(th) * length, y + math.sin(th) * length, th)


def dubins_sample_many(q0, word, params, rho, step):
    """Configurations every `step` along the path (the `dubins` package's sample_many): [(x, y, heading)]."""
    total = sum(params) * rho
    out = []
    s = 0.0
    while s < total:
        tprime = s / rho
        q = (0.0, 0.0, q0[2])
        rem = tprime
        for seg_len, kind in zip(params, word):
            use = min(rem, seg_len)
            q = _segment(q, use, kind)
            rem -= use
            if rem <= 0:
                break
        out.append((q[0] * rho + q0[0], q[1] * rho + q0[1], _mod2pi(q[2])))
        s += step
    return out


class Dubins_Path_Generator(Path_Generator):
    """reference :141-189"""

    def buffered_paths(self):
        sampling_path, true_path = {}, {}
        ext, tr = self.extent, self.tr
        for i, goal in enumerate(self.goals):
            word, params = dubins_shortest_path(self.cp, goal, tr)
            fconfig = dubins_sample_many(self.cp, word, params, tr, self.ss / 10)
            ftemp = []
            for c in fconfig:
                if ext[0] < c[0] < ext[1] and ext[2] < c[1] < ext[3] and not self.obstacle_world.in_obstacle((c[0], c[1]), buff=0.0):
                    ftemp.append(c)
                else:
                    break
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
