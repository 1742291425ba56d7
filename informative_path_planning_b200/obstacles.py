"""Drop-in for the reference's obstacles.py (FreeWorld :16-24, BlockWorld :26-102, BugTrap :104-163,
ChannelWorld :165-193) without shapely / matplotlib.

Every world of the reference is a union of axis-aligned boxes, some with a box-shaped hole, so next to the reference's
own `in_obstacle(point, buff)` predicate each world exposes `boxes()`: the ipp_obstacle_box list (include/ipp_b200.h)
that lets the native and device Dubins generators (csrc/paths.cu) trim candidate paths against it with the same
comparisons.  `get_obstacles()` returns the corner lists (the reference returns shapely polygons of the same corners,
used for plotting only).
"""
import ctypes

import numpy as np

_INF = float('inf')


class IppObstacleBox(ctypes.Structure):
    _fields_ = [('outer', ctypes.c_double * 4), ('hole', ctypes.c_double * 4), ('hole_buff', ctypes.c_double * 4),
                ('has_hole', ctypes.c_int32), ('reserved', ctypes.c_int32)]


def _box(outer, hole=None, hole_buff=(0., 0., 0., 0.)):
    b = IppObstacleBox()
    b.outer[:] = [float(v) for v in outer]
    b.has_hole = 0 if hole is None else 1
    b.hole[:] = [0.0] * 4 if hole is None else [float(v) for v in hole]
    b.hole_buff[:] = [float(v) for v in hole_buff]
    return b


def box_array(boxes):
    """list of IppObstacleBox -> (ctypes array or None, count) for the C ABI."""
    if not boxes:
        return None, 0
    arr = (IppObstacleBox * len(boxes))(*boxes)
    return arr, len(boxes)


def in_boxes(boxes, point, buff):
    """The box-list predicate of include/ipp_b200.h in Python (what csrc/paths.cu evaluates)."""
    x, y = point[0], point[1]
    for b in boxes:
        if x > b.outer[0] - buff and x < b.outer[1] + buff and y > b.outer[2] - buff and y < b.outer[3] + buff:
            if (b.has_hole and x >= b.hole[0] + b.hole_buff[0] * buff and x <= b.hole[1] + b.hole_buff[1] * buff and
                    y >= b.hole[2] + b.hole_buff[2] * buff and y <= b.hole[3] + b.hole_buff[3] * buff):
                continue
            return True
    return False


class FreeWorld(object):
    def __init__(self):
        self.obstacles = []

    def in_obstacle(self, point, buff=0.1):
        return False

    def get_obstacles(self):
        return []

    def boxes(self):
        return []


class BlockWorld(object):
    """num_blocks rectangles of size dim_blocks at `centers` (random inside the extent when None; the draws use the
    global np.random in the reference's order: x then y per block)."""

    def __init__(self, extent, num_blocks=1, dim_blocks=(2., 2.), centers=None):
        self.extent = extent
        self.num_blocks = num_blocks
        self.dim_blocks = dim_blocks
        if centers is None:
            self.centers = []
            for _ in range(num_blocks):
                self.centers.append(((extent[1] - extent[0]) * np.random.random_sample() + extent[0],
                                     (extent[3] - extent[2]) * np.random.random_sample() + extent[2]))
        else:
            self.centers = centers
        self.points = []
        for c in self.centers:
            self.points.append([(c[0] + dim_blocks[0] / 2, c[1] + dim_blocks[1] / 2),
                                (c[0] + dim_blocks[0] / 2, c[1] - dim_blocks[1] / 2),
                                (c[0] - dim_blocks[0] / 2, c[1] - dim_blocks[1] / 2),
                                (c[0] - dim_blocks[0] / 2, c[1] + dim_blocks[1] / 2)])
        self.obstacles = [list(p) for p in self.points]

    def in_obstacle(self, point, buff=0.1):
        for obs in self.points:
            if point[0] > obs[2][0] - buff and point[0] < obs[0][0] + buff:
                if point[1] > obs[1][1] - buff and point[1] < obs[0][1] + buff:
                    return True
        return False

    def boxes(self):
        return [_box((o[2][0], o[0][0], o[1][1], o[0][1])) for o in self.points]

    def get_obstacles(self):
        return self.obstacles

    def get_centers(self):
        return self.centers

    def get_coordinates(self):
        return [([p[0] for p in o] + [o[0][0]], [p[1] for p in o] + [o[0][1]]) for o in self.obstacles]


class BugTrap(BlockWorld):
    def __init__(self, extent, opening_location, opening_size, channel_size=0.5, width=3., orientation='left'):
        self.extent = extent
        self.opening_location = opening_location
        self.opening_size = opening_size
        self.orientation = orientation
        ol, os_, cs = opening_location, opening_size, channel_size
        if orientation not in ('left', 'right'):
            raise ValueError("orientation must be 'left' or 'right'")
        sgn = 1.0 if orientation == 'left' else -1.0
        far, wall = ol[0] + sgn * (width + cs), ol[0] + sgn * width
        points = [(ol[0], ol[1] + os_ / 2), (ol[0], ol[1] + os_ / 2 + cs), (far, ol[1] + os_ / 2 + cs),
                  (far, ol[1] - os_ / 2 - cs), (ol[0], ol[1] - os_ / 2 - cs), (ol[0], ol[1] - os_ / 2),
                  (wall, ol[1] - os_ / 2), (wall, ol[1] + os_ / 2), (ol[0], ol[1] + os_ / 2)]
        if orientation == 'left':
            self.inner_xlim = [ol[0], ol[0] + width]
            self.outer_xlim = [ol[0], ol[0] + width + cs]
        else:
            self.inner_xlim = [ol[0] - width, ol[0]]
            self.outer_xlim = [ol[0] - width - cs, ol[0]]
        self.inner_ylim = [ol[1] + os_ / 2, ol[1] - os_ / 2]
        self.outer_ylim = [ol[1] + os_ / 2 + cs, ol[1] - os_ / 2 - cs]
        self.obstacles = [points]

    def in_obstacle(self, point, buff=0.1):
        if point[0] > self.outer_xlim[0] - buff and point[0] < self.outer_xlim[1] + buff:
            if point[1] < self.outer_ylim[0] + buff and point[1] > self.outer_ylim[1] - buff:
                if self.orientation == 'left':
                    in_x = point[0] >= self.inner_xlim[0] - buff and point[0] <= self.inner_xlim[1] - buff
                else:
                    in_x = point[0] >= self.inner_xlim[0] + buff and point[0] <= self.inner_xlim[1] + buff
                if in_x and point[1] <= self.inner_ylim[0] - buff and point[1] >= self.inner_ylim[1] + buff:
                    return False
                return True
        return False

    def boxes(self):
        sx = -1.0 if self.orientation == 'left' else 1.0
        return [_box((self.outer_xlim[0], self.outer_xlim[1], self.outer_ylim[1], self.outer_ylim[0]),
                     hole=(self.inner_xlim[0], self.inner_xlim[1], self.inner_ylim[1], self.inner_ylim[0]),
                     hole_buff=(sx, sx, 1.0, -1.0))]


class ChannelWorld(BlockWorld):
    def __init__(self, extent, opening_location, opening_size, wall_thickness):
        self.extent = extent
        self.opening_location = opening_location
        self.opening_size = opening_size
        self.wall_thickness = wall_thickness
        ol, os_, wt = opening_location, opening_size, wall_thickness
        top_wall = [(ol[0] - wt / 2, ol[1] + os_ / 2), (ol[0] - wt / 2, extent[3]),
                    (ol[0] + wt / 2, extent[3]), (ol[0] + wt / 2, ol[1] + os_ / 2)]
        bottom_wall = [(ol[0] - wt / 2, ol[1] - os_ / 2), (ol[0] - wt / 2, extent[2]),
                       (ol[0] + wt / 2, extent[2]), (ol[0] + wt / 2, ol[1] - os_ / 2)]
        self.obstacles = [top_wall, bottom_wall]
        self.xlim = [ol[0] - wt / 2, ol[0] + wt / 2]
        self.ylim = [ol[1] - os_ / 2, ol[1] + os_ / 2]

    def in_obstacle(self, point, buff=0.1):
        if point[0] > self.xlim[0] - buff and point[0] < self.xlim[1] + buff:
            if point[1] < self.ylim[0] + buff or point[1] > self.ylim[1] - buff:
                return True
        return False

    def boxes(self):
        # the wall strip is unbounded in y (the reference never tests the extent here); the opening is its hole
        return [_box((self.xlim[0], self.xlim[1], -_INF, _INF),
                     hole=(-_INF, _INF, self.ylim[0], self.ylim[1]), hole_buff=(0.0, 0.0, 1.0, -1.0))]
