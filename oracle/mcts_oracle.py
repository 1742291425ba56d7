"""CPU restatement of the tree search that calls the hot path.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/mcts_library.py: Node :277-312, Tree :314-488 (DPW), BeliefTree :491-632,
cMCTS :636-713; and the straight-line path generator /root/reference/paths_library.py:18-117
(the one generator that needs no `dubins` C library).  Global ``np.random`` / ``random`` are
consumed in the reference's order so seeded runs select the same actions.
"""
import copy
import random
import time

import numpy as np

from . import aq_oracle as aq


def _py2_round(x):
    """int(round(x)) under the reference's Python 2 (halves away from zero; Python 3 rounds halves to even)."""
    x = float(x)
    f = np.floor(abs(x))
    r = f + 1.0 if abs(x) - f >= 0.5 else f
    return int(-r if x < 0 else r)


class PathGeneratorOracle(object):
    """paths_library.py:18-117 (frontier fan of straight segments)."""

    def __init__(self, frontier_size, horizon_length, turning_radius, sample_step, extent):
        self.fs, self.hl, self.tr, self.ss, self.extent = frontier_size, horizon_length, turning_radius, sample_step, extent
        self.goals = []
        self.samples = {}
        self.cp = (0, 0, 0)

    def generate_frontier_points(self):                      # :44-83
        angle = np.linspace(-2.35, 2.35, self.fs)
        goals = []
        for a in angle:
            x = self.hl * np.cos(self.cp[2] + a) + self.cp[0]
            y = self.hl * np.sin(self.cp[2] + a) + self.cp[1]
            p = self.cp[2] + a
            if np.linalg.norm([self.cp[0] - x, self.cp[1] - y]) <= self.tr:
                pass
            elif x > self.extent[1] - 3 * self.tr or x < self.extent[0] + 3 * self.tr:
                pass
            elif y > self.extent[3] - 3 * self.tr or y < self.extent[2] + 3 * self.tr:
                pass
            else:
                goals.append((x, y, p))
        goals.append(self.cp)
        self.goals = goals
        return goals

    def make_sample_paths(self):                             # :85-105
        cp = np.array(self.cp)
        coords = {}
        for i, goal in enumerate(self.goals):
            g = np.array(goal)
            distance = np.sqrt((cp[0] - g[0]) ** 2 + (cp[1] - g[1]) ** 2)
            samples = _py2_round(distance / self.ss)
            for j in range(0, samples):
                x = cp[0] + ((j + 1) * self.ss) * np.cos(g[2])
                y = cp[1] + ((j + 1) * self.ss) * np.sin(g[2])
                coords.setdefault(i, []).append((x, y, g[2]))
        self.samples = coords
        return coords, coords

    def get_path_set(self, current_pose):                    # :107-117
        self.cp = current_pose
        self.generate_frontier_points()
        return self.make_sample_paths()


class _FreeWorldOracle(object):                             # obstacles.py FreeWorld: nothing is an obstacle
    def in_obstacle(self, point, buff=0.0):
        return False


class DubinsPathGeneratorOracle(PathGeneratorOracle):
    """paths_library.py:141-189 over oracle/dubins_oracle.py (the restated `dubins` library): dense configurations every
    ss / 10 cut at the first one outside the open extent or inside an obstacle, every 10th kept, then the reference's
    border / obstacle trimming loop with its `ttemp[0:m-1]` slice (negative-index quirk included) and its try/except."""

    def __init__(self, frontier_size, horizon_length, turning_radius, sample_step, extent, obstacle_world=None):
        PathGeneratorOracle.__init__(self, frontier_size, horizon_length, turning_radius, sample_step, extent)
        self.obstacle_world = _FreeWorldOracle() if obstacle_world is None else obstacle_world

    def buffered_paths(self):                                # :147-175
        from . import dubins_oracle
        sampling_path, true_path = {}, {}
        for i, goal in enumerate(self.goals):
            word, params = dubins_oracle.shortest_path(self.cp, goal, self.tr)
            fconfig, _ = dubins_oracle.sample_many(self.cp, word, params, self.tr, self.ss / 10)
            ftemp = []
            for c in fconfig:
                if c[0] > self.extent[0] and c[0] < self.extent[1] and c[1] > self.extent[2] and c[1] < self.extent[3] \
                        and not self.obstacle_world.in_obstacle((c[0], c[1]), buff=0.0):
                    ftemp.append(c)
                else:
                    break
            try:
                ttemp = ftemp[0::10]
                for m, c in enumerate(ttemp):
                    if c[0] <= self.extent[0] + 3 * self.tr or c[0] >= self.extent[1] - 3 * self.tr or \
                            c[1] <= self.extent[2] + 3 * self.tr or c[1] >= self.extent[3] - 3 * self.tr or \
                            self.obstacle_world.in_obstacle((c[0], c[1]), buff=3 * self.tr):
                        ttemp = ttemp[0:m - 1]
                if len(ttemp) >= 2:
                    sampling_path[i] = ttemp
                    true_path[i] = ftemp[0:ftemp.index(ttemp[-1]) + 1]
            except Exception:
                pass
        return sampling_path, true_path

    def make_sample_paths(self):                             # :178-189
        coords, true_coords = self.buffered_paths()
        self.samples = coords
        return coords, true_coords


class Node(object):                                          # :277-312
    def __init__(self, pose, parent, name, action=None, dense_path=None, zvals=None):
        self.pose, self.name, self.zvals = pose, name, zvals
        self.reward = 0.0
        self.nqueries = 0
        self.parent = parent
        self.children = None
        if action is None:
            self.node_type = 'B'
            self.action = None
            self.dense_path = None
            self.depth = 0 if parent is None else parent.depth + 1
        else:
            self.node_type = 'BA'
            self.action = action
            self.dense_path = dense_path
            self.depth = parent.depth

    def add_children(self, child):
        if self.children is None:
            self.children = []
        self.children.append(child)


class Tree(object):                                          # :314-488
    def __init__(self, f_rew, f_aqu, belief, pose, path_generator, t, depth, param, c):
        self.path_generator, self.max_depth, self.param, self.t = path_generator, depth, param, t
        self.f_rew, self.aquisition_function, self.c = f_rew, f_aqu, c
        self.root = Node(pose, parent=None, name='root')

    def backprop(self, node, reward):                        # :330-345
        while node is not None:
            node.nqueries += 1
            node.reward += reward
            node = node.parent

    def get_next_leaf(self, belief):                         # :347-352
        leaf, reward = self.leaf_helper(self.root, 0.0, belief)
        self.backprop(leaf, reward)

    def _reward(self, xobs, belief):                         # :386-396
        if self.f_rew in ('mes', 'maxs-mes', 'exp_improve', 'naive', 'naive_value'):
            return self.aquisition_function(time=self.t, xvals=xobs, robot_model=belief, param=self.param)
        return self.aquisition_function(time=self.t, xvals=xobs, robot_model=belief)

    def _simulate_obs(self, xobs, belief):                   # :418-424 (belief.model is always None: Q5)
        n_points = xobs.shape[0]
        if belief.model is None:
            zobs = np.random.multivariate_normal(mean=np.zeros((n_points,)), cov=np.eye(n_points) * belief.variance)
            return np.reshape(zobs, (n_points, 1))
        return belief.posterior_samples(xobs, full_cov=False, size=1)

    def leaf_helper(self, node, reward, belief):             # :354-442
        if node.node_type == 'B':
            if node.depth == self.max_depth:
                return node, reward
            if node.children is None:
                self.build_action_children(node)
            if node.children is None:
                return node, reward
            return self.leaf_helper(self.get_next_child(node), reward, belief)
        obs = np.array(node.action)
        xobs = np.vstack([obs[:, 0], obs[:, 1]]).T
        r = self._reward(xobs, belief)
        if node.children is not None:                        # :398-415 progressive widening
            alpha = 3.0 / (10.0 * (self.max_depth - node.depth) - 3.0)
            nchild = len(node.children)
            if node.depth < self.max_depth - 1 and np.floor(nchild ** alpha) == np.floor((nchild - 1) ** alpha):
                child = random.choice(node.children)         # :408 (draw consumed, result discarded)
                nq = [n.nqueries for n in node.children]
                child = random.choice([n for n in node.children if n.nqueries == min(nq)])
                belief.add_data(xobs, child.zvals)
                return self.leaf_helper(child, reward + r, belief)
        zobs = self._simulate_obs(xobs, belief)
        belief.add_data(xobs, zobs)                          # :426
        belief.add_data(xobs, zobs)                          # :430 (Q4: added twice)
        child = Node(pose=node.dense_path[-1], parent=node, name=node.name + '_belief' + str(node.depth + 1), zvals=zobs)
        node.add_children(child)
        return self.leaf_helper(child, reward + r, belief)

    def get_next_child(self, node):                          # :444-455
        vals = {}
        e_d = 0.5 * (1.0 - (3.0 / (10.0 * (self.max_depth - node.depth))))
        for child in node.children:
            if child.nqueries == 0:
                return child
            vals[child] = child.reward / float(child.nqueries) + self.c * np.sqrt(
                (float(node.nqueries) ** e_d) / float(child.nqueries))
        best = max(vals.values())
        return random.choice([k for k in vals.keys() if vals[k] == best])

    def build_action_children(self, parent):                 # :457-471
        actions, dense_paths = self.path_generator.get_path_set(parent.pose)
        if len(actions) == 0:
            return
        for i, action in enumerate(list(actions.keys())):
            parent.add_children(Node(pose=parent.pose, parent=parent, name=parent.name + '_action' + str(i),
                                     action=actions[action], dense_path=dense_paths[action]))


class BeliefTree(Tree):                                      # :491-632
    def _mle_obs(self, xobs, belief):                        # :598-603
        if belief.model is None:
            return np.zeros((xobs.shape[0], 1))
        return belief.predict_value(xobs)[0]

    def random_rollouts(self, node, reward, belief):         # :499-543
        cur_depth = node.depth
        pose = node.pose
        while cur_depth <= self.max_depth:
            actions, dense_paths = self.path_generator.get_path_set(pose)
            keys = list(actions.keys())
            if len(actions) <= 1:
                return reward
            a = np.random.randint(0, len(actions) - 1)       # Q12: never the last action
            obs = np.array(actions[keys[a]])
            xobs = np.vstack([obs[:, 0], obs[:, 1]]).T
            r = self._reward(xobs, belief)
            zobs = self._mle_obs(xobs, belief)
            belief.add_data(xobs, zobs)
            belief.add_data(xobs, zobs)
            pose = dense_paths[keys[a]][-1]
            reward += r
            cur_depth += 1
        return reward

    def leaf_helper(self, node, reward, belief):             # :545-621
        if node.node_type == 'B':
            if node.depth == self.max_depth:
                return node, reward
            if node.children is None:
                self.build_action_children(node)
            if node.children is None:
                return node, reward
            child, full = self.get_next_child(node)
            if full:
                return self.leaf_helper(child, reward, belief)
            return child, self.random_rollouts(node, reward, belief)
        obs = np.array(node.action)
        xobs = np.vstack([obs[:, 0], obs[:, 1]]).T
        r = self._reward(xobs, belief)
        zobs = self._mle_obs(xobs, belief)
        belief.add_data(xobs, zobs)
        belief.add_data(xobs, zobs)
        child = Node(pose=node.dense_path[-1], parent=node, name=node.name + '_belief' + str(node.depth + 1), zvals=zobs)
        node.add_children(child)
        return self.leaf_helper(child, reward + r, belief)

    def get_next_child(self, node):                          # :624-632
        vals = {}
        for child in node.children:
            if child.nqueries == 0:
                return child, False
            vals[child] = child.reward / float(child.nqueries) + self.c * np.sqrt(
                2.0 * np.log(float(node.nqueries)) / float(child.nqueries))
        best = max(vals.values())
        return random.choice([k for k in vals.keys() if vals[k] == best]), True


def uct_constant(f_rew, tree_type):
    """mcts_library.py:645-660."""
    if f_rew == 'mean':
        return 1000 if tree_type == 'belief' else 5000
    if f_rew == 'exp_improve':
        return 200
    if f_rew == 'mes':
        return 1.0 / np.sqrt(2.0) if tree_type == 'belief' else 1.0
    return 1.0


class cMCTSOracle(object):                                   # :636-713
    def __init__(self, computation_budget, belief, initial_pose, rollout_length, path_generator,
                 aquisition_function, f_rew, T, aq_param=None, use_cost=False, tree_type='dpw',
                 sample_max_vals=aq.sample_max_vals):
        self.comp_budget, self.GP, self.cp, self.rl = computation_budget, belief, initial_pose, rollout_length
        self.path_generator, self.aquisition_function, self.f_rew = path_generator, aquisition_function, f_rew
        self.current_max = aq_param
        self.aq_param = aq_param
        self.tree_type = tree_type
        self.c = uct_constant(f_rew, tree_type)
        self.max_val = self.max_locs = self.target = None
        self._sample_max_vals = sample_max_vals

    def choose_trajectory(self, t):
        if self.f_rew == 'mes':
            self.max_val, self.max_locs, self.target = self._sample_max_vals(self.GP, t=t)
            param = (self.max_val, self.max_locs, self.target)
        elif self.f_rew == 'exp_improve':
            param = [self.current_max]
        elif self.f_rew in ('naive', 'naive_value'):
            self.max_val, self.max_locs, self.target = self._sample_max_vals(self.GP, t=t, nK=int(self.aq_param[0]),
                                                                             f_rew=self.f_rew)
            param = ((self.max_val, self.max_locs, self.target), self.aq_param[1])
        else:
            param = None
        cls = {'dpw': Tree, 'belief': BeliefTree}.get(self.tree_type)
        if cls is None:
            raise ValueError('Tree type must be one of either \'dpw\' or \'belief\'')
        self.tree = cls(self.f_rew, self.aquisition_function, self.GP, self.cp, self.path_generator, t,
                        depth=self.rl, param=param, c=self.c)
        time_start = time.time()                             # :689: the reference's "Rollouts completed in ...s"
        for _ in range(self.comp_budget):                    # :692-698
            self.tree.get_next_leaf(copy.copy(self.GP))
        self.rollout_seconds = time.time() - time_start
        top = max(n.nqueries for n in self.tree.root.children)
        best = random.choice([n for n in self.tree.root.children if n.nqueries == top])     # :707
        all_vals = {i: c.reward / float(c.nqueries) for i, c in enumerate(self.tree.root.children)}
        paths, _ = self.path_generator.get_path_set(self.cp)
        return (best.action, best.dense_path, best.reward / float(best.nqueries), paths, all_vals,
                self.max_locs, self.max_val, self.target)
