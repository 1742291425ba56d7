"""Tree search front-end (reference mcts_library.py:277-713) with the rewards on the GPU.

The tree, its selection rules and its random-number consumption follow the reference exactly
(SURVEY.md 8a row a12), so that on fixed seeds the same actions are chosen:
  * DPW tree (`Tree`): widening exponent alpha = 3 / (10 (D - d) - 3); an existing child is re-used
    iff d < D - 1 and floor(n^alpha) == floor((n - 1)^alpha), picked uniformly among the least visited
    (after one discarded random.choice draw, reference :408); otherwise observations are simulated
    from the PRIOR N(0, variance I) because belief.model is None for OnlineGPModel (:418-422) and are
    appended TWICE (:426, :430);
  * action choice Q/n + c sqrt(N^e_d / n), e_d = (1 - 3 / (10 (D - d))) / 2, first unvisited child wins,
    exact ties broken with random.choice (:444-455);
  * `BeliefTree`: Q/n + c sqrt(2 ln N / n), zero ("maximum likelihood") observations, random rollouts
    below a not-yet-expanded node with np.random.randint(0, len - 1) (never the last action, :510);
  * final action = most visited root child, ties -> random.choice (:707).

Each rollout works on copy.copy(belief): an O(1) fork, because the device state of a belief is
immutable and every append writes new buffers.

Rollout batching (the north-star seam "hand each rollout's batched query points to the GPU in one call"):
within a rollout the tree walk never looks at a reward -- child selection uses statistics of EARLIER
rollouts, paths depend on poses only, and the simulated observations are prior draws / zeros because
belief.model is None (:418-422, :598-601).  So the walk first records (sample points, observation,
1-or-2 appends) per level, consuming the random streams in the reference's order, and then ONE call
(GPModel.rollout_rewards -> ipp_rollout_rewards) returns every level's reward under the belief that level
would have seen.  The sequential path (one reward + up to two add_data per level) remains for beliefs
with a GPy-style model handle, foreign acquisition callables, empty beliefs and non-PD fallbacks.

Batching seam (`evaluate_root_actions`, `choose_myopic`): all candidate paths of a pose are scored by
ONE call of aq_library.reward_paths instead of one predict per path (robot_library.py:175-203).
"""
import copy
import random
import time

import numpy as np
import torch

from . import aq_library as aqlib
from .gpmodel_library import GPModel

_PARAM_REWARDS = ('mes', 'maxs-mes', 'exp_improve', 'naive', 'naive_value')


class Node(object):
    def __init__(self, pose, parent, name, action=None, dense_path=None, zvals=None):
        self.pose = pose
        self.name = name
        self.zvals = zvals
        self.reward = 0.0
        self.nqueries = 0
        self.parent = parent
        self.children = None
        self.action = action
        self.dense_path = dense_path if action is not None else None
        if action is None:
            self.node_type = 'B'
            self.depth = 0 if parent is None else parent.depth + 1
        else:
            self.node_type = 'BA'
            self.depth = parent.depth

    def add_children(self, child_node):
        if self.children is None:
            self.children = []
        self.children.append(child_node)

    def print_self(self):
        print(self.name)


_prior_transforms = {}


def _prior_draw(k, variance):
    """np.random.multivariate_normal(zeros(k), eye(k) * variance) (reference mcts_library.py:418-422), bit for bit
    and consuming the same global random stream, without the per-call SVD + PSD check: numpy's legacy sampler is
    standard_normal(k) @ (sqrt(s)[:, None] * v) with (u, s, v) = svd(cov), which for a fixed (k, variance) is one
    constant k x k matrix."""
    key = (int(k), float(variance))
    tr = _prior_transforms.get(key)
    if tr is None:
        _, s, v = np.linalg.svd(np.eye(k) * float(variance))
        tr = _prior_transforms[key] = np.sqrt(s)[:, None] * v
    x = np.random.standard_normal((k,)).reshape(-1, k)
    x = np.dot(x, tr)
    x += np.zeros((k,))
    return x.reshape(k)


_native_stage = {}
_prior_scales = {}


def _prior_scale_table(variance):
    """[33][32] table d with _prior_draw(k, variance)[i] == standard_normal_i * d[k][i]: the SVD transform of a scaled
    identity is a constant diagonal (checked, else None and the interpreter keeps the search)."""
    key = float(np.asarray(variance, dtype=float).reshape(-1)[0])
    if key not in _prior_scales:
        kmax = aqlib.L.IPP_APPEND_MAX_K
        table = np.zeros((kmax + 1, kmax), dtype=np.float64)
        for k in range(1, kmax + 1):
            _, s, v = np.linalg.svd(np.eye(k) * key)
            tr = np.sqrt(s)[:, None] * v
            if np.count_nonzero(tr - np.diag(np.diag(tr))) != 0:
                table = None
                break
            table[k, :k] = np.diag(tr)
        _prior_scales[key] = table
    return _prior_scales[key]


def _xobs(action):
    obs = np.array(action)
    return np.vstack([obs[:, 0], obs[:, 1]]).T


class Tree(object):
    FUSE_ROLLOUTS = True       # default of .fuse_rollouts: score each rollout with one device call when possible

    def __init__(self, f_rew, f_aqu, belief, pose, path_generator, t, depth, param, c):
        self.path_generator = path_generator
        self.max_depth = depth
        self.param = param
        self.t = t
        self.f_rew = f_rew
        self.aquisition_function = f_aqu
        self.c = c
        self._maxes_d = None
        self.reward_evals = 0                       # one eval = one path's scalar reward (the benchmark's unit)
        self.fuse_rollouts = self.FUSE_ROLLOUTS     # False forces the per-level reward / add_data path
        self.root = Node(pose, parent=None, name='root')

    def get_best_child(self):
        return self.root.children[int(np.argmax([node.nqueries for node in self.root.children]))]

    def backprop(self, leaf_node, reward):
        node = leaf_node
        while node is not None:
            node.nqueries += 1
            node.reward += reward
            node = node.parent

    def get_next_leaf(self, belief):
        leaf, reward = self.leaf_helper(self.root, reward=0.0, belief=belief)
        self.backprop(leaf, reward)

    def action_reward(self, xobs, belief):
        """Reward of one action.  When the acquisition function is one of this package's own callables the
        value stays on the device as a 0-d tensor (no host sync; the tree only needs it at back-propagation,
        reference :352); otherwise the callable is invoked exactly as the reference does (:386-396)."""
        self.reward_evals += 1
        fast = _FAST_REWARDS.get(self.aquisition_function)
        if fast is not None and fast == self.f_rew and isinstance(belief, GPModel):
            param = self.param
            if fast == 'mes' and param is not None and param[0] is not None:
                if self._maxes_d is None:
                    self._maxes_d = aqlib.L.to_device(np.asarray(param[0], dtype=float).reshape(-1))
                param = (self._maxes_d,) + tuple(param[1:])
            r = aqlib.reward_device(fast, self.t, xobs, belief, param)
            if r is not None:
                return r
        if self.f_rew in _PARAM_REWARDS:
            return self.aquisition_function(time=self.t, xvals=xobs, robot_model=belief, param=self.param)
        return self.aquisition_function(time=self.t, xvals=xobs, robot_model=belief)

    # ---------------------------------------------------------------- fused rollouts
    def _fused_kind(self, belief):
        """Reward kind + parameters when the whole rollout can be scored by one device call, else None."""
        if not self.fuse_rollouts:
            return None
        if (_BELIEF_FREE_REWARDS.get(self.aquisition_function) == self.f_rew and isinstance(belief, GPModel)
                and belief.model is None):
            return 'host', None, None        # naive / naive_value never read the belief: no simulated updates needed
        info = _INFO_REWARDS.get(self.aquisition_function)
        if info is not None and info == self.f_rew and isinstance(belief, GPModel):
            if belief.model is not None or belief.xvals is None or belief._winv_d is None:
                return None
            if info == 'info_gain':
                return aqlib.L.REWARD_INFO, None, None
            return 'hotspot', [np.sqrt(aqlib.ucb_beta(self.t))], None        # info_gain + mean_UCB, two device passes
        fast = _FAST_REWARDS.get(self.aquisition_function)
        if fast is None or fast != self.f_rew or not isinstance(belief, GPModel):
            return None
        if belief.model is not None or belief.xvals is None or belief._winv_d is None:
            return None
        if fast == 'mean':
            return aqlib.L.REWARD_UCB, [np.sqrt(aqlib.ucb_beta(self.t))], None
        if fast == 'exp_improve':
            eta = 0.5 if self.param is None else sum(self.param) / len(self.param)
            return aqlib.L.REWARD_EI, [float(np.asarray(eta, dtype=float).reshape(-1)[0])], None
        if self.param is None or self.param[0] is None:
            return None
        if self._maxes_d is None:
            self._maxes_d = aqlib.L.to_device(np.asarray(self.param[0], dtype=float).reshape(-1))
        return aqlib.L.REWARD_MES, None, self._maxes_d

    def _replay(self, levels, pending, belief):
        """The reference's per-level sequence for recorded levels (fallback of the fused path)."""
        for xobs, zobs, reps in levels:
            pending.append(self.action_reward(xobs, belief))
            for _ in range(reps):
                belief.add_data(xobs, zobs)

    def settle_fused(self, reward, levels, belief, fused):
        """Score the recorded levels with one device call; rewards are added level by level like settle()."""
        if not levels:
            return reward
        kind, params, maxes_d = fused
        if kind == 'host':
            for r in aqlib.naive_levels(self.f_rew, self.t, [x for x, _, _ in levels], belief, self.param):
                reward = reward + r
            self.reward_evals += len(levels)
            return reward
        n_pts = sum(x.shape[0] for x, _, _ in levels)
        ok = (len(levels) <= aqlib.L.IPP_ROLLOUT_MAX_LEVELS and n_pts <= aqlib.L.IPP_ROLLOUT_MAX_POINTS
              and max(x.shape[0] for x, _, _ in levels) <= aqlib.L.IPP_APPEND_MAX_K)
        if kind in (aqlib.L.REWARD_INFO, 'hotspot'):
            ok = ok and n_pts <= aqlib.L.IPP_ROLLOUT_INFO_MAX_POINTS
        if ok:
            queries = [(aqlib._queries(self.t, x, belief)[1], z, r) for x, z, r in levels]
            if kind == 'hotspot':                 # hotspot_info_UCB = info_gain + mean_UCB (aq_library.hotspot_info_UCB)
                ig, info = belief.rollout_rewards(aqlib.L.REWARD_INFO, queries)
                ucb, info2 = belief.rollout_rewards(aqlib.L.REWARD_UCB, queries, params=params)
                vals, info = ig + ucb, (info or info2)
            else:
                vals, info = belief.rollout_rewards(kind, queries, params=params, maxes_d=maxes_d)
            ok = info == 0
            if ok:
                self.reward_evals += len(levels)
        if not ok:
            pending = []
            belief.defer_checks(True)
            self._replay(levels, pending, belief)
            return self.settle(reward, pending, belief)
        for r in vals:
            reward = reward + r
        return reward

    @staticmethod
    def settle(reward, pending, belief):
        """One device->host transfer for all the rewards of a rollout; summed in the reference's order
        (reward + r level by level) so the float64 total is bit-identical to per-call accumulation."""
        if pending:
            dev = [r for r in pending if isinstance(r, torch.Tensor)]
            host = iter(torch.stack(dev).cpu().numpy()) if dev else iter(())
            for r in pending:
                reward = reward + (next(host) if isinstance(r, torch.Tensor) else r)
        if isinstance(belief, GPModel):
            belief.check_pending()
        return reward

    def simulate_observation(self, xobs, belief):
        n_points = xobs.shape[0]
        if belief.model is None:
            return np.reshape(_prior_draw(n_points, belief.variance), (n_points, 1))
        return belief.posterior_samples(xobs, full_cov=False, size=1)

    def _expand_action(self, node, pending, belief, levels=None):
        """One 'BA' step: score the action, then extend the rollout's belief.  Returns the next belief node.
        With `levels` (fused rollouts) the step is only recorded: (points, observation, number of appends)."""
        xobs = _xobs(node.action)
        if levels is not None:
            return self._record_action(node, xobs, belief, levels)
        pending.append(self.action_reward(xobs, belief))
        if node.children is not None:
            alpha = 3.0 / (10.0 * (self.max_depth - node.depth) - 3.0)
            nchild = len(node.children)
            if node.depth < self.max_depth - 1 and np.floor(nchild ** alpha) == np.floor((nchild - 1) ** alpha):
                random.choice(node.children)                                   # draw consumed, result unused
                fewest = min(n.nqueries for n in node.children)
                child = random.choice([n for n in node.children if n.nqueries == fewest])
                belief.add_data(xobs, child.zvals)
                return child
        zobs = self.simulate_observation(xobs, belief)
        belief.add_data(xobs, zobs)
        belief.add_data(xobs, zobs)
        child = Node(pose=node.dense_path[-1], parent=node, name=node.name + '_belief' + str(node.depth + 1),
                     zvals=zobs)
        node.add_children(child)
        return child

    def _record_action(self, node, xobs, belief, levels):
        """_expand_action without device work: same branches, same random draws, in the same order."""
        if node.children is not None:
            alpha = 3.0 / (10.0 * (self.max_depth - node.depth) - 3.0)
            nchild = len(node.children)
            if node.depth < self.max_depth - 1 and np.floor(nchild ** alpha) == np.floor((nchild - 1) ** alpha):
                random.choice(node.children)
                fewest = min(n.nqueries for n in node.children)
                child = random.choice([n for n in node.children if n.nqueries == fewest])
                levels.append((xobs, child.zvals, 1))
                return child
        zobs = self.simulate_observation(xobs, belief)
        levels.append((xobs, zobs, 2))
        child = Node(pose=node.dense_path[-1], parent=node, name=node.name + '_belief' + str(node.depth + 1),
                     zvals=zobs)
        node.add_children(child)
        return child

    def leaf_helper(self, current_node, reward, belief):
        node = current_node
        pending = []
        fused = self._fused_kind(belief)
        levels = [] if fused is not None else None
        if isinstance(belief, GPModel):
            belief.defer_checks(True)
        while True:
            if node.node_type == 'B':
                if node.depth == self.max_depth:
                    break
                if node.children is None:
                    self.build_action_children(node)
                if node.children is None:
                    break
                node = self.get_next_child(node)
            else:
                node = self._expand_action(node, pending, belief, levels)
        if fused is not None:
            return node, self.settle_fused(reward, levels, belief, fused)
        return node, self.settle(reward, pending, belief)

    def get_next_child(self, current_node):
        vals = {}
        e_d = 0.5 * (1.0 - (3.0 / (10.0 * (self.max_depth - current_node.depth))))
        for child in current_node.children:
            if child.nqueries == 0:
                return child
            vals[child] = child.reward / float(child.nqueries) + self.c * np.sqrt(
                (float(current_node.nqueries) ** e_d) / float(child.nqueries))
        top = max(vals.values())
        return random.choice([key for key in vals.keys() if vals[key] == top])

    def build_action_children(self, parent):
        lazy = getattr(self.path_generator, 'get_path_set_lazy', None)     # same paths, dense ones materialised on demand
        actions, dense_paths = lazy(parent.pose) if lazy is not None else self.path_generator.get_path_set(parent.pose)
        if len(actions) == 0:
            return
        for i, action in enumerate(list(actions.keys())):
            parent.add_children(Node(pose=parent.pose, parent=parent, name=parent.name + '_action' + str(i),
                                     action=actions[action], dense_path=dense_paths[action]))

    def count_leaves(self, node=None):
        node = self.root if node is None else node
        if node.children is None:
            return 1
        return sum(self.count_leaves(c) for c in node.children)


class BeliefTree(Tree):
    def mle_observation(self, xobs, belief):
        if belief.model is None:
            return np.zeros((xobs.shape[0], 1))
        return belief.predict_value(xobs)[0]

    def random_rollouts(self, current_node, pending, belief, levels=None):
        cur_depth = current_node.depth
        pose = current_node.pose
        while cur_depth <= self.max_depth:
            actions, dense_paths = self.path_generator.get_path_set(pose)
            keys = list(actions.keys())
            if len(actions) <= 1:
                return
            a = np.random.randint(0, len(actions) - 1)
            xobs = _xobs(actions[keys[a]])
            if levels is not None:
                levels.append((xobs, self.mle_observation(xobs, belief), 2))
            else:
                pending.append(self.action_reward(xobs, belief))
                zobs = self.mle_observation(xobs, belief)
                belief.add_data(xobs, zobs)
                belief.add_data(xobs, zobs)
            pose = dense_paths[keys[a]][-1]
            cur_depth += 1

    def leaf_helper(self, current_node, reward, belief):
        node = current_node
        pending = []
        fused = self._fused_kind(belief)
        levels = [] if fused is not None else None
        if isinstance(belief, GPModel):
            belief.defer_checks(True)
        while True:
            if node.node_type == 'B':
                if node.depth == self.max_depth:
                    break
                if node.children is None:
                    self.build_action_children(node)
                if node.children is None:
                    break
                child, full_action_set = self.get_next_child(node)
                if not full_action_set:
                    self.random_rollouts(node, pending, belief, levels)
                    node = child
                    break
                node = child
            else:
                xobs = _xobs(node.action)
                if levels is not None:
                    zobs = self.mle_observation(xobs, belief)
                    levels.append((xobs, zobs, 2))
                else:
                    pending.append(self.action_reward(xobs, belief))
                    zobs = self.mle_observation(xobs, belief)
                    belief.add_data(xobs, zobs)
                    belief.add_data(xobs, zobs)
                child = Node(pose=node.dense_path[-1], parent=node,
                             name=node.name + '_belief' + str(node.depth + 1), zvals=zobs)
                node.add_children(child)
                node = child
        if fused is not None:
            return node, self.settle_fused(reward, levels, belief, fused)
        return node, self.settle(reward, pending, belief)

    def get_next_child(self, current_node):
        vals = {}
        for child in current_node.children:
            if child.nqueries == 0:
                return child, False
            vals[child] = child.reward / float(child.nqueries) + self.c * np.sqrt(
                2.0 * np.log(float(current_node.nqueries)) / float(child.nqueries))
        top = max(vals.values())
        return random.choice([key for key in vals.keys() if vals[key] == top]), True


# acquisition callables of this package that have a sync-free device form, by reward name
_FAST_REWARDS = {aqlib.mean_UCB: 'mean', aqlib.mves: 'mes', aqlib.exp_improvement: 'exp_improve'}
# information-gain rewards (reference aq_library.py:34-80, :117-135): fused rollouts on the information-gain belief
_INFO_REWARDS = {aqlib.info_gain: 'info_gain', aqlib.hotspot_info_UCB: 'hotspot_info'}
# rewards that do not depend on the belief at all (reference aq_library.py:339-413)
_BELIEF_FREE_REWARDS = {aqlib.naive: 'naive', aqlib.naive_value: 'naive_value'}


def exploration_constant(f_rew, tree_type):
    """reference :645-660"""
    if f_rew == 'mean':
        return {'belief': 1000, 'dpw': 5000}.get(tree_type, 1.0)
    if f_rew == 'exp_improve':
        return 200
    if f_rew == 'mes':
        return {'belief': 1.0 / np.sqrt(2.0), 'dpw': 1.0}.get(tree_type, 1.0)
    return 1.0


class cMCTS(object):
    """reference :636-713 (the legacy dict-tree base class MCTS :27-274 is not carried over)."""

    def __init__(self, computation_budget, belief, initial_pose, rollout_length, path_generator,
                 aquisition_function, f_rew, T, aq_param=None, use_cost=False, tree_type='dpw'):
        self.GP = belief
        self.cp = initial_pose
        self.path_generator = path_generator
        self.comp_budget = computation_budget
        self.rl = rollout_length
        self.tree = None
        self.aquisition_function = aquisition_function
        self.max_val = None
        self.max_locs = None
        self.target = None
        self.current_max = aq_param
        self.aq_param = aq_param
        self.f_rew = f_rew
        self.t = T
        self.use_cost = use_cost
        self.tree_type = tree_type
        self.c = exploration_constant(f_rew, tree_type)

    def reward_param(self, t):
        """The acquisition parameter of this planning step (reference :668-679); for MES / the naive rewards this draws
        the max-value samples."""
        if self.f_rew == 'mes':
            self.max_val, self.max_locs, self.target = aqlib.sample_max_vals(self.GP, t=t, visualize=True)
            return (self.max_val, self.max_locs, self.target)
        if self.f_rew == 'exp_improve':
            return [self.current_max]
        if self.f_rew in ('naive', 'naive_value'):
            self.max_val, self.max_locs, self.target = aqlib.sample_max_vals(self.GP, t=t, nK=int(self.aq_param[0]),
                                                                             visualize=True, f_rew=self.f_rew)
            return ((self.max_val, self.max_locs, self.target), self.aq_param[1])
        return None

    NATIVE_SEARCH = True       # dpw or belief tree + Dubins / straight-fan generator + a fused reward: the whole search in libipp_b200.so
    native_searches = 0        # how many searches (all instances) went through ipp_tree_search_dpw

    def search(self, t, param, budget=None):
        """Build the tree and run `budget` rollouts (reference :681-695); returns the root's children."""
        if self.tree_type == 'dpw':
            tree_cls = Tree
        elif self.tree_type == 'belief':
            tree_cls = BeliefTree
        else:
            raise ValueError('Tree type must be one of either \'dpw\' or \'belief\'')
        self.tree = tree_cls(self.f_rew, self.aquisition_function, self.GP, self.cp, self.path_generator, t,
                             depth=self.rl, param=param, c=self.c)
        budget = self.comp_budget if budget is None else budget
        self.native_search_used = False
        # the reference's own performance figure brackets exactly this loop (:689-700, "Rollouts completed in ...s")
        time_start = time.time()
        if self.NATIVE_SEARCH and self._search_native(t, budget):
            self.native_search_used = True
            cMCTS.native_searches += 1
        else:
            for _ in range(budget):
                self.tree.get_next_leaf(copy.copy(self.GP))
        self.rollout_seconds = time.time() - time_start
        self.rollouts = budget
        return self.tree.root.children

    def _search_native(self, t, budget):
        """The same search by ipp_tree_search_dpw (csrc/tree.cu): walk, widening (Tree) or random rollouts (BeliefTree),
        candidate paths (Dubins or straight fan), simulated observations and both random streams in native code; per
        rollout one device pass for UCB / EI / MES (points by value, rewards to pinned memory, polled completion), one
        small kernel for naive_value, nothing at all for naive.  Returns False -- with nothing consumed from the random
        streams -- when the configuration is not covered (generator subclasses, other rewards, no data yet) or a rollout
        needs the per-level fallback."""
        import ctypes
        from .paths_library import Dubins_Path_Generator, Path_Generator
        L = aqlib.L
        pg, belief, tree = self.path_generator, self.GP, self.tree
        if not isinstance(belief, GPModel):
            return False
        if type(pg) is Dubins_Path_Generator and pg.use_native:
            generator = 0
            boxes = pg._obstacle_boxes()
            if boxes is None or boxes[1] > 16:
                return False
        elif type(pg) is Path_Generator:
            generator, boxes = 1, (None, 0)          # the straight fan (reference :85-105) looks at no obstacle
        else:
            return False
        fused = tree._fused_kind(belief)
        scale = _prior_scale_table(belief.variance)
        if fused is None or scale is None or belief.dimension not in (2, 3):
            return False
        kind, params, maxes_d = fused
        operand = (belief.kern._c, belief._winv_d, belief._ld, belief._noise_var(), belief._diag_add())
        hotspot = None
        if kind == 'hotspot':                         # two passes per rollout: UCB on this belief, information gain on (I + vK)^-1
            hotspot = belief._info_gain_state()
            kind = L.REWARD_HOTSPOT
        if kind == L.REWARD_INFO:                     # the search runs on the information-gain belief (ipp_b200.h)
            kern2, binv, ldb, _ = belief._info_gain_state()
            operand, params = (kern2, binv, ldb, 0.0, 1.0), [2 * np.pi * np.e]
        naive_locs, rff = None, None
        if kind == 'host':
            # naive: pure geometry against the sampled maximisers, computed inside the native search (no device work at
            # all); naive_value evaluates the sampled functions and stays with the Python tree
            if tree.param is None or tree.param[0][0] is None:
                return False
            max_vals, max_locs, funcs = tree.param[0]
            if self.f_rew == 'naive':
                if max_locs is None:
                    return False
                kind, params = L.REWARD_NAIVE, [float(tree.param[1])]
                naive_locs = np.ascontiguousarray(np.asarray(max_locs, dtype=np.float64)[:, :2])
            else:
                # naive_value: the sampled functions evaluated on the device for all points of a rollout in one pass
                if (funcs is None or len(funcs) != max_vals.shape[0] or len(funcs) == 0
                        or not all(isinstance(f, aqlib.RFFTarget) for f in funcs)
                        or len({f.nFeatures for f in funcs}) != 1):
                    return False
                kind, params = L.REWARD_NAIVE_VALUE, [float(tree.param[1])]
                S_, F_ = len(funcs), funcs[0].nFeatures
                rff = {'W': L.to_device(np.stack([np.asarray(f.W, dtype=np.float64) for f in funcs])),
                       'b': L.to_device(np.stack([np.asarray(f.b, dtype=np.float64).reshape(-1) for f in funcs])),
                       'theta': L.to_device(np.stack([np.asarray(f.theta, dtype=np.float64).reshape(-1) * f.scale for f in funcs])),
                       'maxes': np.ascontiguousarray(np.asarray(max_vals, dtype=np.float64).reshape(-1)),
                       'vals_d': torch.empty(S_ * L.IPP_ROLLOUT_MAX_POINTS + 2 * S_ + 2, dtype=torch.float64, device=L.device()),
                       'vals_h': torch.empty(S_ * L.IPP_ROLLOUT_MAX_POINTS, dtype=torch.float64).pin_memory(),
                       'work': torch.empty(int(L.lib.ipp_rff_workspace(F_, S_, L.IPP_ROLLOUT_MAX_POINTS, 0)),
                                           dtype=torch.float64, device=L.device())}
        pose = np.array([float(v) for v in self.cp], dtype=np.float64)
        if pose.shape[0] != 3:
            return False
        angles = np.ascontiguousarray(np.linspace(-2.35, 2.35, pg.fs), dtype=np.float64)
        prm = L.IppTreeParams()
        prm.budget, prm.max_depth, prm.kind = int(budget), int(self.rl), int(kind)
        prm.frontier_size, prm.keep_every, prm.n_obstacles = int(pg.fs), 10, int(boxes[1])
        prm.c = float(self.c)
        prm.reward_param = float(params[0]) if params is not None else 0.0
        prm.time = float(t)
        prm.horizon, prm.turning_radius = float(pg.hl), float(pg.tr)
        prm.dense_step, prm.border = float(pg.ss / 10), float(3 * pg.tr)
        prm.generator, prm.sample_step = generator, float(pg.ss)
        log_table = None
        if isinstance(tree, BeliefTree):
            # UCB1 uses np.log of the parent's visit count: hand the interpreter's own values over (numpy's SIMD log need
            # not agree with libm in the last bit), computed the way the Python tree computes them: one scalar at a time
            log_table = np.array([0.0] + [np.log(float(n)) for n in range(1, int(budget) + 2)], dtype=np.float64)
            prm.tree_type, prm.log_table = 1, log_table.ctypes.data
        prm.extent[:] = [float(e) for e in pg.extent]
        prm.angles = angles.ctypes.data
        prm.obstacles = ctypes.cast(boxes[0], ctypes.c_void_p).value if boxes[1] else None
        prm.prior_scale = scale.ctypes.data
        if hotspot is not None:
            prm.info_kern = ctypes.cast(ctypes.pointer(hotspot[0]), ctypes.c_void_p).value
            prm.info_winv, prm.info_ldw, prm.info_param = hotspot[1].data_ptr(), int(hotspot[2]), float(2 * np.pi * np.e)
        if naive_locs is not None:
            prm.naive_locs, prm.n_naive = naive_locs.ctypes.data, int(naive_locs.shape[0])
        if rff is not None:
            prm.n_naive, prm.rff_features = int(rff['W'].shape[0]), int(rff['W'].shape[1])
            prm.rff_W, prm.rff_b, prm.rff_theta = rff['W'].data_ptr(), rff['b'].data_ptr(), rff['theta'].data_ptr()
            prm.naive_maxes = rff['maxes'].ctypes.data
            prm.rff_vals_dev, prm.rff_vals_host = rff['vals_d'].data_ptr(), rff['vals_h'].data_ptr()
            prm.rff_work = rff['work'].data_ptr()
        dev = L.device()
        stage = _native_stage.get(dev.index)
        if stage is None:
            stage = _native_stage[dev.index] = (torch.empty(L.IPP_TREE_STAGE_DOUBLES, dtype=torch.float64).pin_memory(),
                                                torch.empty(L.IPP_TREE_STAGE_DOUBLES, dtype=torch.float64, device=dev))
        N = belief.n_obs
        work = L.workspace(L.lib.ipp_rollout_workspace(N, L.IPP_ROLLOUT_MAX_POINTS), slot='rollout')
        G = int(pg.fs) + 1
        n_root = np.zeros(1, dtype=np.int32)
        root_goal = np.zeros(G, dtype=np.int32)
        root_visits = np.zeros(G, dtype=np.int64)
        root_reward = np.zeros(G, dtype=np.float64)
        evals = np.zeros(1, dtype=np.int64)
        status = np.zeros(1, dtype=np.int32)
        py_state, py_extra = L.py_rng_state()
        np_state = L.np_rng_state()
        vp = ctypes.c_void_p
        S = 0 if maxes_d is None else int(maxes_d.numel())
        L.check(L.lib.ipp_tree_search_dpw(
            ctypes.byref(prm), pose.ctypes.data_as(vp), ctypes.byref(operand[0]), L.ptr(belief._X_d), N,
            L.ptr(belief._alpha_d), L.ptr(operand[1]), operand[2], operand[3], operand[4],
            L.ptr(maxes_d), S, ctypes.byref(py_state), ctypes.byref(np_state), L.ptr(stage[0]), L.ptr(stage[1]),
            L.ptr(work), n_root.ctypes.data_as(vp), root_goal.ctypes.data_as(vp), root_visits.ctypes.data_as(vp),
            root_reward.ctypes.data_as(vp), evals.ctypes.data_as(vp), status.ctypes.data_as(vp), L.stream_ptr()))
        if int(status[0]) != 0:
            return False
        L.set_py_rng_state(py_state, py_extra)
        L.set_np_rng_state(np_state)
        paths, dense_paths = pg.get_path_set_lazy(self.cp) if generator == 0 else pg.get_path_set(self.cp)
        keys = list(paths.keys())
        n = int(n_root[0])
        if keys != [int(g) for g in root_goal[:n]]:
            raise RuntimeError('native tree search and the path generator disagree on the root actions')
        root = tree.root
        for i, key in enumerate(keys):
            child = Node(pose=root.pose, parent=root, name=root.name + '_action' + str(i), action=paths[key],
                         dense_path=dense_paths[key])
            child.nqueries = int(root_visits[i])
            child.reward = float(root_reward[i])
            root.add_children(child)
        root.nqueries = int(budget)
        root.reward = float(np.sum(root_reward[:n]))
        tree.reward_evals = int(evals[0])
        return True

    def result(self, best_child, all_vals):
        """The tuple choose_trajectory returns (reference :709-713)."""
        paths, dense_paths = self.path_generator.get_path_set(self.cp)
        best_dense = best_child.dense_path
        if hasattr(best_dense, 'tolist'):
            best_dense = best_dense.tolist()                  # the reference returns a list of (x, y, heading) tuples
        return (best_child.action, best_dense, all_vals[self.tree.root.children.index(best_child)], paths,
                all_vals, self.max_locs, self.max_val, self.target)

    def choose_trajectory(self, t):
        children = self.search(t, self.reward_param(t))
        most = max(n.nqueries for n in children)
        best_child = random.choice([node for node in children if node.nqueries == most])
        all_vals = {i: child.reward / float(child.nqueries) for i, child in enumerate(children)}
        return self.result(best_child, all_vals)


# ------------------------------------------------------------------------------------------ batched seams
_F_REW_TO_BATCH = {'mean': 'mean', 'mes': 'mes', 'exp_improve': 'exp_improve', 'info_gain': 'info_gain',
                   'hotspot_info': 'hotspot_info'}


def evaluate_root_actions(belief, paths, f_rew, t, param=None):
    """Score every candidate path of one pose in a single GPU call.  `paths`: dict key -> point list."""
    keys = list(paths.keys())
    values = aqlib.reward_paths(_F_REW_TO_BATCH[f_rew], t, [paths[k] for k in keys], belief, param)
    return dict(zip(keys, values))


def choose_myopic(belief, path_generator, pose, f_rew, t, param=None, use_cost=False):
    """Robot.choose_trajectory's path loop (reference robot_library.py:173-208) with one batched reward
    call.  Returns (sampling path, dense path, value, all paths, all values) or None."""
    paths, true_paths = path_generator.get_path_set(pose)
    if len(paths) == 0:
        return None
    value = evaluate_root_actions(belief, paths, f_rew, t, param)
    if use_cost:
        for k in list(value.keys()):
            cost = float(path_generator.path_cost(true_paths[k]))
            value[k] = value[k] / (cost if cost != 0.0 else 100.0)
    try:
        top = max(value.values())
        best_key = np.random.choice([key for key in value.keys() if value[key] == top])
        return paths[best_key], true_paths[best_key], value[best_key], paths, value
    except Exception:
        return None


_grid_cache = {}


def predict_max(belief, ranges, t=0, time=0):
    """Robot.predict_max (reference robot_library.py:222-254): location and value of the posterior-mean maximum over the
    100 x 100 grid of the world -- one batched device predict (the grid stays resident on the device between steps)."""
    if belief.xvals is None:
        return np.array([0., 0.]), 0.
    key = (tuple(float(r) for r in ranges), int(belief.dimension), float(time) if belief.dimension == 3 else 0.0,
           torch.cuda.current_device())
    entry = _grid_cache.get(key)
    if entry is None:
        x1, x2 = np.meshgrid(np.linspace(ranges[0], ranges[1], 100), np.linspace(ranges[2], ranges[3], 100),
                             sparse=False, indexing='xy')
        if belief.dimension == 2:
            data = np.vstack([x1.ravel(), x2.ravel()]).T
        else:
            data = np.vstack([x1.ravel(), x2.ravel(), time * np.ones(len(x1.ravel()))]).T
        if len(_grid_cache) > 8:
            _grid_cache.clear()
        entry = _grid_cache[key] = (data, aqlib.L.to_device(data))
    data, data_d = entry
    mean, _ = belief.predict_device(data_d, precision='fp64')
    top = torch.argmax(mean)                         # first index of the maximum, like np.argmax
    return data[int(top.item()), :], float(mean[top].item())
