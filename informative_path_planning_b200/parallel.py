"""Multi-GPU partitioning of the hot path: one process per GPU, torch.distributed.

The path shards naturally: query points and whole candidate paths are independent given the belief
(X, alpha, W^-1).  So
  * the belief is REPLICATED: a new observation is broadcast from the source rank (k x (D+1) doubles)
    and every rank applies the same FP64 refresh / block append locally -- cheaper than shipping the
    N x N inverse (134 MB at N=4096) and bit-identical across ranks because every rank runs the same
    deterministic kernels on the same inputs; `mode='broadcast'` ships the factor instead;
  * queries / paths are SHARDED contiguously (never splitting a path), with no data-path collective;
  * per-path rewards are returned with one all_gather (64k paths x 8 B: latency-bound);
  * max-value sampling over a large grid: mesh rows sharded, the S (max, arg-max) pairs gathered and reduced.
Collectives run on NCCL over NVLink when the tensors are CUDA tensors and on gloo for the CPU tests
of the host-side logic (the functions below never look at the device of what they move).
"""
import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n, rank, size):
    """Contiguous, balanced [lo, hi) split of n items: the first n % size ranks get one extra."""
    base, extra = divmod(int(n), int(size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_paths(offsets, rank, size):
    """Whole paths only: rank r owns paths [plo, phi) and therefore points [offsets[plo], offsets[phi])."""
    offsets = np.asarray(offsets, dtype=np.int64)
    plo, phi = shard_bounds(offsets.shape[0] - 1, rank, size)
    return plo, phi, int(offsets[plo]), int(offsets[phi])


def _comm_device():
    if dist.get_backend() == 'nccl':
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


def broadcast_array(a, src=0):
    """Broadcast a float64 numpy array (None on non-source ranks) -> numpy array on every rank."""
    rank, size = world()
    if size == 1:
        return np.asarray(a, dtype=np.float64)
    dev = _comm_device()
    shape = torch.zeros(4, dtype=torch.int64, device=dev)
    if rank == src:
        a = np.ascontiguousarray(a, dtype=np.float64)
        shape[:a.ndim + 1] = torch.tensor([a.ndim] + list(a.shape), dtype=torch.int64)
    dist.broadcast(shape, src)
    ndim = int(shape[0])
    dims = [int(s) for s in shape[1:1 + ndim]]
    buf = torch.from_numpy(a).to(dev) if rank == src else torch.empty(dims, dtype=torch.float64, device=dev)
    dist.broadcast(buf, src)
    return buf.cpu().numpy()


def add_data_replicated(model, xvals, zvals, src=0, mode='recompute'):
    """model.add_data on every rank with the observation that only `src` holds.
    mode='recompute': broadcast (xvals, zvals), every rank refreshes locally (default).
    mode='broadcast': `src` refreshes, then W^-1 and alpha are broadcast (NCCL over NVLink)."""
    rank, size = world()
    xvals = broadcast_array(xvals, src)
    zvals = broadcast_array(zvals, src)
    if size == 1 or mode == 'recompute':
        model.add_data(xvals, zvals)
        return model
    if mode != 'broadcast':
        raise ValueError("mode must be 'recompute' or 'broadcast'")
    if rank == src:
        model.add_data(xvals, zvals)
    else:
        model.attach_state_shell(xvals, zvals)           # host arrays + empty device buffers of the right shape
    for t in model.state_tensors():
        dist.broadcast(t, src)
    return model


def gather_concat(local, counts):
    """all_gather of ragged 1-D float64 tensors (counts[r] items from rank r) -> concatenated tensor."""
    rank, size = world()
    if size == 1:
        return local
    width = int(max(counts))
    pad = torch.zeros(width, dtype=local.dtype, device=local.device)
    pad[:local.numel()] = local
    out = [torch.empty_like(pad) for _ in range(size)]
    dist.all_gather(out, pad)
    return torch.cat([o[:int(c)] for o, c in zip(out, counts)])


def predict_sharded(model, xq):
    """Posterior mean/var of (M, D) numpy queries: each rank predicts its contiguous slice on its GPU,
    the slices are all-gathered.  Returns (mean (M,1), var (M,1)) on every rank."""
    rank, size = world()
    M = xq.shape[0]
    lo, hi = shard_bounds(M, rank, size)
    counts = [shard_bounds(M, r, size)[1] - shard_bounds(M, r, size)[0] for r in range(size)]
    from . import _lib as L
    mean, var = model.predict_device(L.to_device(xq[lo:hi])) if hi > lo else (
        torch.empty(0, dtype=torch.float64, device=L.device()), torch.empty(0, dtype=torch.float64, device=L.device()))
    mean, var = gather_concat(mean, counts), gather_concat(var, counts)
    return mean.cpu().numpy().reshape(-1, 1), var.cpu().numpy().reshape(-1, 1)


def reward_paths_sharded(model, f_rew, time, queries, offsets, param=None, reward_fn=None):
    """Rewards of P paths, P/size whole paths per rank, one all_gather.  `reward_fn(f_rew, time, q, offs,
    model, param) -> float64 [p]` defaults to aq_library.reward_paths (injected in the CPU tests)."""
    rank, size = world()
    offsets = np.asarray(offsets, dtype=np.int64)
    P = offsets.shape[0] - 1
    plo, phi, qlo, qhi = shard_paths(offsets, rank, size)
    if reward_fn is None:
        from . import aq_library as aq

        def reward_fn(f, t, q, o, m, p):
            return aq.reward_paths(f, t, q, m, p, offsets=o)
    if phi > plo:
        local = np.asarray(reward_fn(f_rew, time, queries[qlo:qhi], offsets[plo:phi + 1] - qlo, model, param), dtype=np.float64)
    else:
        local = np.zeros(0)
    counts = [shard_bounds(P, r, size)[1] - shard_bounds(P, r, size)[0] for r in range(size)]
    dev = _comm_device() if size > 1 else torch.device('cpu')
    out = gather_concat(torch.from_numpy(local).to(dev), counts)
    return out.cpu().numpy()


def combine_sharded_argmax(max_local, arg_global_local):
    """(max, first arg-max) over grid slices held by different ranks ("MES additionally needs the S
    max-values: computed on sharded grid slices, then reduced over the (value, index) pairs").  max_local [S] float64 and
    arg_global_local [S] int64 (indices into the WHOLE grid; a rank without points passes -inf / any index) on the
    communication device.  One all_gather of 2 S numbers per rank; ties go to the smallest index, like np.argmax.
    Returns (max [S], arg-max [S]) on every rank."""
    rank, size = world()
    if size == 1:
        return max_local, arg_global_local
    pair = torch.stack([max_local.to(torch.float64), arg_global_local.to(torch.float64)])      # indices < 2^53: exact
    out = [torch.empty_like(pair) for _ in range(size)]
    dist.all_gather(out, pair)
    vals = torch.stack([o[0] for o in out])                 # [size][S]
    idxs = torch.stack([o[1] for o in out])
    best = vals.max(dim=0).values
    cand = torch.where(vals == best.unsqueeze(0), idxs, torch.full_like(idxs, float('inf')))
    return best, cand.min(dim=0).values.to(torch.int64)


def sample_max_vals_shared_sharded(model, xs, ys, nK=100, nFeatures=1000, seed=0, eval_fn=None):
    """Shared-feature max-value sampling (aq_library.sample_max_vals_shared, BASELINE config 4) over the meshgrid of xs and
    ys with the MESH ROWS sharded across the ranks: every rank draws the same features and the same posterior weights
    from the replicated belief (same seed, deterministic kernels), evaluates the nK sampled functions on its rows
    ys[lo:hi] only, and the per-sample (max, arg-max) pairs meet in combine_sharded_argmax.
    `eval_fn(W, b, Theta, variance, xs, ys_slice) -> (max [S], arg-max [S] within the slice's mesh)` defaults to the
    device grid stage (injected in the CPU tests).  Returns (max_vals (nK, 1), max_locs (nK, 2), (W, b, Theta))."""
    rank, size = world()
    xs, ys = np.asarray(xs, dtype=np.float64).reshape(-1), np.asarray(ys, dtype=np.float64).reshape(-1)
    nx, ny = xs.shape[0], ys.shape[0]
    lo, hi = shard_bounds(ny, rank, size)
    variance = float(np.asarray(model.variance, dtype=float).reshape(-1)[0])
    if eval_fn is None:
        from . import _lib as L
        from . import aq_library as aq
        W, b, Theta = aq.shared_feature_draws(model, nK, nFeatures, np.random.default_rng(seed))

        def eval_fn(W_, b_, Th_, var_, xs_, ys_):
            return aq.rff_eval_argmax_mesh(W_, b_, Th_, var_, L.to_device(xs_), L.to_device(ys_))
    else:
        W, b, Theta = model.shared_feature_draws(nK, nFeatures, np.random.default_rng(seed))
    dev = _comm_device() if size > 1 else None
    if hi > lo:
        mx, am = eval_fn(W, b, Theta, variance, xs, ys[lo:hi])
        mx = torch.as_tensor(mx, dtype=torch.float64)
        am = torch.as_tensor(am, dtype=torch.int64) + lo * nx          # index into the whole mesh: g = j * nx + i
    else:
        mx = torch.full((nK,), -float('inf'), dtype=torch.float64)
        am = torch.zeros(nK, dtype=torch.int64)
    if dev is not None:
        mx, am = mx.to(dev), am.to(dev)
    mx, am = combine_sharded_argmax(mx, am)
    am_h = am.cpu().numpy()
    locs = np.column_stack([xs[am_h % nx], ys[am_h // nx]])
    Theta_h = np.ascontiguousarray(Theta.cpu().numpy().T) if isinstance(Theta, torch.Tensor) else Theta
    return mx.cpu().numpy().reshape(-1, 1), locs, (W, b, Theta_h)


def mcts_root_parallel(mcts, t, seed=None):
    """Root-parallel tree search for one planning step (rollouts per rank).

    Every rank holds the replicated belief and runs its OWN tree (mcts_library.cMCTS.search) on its GPU with
    computation_budget / world_size rollouts and its own random stream; the root statistics (visit count and reward
    sum per root action: <= 16 x 2 doubles) are summed with ONE all_reduce and every rank picks the same action: most
    visits, ties by mean reward, then by index.  The candidate paths of the root pose are deterministic, so the root
    actions line up across ranks.  The acquisition parameter (MES max-value samples) is drawn from an identically
    seeded stream on every rank, so it is the same everywhere without shipping the sampled functions.

    This is a different search from the reference's single sequential tree (the ranks do not see each other's
    statistics while searching), so it is an opt-in mode; with world_size 1 it IS mcts.choose_trajectory(t).
    Returns the tuple of cMCTS.choose_trajectory."""
    import random
    rank, size = world()
    if size == 1:
        return mcts.choose_trajectory(t)
    dev = _comm_device()
    if seed is None:                                             # agree on a seed: rank 0 draws it
        s = torch.tensor([np.random.randint(0, 2 ** 31 - 1) if rank == 0 else 0], dtype=torch.int64, device=dev)
        dist.broadcast(s, 0)
        seed = int(s.item())
    np.random.seed(seed % (2 ** 32))
    random.seed(seed)
    param = mcts.reward_param(t)                                 # same draws on every rank
    np.random.seed((seed + 7919 * (rank + 1)) % (2 ** 32))
    random.seed(seed + 7919 * (rank + 1))
    lo, hi = shard_bounds(mcts.comp_budget, rank, size)
    # every rank runs at least one rollout (comp_budget < world_size would leave a rank without root children), and a
    # rank whose search produced no children (no feasible action from the root pose) still reaches the collectives
    children = mcts.search(t, param, budget=max(1, hi - lo)) or []
    stats = torch.tensor([[float(c.nqueries), float(c.reward)] for c in children], dtype=torch.float64,
                         device=dev).reshape(-1, 2)
    count = torch.tensor([stats.shape[0]], dtype=torch.int64, device=dev)
    lo_c, hi_c = count.clone(), count.clone()
    dist.all_reduce(lo_c, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi_c, op=dist.ReduceOp.MAX)
    if int(lo_c.item()) != int(hi_c.item()):
        raise RuntimeError('root actions differ across ranks (%d vs %d): the belief or the path generator is not '
                           'replicated' % (int(lo_c.item()), int(hi_c.item())))
    if not children:
        raise RuntimeError('no feasible action from the root pose (every rank agrees)')
    mcts.local_root_stats = stats.cpu().numpy().copy()
    dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    stats = stats.cpu().numpy()
    visits, sums = stats[:, 0], stats[:, 1]
    means = np.where(visits > 0, sums / np.maximum(visits, 1.0), 0.0)
    order = sorted(range(len(children)), key=lambda i: (-visits[i], -means[i], i))
    mcts.root_stats = stats
    return mcts.result(children[order[0]], {i: float(means[i]) for i in range(len(children))})
