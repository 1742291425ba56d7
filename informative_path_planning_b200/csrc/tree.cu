// Native DPW tree search for one planning step (include/ipp_b200.h: ipp_tree_search_dpw).
//
// What runs here is the host side of the reference's Tree (mcts_library.py:314-488): get_next_leaf / leaf_helper (the
// walk), get_next_child (DPW-UCT), build_action_children (frontier goals + Dubins candidate paths, shared with
// paths.cu through dubins.cuh), the progressive-widening rule and the simulated prior observations of the 'BA' step,
// and backprop -- with the two random streams the reference consumes (CPython's `random`, numpy's legacy global
// RandomState) continued bit for bit, so that the search is the one the interpreter would have run.  Rewards: one
// ipp_rollout_rewards launch sequence per rollout (rollout.cu), inputs staged through pinned memory, one stream
// synchronisation per rollout (back-propagation needs the value before the next selection).
#include <math.h>
#include <string.h>

#include <chrono>
#include <vector>

#include "dubins.cuh"
#include "internal.cuh"

namespace {

using namespace ipp_dubins;

// ---- MT19937, the generator under both `random` and numpy's legacy RandomState
struct Mt {
    uint32_t* mt;
    int32_t* pos;
    void regen() {
        const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MAG = 0x9908b0dfu;
        int kk = 0;
        uint32_t y;
        for (; kk < 624 - 397; ++kk) {
            y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
            mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
        }
        for (; kk < 623; ++kk) {
            y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
            mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
        }
        y = (mt[623] & UPPER) | (mt[0] & LOWER);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
        *pos = 0;
    }
    uint32_t next() {
        if (*pos >= 624) regen();
        uint32_t y = mt[(*pos)++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
};

// random.choice(seq) -> index: seq[_randbelow(len(seq))], _randbelow_with_getrandbits (CPython Lib/random.py)
inline int py_choice(Mt& g, int n) {
    int k = 0;
    for (unsigned v = (unsigned)n; v; v >>= 1) ++k;          // n.bit_length()
    uint32_t r = g.next() >> (32 - k);
    while (r >= (uint32_t)n) r = g.next() >> (32 - k);
    return (int)r;
}

// numpy legacy: random_standard_normal of the global RandomState = legacy_gauss (polar method, cached second variate)
inline double np_double(Mt& g) {
    const int32_t a = (int32_t)(g.next() >> 5), b = (int32_t)(g.next() >> 6);
    return (a * 67108864.0 + b) / 9007199254740992.0;
}
inline double np_gauss(Mt& g, ipp_mt_state* st) {
    if (st->has_gauss) {
        const double t = st->gauss;
        st->has_gauss = 0;
        st->gauss = 0.0;
        return t;
    }
    double f, x1, x2, r2;
    do {
        x1 = 2.0 * np_double(g) - 1.0;
        x2 = 2.0 * np_double(g) - 1.0;
        r2 = x1 * x1 + x2 * x2;
    } while (r2 >= 1.0 || r2 == 0.0);
    f = sqrt(-2.0 * log(r2) / r2);
    st->gauss = f * x1;
    st->has_gauss = 1;
    return f * x2;
}

// np.random.randint(0, high) of the legacy global RandomState (int64 path, range < 2^32): masked rejection over 32-bit
// draws; a one-value range consumes nothing
inline int np_randint(Mt& g, int high) {
    const uint32_t rng = (uint32_t)(high - 1);
    if (rng == 0) return 0;
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t val;
    do { val = g.next() & mask; } while (val > rng);
    return (int)val;
}

// np.sum of a contiguous float64 vector (numpy's pairwise summation: < 8 elements sequentially, else 8 running sums
// combined as a tree, remainder appended; blocks of 128 are not reached by a path's <= 32 points)
inline double np_pairwise_sum(const double* a, int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += a[i];
        return r;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

// ---- candidate paths of one pose: Path_Generator.generate_frontier_points (paths_library.py:44-83) + buffered_paths
struct PathSet {
    std::vector<int32_t> goal;       // goal index of each kept path
    std::vector<int32_t> count;      // kept samples
    std::vector<double> samp;        // [path][max_samp][3]
    int max_samp = 0;
};

int path_limits(const ipp_tree_params& p, int* max_samp) {
    const double longest = p.horizon + 8.0 * 3.14159265358979323846 * p.turning_radius;   // hl + a few turning circles
    const int max_dense = (int)(longest / p.dense_step) + 16;
    *max_samp = max_dense / p.keep_every + 2;
    return max_dense;
}

// goals[(fs + 1) * 3] -> number of goals (the pose itself closes the list)
int frontier_goals(const ipp_tree_params& p, const double* cp, double* goals) {
    const int fs = p.frontier_size;
    int G = 0;
    const double cx = cp[0], cy = cp[1], ch = cp[2], hl = p.horizon, tr = p.turning_radius;
    const double x_lo = p.extent[0] + 3 * tr, x_hi = p.extent[1] - 3 * tr;
    const double y_lo = p.extent[2] + 3 * tr, y_hi = p.extent[3] - 3 * tr;
    for (int i = 0; i < fs; ++i) {
        const double a = ch + p.angles[i];
        const double x = hl * cos(a) + cx, y = hl * sin(a) + cy;
        if (sqrt((cx - x) * (cx - x) + (cy - y) * (cy - y)) <= tr) continue;
        if (x > x_hi || x < x_lo) continue;
        if (y > y_hi || y < y_lo) continue;
        goals[3 * G] = x; goals[3 * G + 1] = y; goals[3 * G + 2] = a;
        ++G;
    }
    goals[3 * G] = cx; goals[3 * G + 1] = cy; goals[3 * G + 2] = ch;
    return G + 1;
}

// kept samples of the path pose -> goal (0 = dropped); samp[max_samp][3]
inline int goal_path(const ipp_tree_params& p, const double* cp, const double* goal, double* samp, int max_samp,
                     int max_dense) {
    if (p.generator == 1) {
        // Path_Generator.make_sample_paths (paths_library.py:85-105): numpy scalar arithmetic restated (x ** 2 is pow(x, 2),
        // round() is Python 2's: C round(), halves away from zero); a path exists as soon as it has one point
        const double distance = sqrt(pow(cp[0] - goal[0], 2.0) + pow(cp[1] - goal[1], 2.0));
        const int n = (int)round(distance / p.sample_step);
        if (n < 1) return 0;
        if (n > max_samp) return -1;
        for (int j = 0; j < n; ++j) {
            const double step = (double)(j + 1) * p.sample_step;
            samp[3 * j] = cp[0] + step * cos(goal[2]);
            samp[3 * j + 1] = cp[1] + step * sin(goal[2]);
            samp[3 * j + 2] = goal[2];
        }
        return n;
    }
    int dn = 0;
    const int n = dubins_one(cp, goal, p.turning_radius, p.dense_step, p.keep_every, p.extent, p.border, p.obstacles,
                             p.n_obstacles, samp, max_samp, nullptr, max_dense, &dn, nullptr);
    return n < 2 ? 0 : n;
}

void build_path_set(const ipp_tree_params& p, const double* cp, PathSet& out) {
    std::vector<double> goals((size_t)(p.frontier_size + 1) * 3);
    const int G = frontier_goals(p, cp, goals.data());
    int max_samp = 0;
    const int max_dense = path_limits(p, &max_samp);
    out.max_samp = max_samp;
    out.goal.clear(); out.count.clear(); out.samp.clear();
    std::vector<double> tmp((size_t)max_samp * 3);
    for (int gi = 0; gi < G; ++gi) {
        const int n = goal_path(p, cp, &goals[3 * gi], tmp.data(), max_samp, max_dense);
        if (n <= 0) continue;
        out.goal.push_back(gi);
        out.count.push_back(n);
        out.samp.insert(out.samp.end(), tmp.begin(), tmp.end());
    }
}

// ---- the tree
struct Node {
    bool action = false;            // 'BA' (an action of its parent belief node) or 'B'
    bool has_children = false;      // python: children is not None
    int depth = 0;
    int parent = -1;
    int goal = -1;                  // action nodes: goal index in the parent's path set
    int k = 0;                      // action nodes: sample points; belief nodes: length of zvals
    size_t data = 0;                // action nodes: offset of samp[k][3] in the pool; belief nodes: offset of zvals[k]
    double pose[3] = {0, 0, 0};
    double end[3] = {0, 0, 0};      // action nodes: last kept sample (pose of the belief children)
    int64_t nq = 0;
    double reward = 0.0;
    std::vector<int> children;
    // belief nodes: candidate actions are materialised one goal at a time (see expand_one)
    int n_goals = -1;               // -1: frontier goals not generated yet
    int g_next = 0;                 // next goal whose path has not been tried
    size_t goals = 0;               // offset of goals[n_goals][3] in the pool
};

bool valid_params(const ipp_tree_params* p) {
    return p && p->angles && p->prior_scale && p->budget >= 0 && p->max_depth >= 1 && p->frontier_size >= 1 &&
           (p->tree_type == 0 || (p->tree_type == 1 && p->log_table)) &&
           (p->generator == 0 || (p->generator == 1 && p->sample_step > 0)) &&
           p->keep_every >= 1 && p->turning_radius > 0 && p->dense_step > 0 && p->n_obstacles >= 0 &&
           p->n_obstacles <= IPP_MAX_OBSTACLES && (p->n_obstacles == 0 || p->obstacles);
}

// naive_value: the n_naive sampled functions at the M points of one rollout.  One block; a warp per (function, point)
// pair, lanes striding the F features (the grid evaluator of rff.cu maps one THREAD to a point, which is the right
// shape for 10^5 grid points and a 30 us serial loop for 20).  Points arrive by value, values go straight to pinned
// host memory, `done` is set to `seq` once they are visible there.
__global__ void __launch_bounds__(1024) rff_points_kernel(const double* __restrict__ W, const double* __restrict__ b,
                                                          const double* __restrict__ theta, int F, int S, int D,
                                                          const __grid_constant__ ipp::RolloutInline pp, int M,
                                                          double* __restrict__ vals, unsigned int* done, unsigned int seq) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int pair = warp; pair < S * M; pair += nwarps) {
        const int sidx = pair / M, m = pair % M;
        const double* w = W + (size_t)sidx * F * D;
        double x[3] = {pp.v[m * D], pp.v[m * D + 1], D == 3 ? pp.v[m * D + 2] : 0.0};
        double acc = 0.0;
        for (int f = lane; f < F; f += 32) {
            double arg = b[(size_t)sidx * F + f];
            for (int d = 0; d < D; ++d) arg = fma(w[(size_t)f * D + d], x[d], arg);
            acc = fma(theta[(size_t)sidx * F + f], cos(arg), acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            vals[(size_t)sidx * M + m] = acc;
            __threadfence_system();
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned int*>(done) = seq;
    }
}

double g_walk_us = 0.0, g_device_us = 0.0;      // host-side split of the last search (ipp_tree_last_timing)
double g_launch_us = 0.0;                       // of g_device_us: time inside the launch calls themselves
inline double now_us() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" int ipp_abi_struct_sizes(int32_t out[4]) {
    IPP_CHECK_ARG(out, "null pointer");
    out[0] = (int32_t)sizeof(ipp_rbf);
    out[1] = (int32_t)sizeof(ipp_obstacle_box);
    out[2] = (int32_t)sizeof(ipp_mt_state);
    out[3] = (int32_t)sizeof(ipp_tree_params);
    return IPP_OK;
}

extern "C" int ipp_tree_last_launch_us(double* launch_us) {
    IPP_CHECK_ARG(launch_us, "null pointer");
    *launch_us = g_launch_us;
    return IPP_OK;
}

extern "C" int ipp_tree_last_timing(double* walk_us, double* device_us) {
    IPP_CHECK_ARG(walk_us && device_us, "null pointer");
    *walk_us = g_walk_us;
    *device_us = g_device_us;
    return IPP_OK;
}

extern "C" int ipp_rng_py_choices(ipp_mt_state* py_rng, const int32_t* ns, int count, int32_t* out) {
    IPP_CHECK_ARG(py_rng && ns && out && count >= 0, "bad argument");
    Mt g{py_rng->key, &py_rng->pos};
    for (int i = 0; i < count; ++i) {
        IPP_CHECK_ARG(ns[i] >= 1, "choice from an empty sequence");
        out[i] = py_choice(g, ns[i]);
    }
    return IPP_OK;
}

extern "C" int ipp_rng_np_normals(ipp_mt_state* np_rng, int count, double* out) {
    IPP_CHECK_ARG(np_rng && out && count >= 0, "bad argument");
    Mt g{np_rng->key, &np_rng->pos};
    for (int i = 0; i < count; ++i) out[i] = np_gauss(g, np_rng);
    return IPP_OK;
}

extern "C" int ipp_tree_path_set(const ipp_tree_params* prm, const double* pose, int32_t* n_paths, int32_t* goal_index,
                                 int32_t* samp_count, double* samp, int max_samp) {
    IPP_CHECK_ARG(valid_params(prm) && pose && n_paths && goal_index && samp_count && samp, "bad argument");
    PathSet ps;
    build_path_set(*prm, pose, ps);
    IPP_CHECK_ARG(max_samp >= ps.max_samp, "max_samp too small");
    *n_paths = (int32_t)ps.goal.size();
    for (size_t i = 0; i < ps.goal.size(); ++i) {
        goal_index[i] = ps.goal[i];
        samp_count[i] = ps.count[i];
        memcpy(samp + i * (size_t)max_samp * 3, &ps.samp[i * (size_t)ps.max_samp * 3], sizeof(double) * 3 * ps.count[i]);
    }
    return IPP_OK;
}

extern "C" int ipp_tree_search_dpw(const ipp_tree_params* prm, const double* root_pose,
                                   const ipp_rbf* kern, const double* X, int64_t N, const double* alpha,
                                   const double* winv, int64_t ldw, double noise_pred, double diag_add,
                                   const double* maxes_dev, int64_t S,
                                   ipp_mt_state* py_rng, ipp_mt_state* np_rng,
                                   double* stage_host, double* stage_dev, double* work,
                                   int32_t* n_root, int32_t* root_goal, int64_t* root_visits, double* root_reward,
                                   int64_t* reward_evals, int32_t* status, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    IPP_CHECK_ARG(valid_params(prm), "bad tree parameters");
    const bool value_reward = prm->kind == IPP_REWARD_NAIVE_VALUE;
    const bool host_reward = prm->kind == IPP_REWARD_NAIVE || value_reward;        // rewards that read no belief
    IPP_CHECK_ARG(root_pose && kern && py_rng && np_rng, "null pointer");
    IPP_CHECK_ARG(host_reward || (X && alpha && winv && stage_host && stage_dev && work), "null pointer");
    IPP_CHECK_ARG(n_root && root_goal && root_visits && root_reward && reward_evals && status, "null output pointer");
    IPP_CHECK_ARG((host_reward || N >= 1) && (kern->dim == 2 || kern->dim == 3), "needs a belief with data, 2 or 3 input dimensions");
    const bool hotspot = prm->kind == IPP_REWARD_HOTSPOT;
    IPP_CHECK_ARG(prm->kind == IPP_REWARD_UCB || prm->kind == IPP_REWARD_EI || prm->kind == IPP_REWARD_MES ||
                  prm->kind == IPP_REWARD_INFO || hotspot || host_reward, "unknown reward kind");
    IPP_CHECK_ARG(!hotspot || (prm->info_kern && prm->info_winv && prm->info_ldw >= N && prm->info_kern->dim == kern->dim),
                  "the hotspot reward needs the information-gain belief over the same observations");
    IPP_CHECK_ARG(prm->kind != IPP_REWARD_NAIVE || (prm->naive_locs && prm->n_naive >= 1),
                  "the naive reward needs the sampled maximisers");
    IPP_CHECK_ARG(!value_reward || (prm->rff_W && prm->rff_b && prm->rff_theta && prm->naive_maxes && prm->rff_vals_dev &&
                                    prm->rff_vals_host && prm->rff_work && prm->n_naive >= 1 && prm->rff_features >= 1 &&
                                    stage_host && stage_dev),
                  "the naive_value reward needs the sampled functions and its scratch buffers");
    const ipp_tree_params& P = *prm;
    const int D = kern->dim;
    // work on copies of the generators: they are only written back when the whole search succeeded
    ipp_mt_state py = *py_rng, np = *np_rng;
    Mt gpy{py.key, &py.pos}, gnp{np.key, &np.pos};

    std::vector<Node> nodes;
    nodes.reserve((size_t)P.budget * (2 * P.max_depth + 2) + 64);
    std::vector<double> pool;
    pool.reserve(1 << 16);
    nodes.emplace_back();
    memcpy(nodes[0].pose, root_pose, sizeof(double) * 3);

    struct Level { size_t data; size_t z; int k; int reps; };      // pool offsets of the points [k][3] and the observation [k]
    std::vector<Level> levels;
    std::vector<int> ties;
    int64_t evals = 0;
    *status = 0;
    double* h_in = stage_host;                                                   // [M][D] points, then [M] observations
    double* h_out = stage_host + IPP_ROLLOUT_MAX_POINTS * (IPP_MAX_DIM + 1);      // [L] rewards + info
    double* d_in = stage_dev;
    double* d_out = stage_dev + IPP_ROLLOUT_MAX_POINTS * (IPP_MAX_DIM + 1);
    const double params_host[1] = {P.reward_param};
    // zero-copy form: the pinned staging buffer as the device sees it (same memory; unified addressing normally makes
    // the two pointers equal).  If the buffer is not mapped, the copy + synchronise form below is used instead.
    double* h_flag = h_out + IPP_ROLLOUT_MAX_LEVELS + 8;
    double* h_rec = h_out + IPP_ROLLOUT_MAX_LEVELS + 16;                          // [L + 1] {value, tag} records (16-byte aligned)
    double* h_rec2 = h_rec + 2 * (IPP_ROLLOUT_MAX_LEVELS + 1);                    // hotspot: the information-gain pass's records
    double* m_out = nullptr;
    double* m_rec = nullptr;
    double* m_rec2 = nullptr;
    unsigned int* m_flag = nullptr;
    bool zero_copy = false;
    {
        void* dp = nullptr;
        if (host_reward && !value_reward) {
        } else if (cudaHostGetDevicePointer(&dp, stage_host, 0) == cudaSuccess && dp) {
            m_out = static_cast<double*>(dp) + (h_out - stage_host);
            m_flag = reinterpret_cast<unsigned int*>(static_cast<double*>(dp) + (h_flag - stage_host));
            m_rec = static_cast<double*>(dp) + (h_rec - stage_host);
            m_rec2 = static_cast<double*>(dp) + (h_rec2 - stage_host);
            zero_copy = true;
        } else {
            cudaGetLastError();                                                  // clear the sticky "not mapped" error
        }
    }
    double* m_vals = nullptr;                                                    // naive_value: the pinned value buffer as the device sees it
    if (value_reward && zero_copy) {
        void* dp = nullptr;
        if (cudaHostGetDevicePointer(&dp, prm->rff_vals_host, 0) == cudaSuccess && dp) m_vals = static_cast<double*>(dp);
        else cudaGetLastError();
    }
    static unsigned int seq_base = 0;                                            // completion words never repeat a value
    unsigned int seq = (seq_base += 1u << 20);
    if (!host_reward || value_reward) {
        *reinterpret_cast<volatile unsigned int*>(h_flag) = seq;
        memset(h_rec, 0, sizeof(double) * 4 * (IPP_ROLLOUT_MAX_LEVELS + 1));     // no stale record can carry a live tag
    }
    ipp::RolloutInline inl;

    int max_samp = 0;
    const int max_dense = path_limits(P, &max_samp);
    std::vector<double> tmp_path((size_t)max_samp * 3);
    std::vector<int> incomplete;                                                  // belief nodes with goals still to try
    bool overflow = false;
    // try the next goal of belief node b; true when it produced an action child
    auto expand_one = [&](int b) -> bool {
        const int gi = nodes[b].g_next++;
        double pose[3], goal[3];
        memcpy(pose, nodes[b].pose, sizeof(pose));
        memcpy(goal, &pool[nodes[b].goals + 3 * (size_t)gi], sizeof(goal));
        const int n = goal_path(P, pose, goal, tmp_path.data(), max_samp, max_dense);
        if (n < 0) overflow = true;                                               // more samples than the buffers hold
        if (n <= 0) return false;
        Node a;
        a.action = true;
        a.depth = nodes[b].depth;
        a.parent = b;
        a.goal = gi;
        a.k = n;
        a.data = pool.size();
        pool.insert(pool.end(), tmp_path.begin(), tmp_path.begin() + 3 * n);
        memcpy(a.pose, pose, sizeof(pose));
        memcpy(a.end, &tmp_path[3 * (size_t)(n - 1)], sizeof(double) * 3);
        nodes[b].children.push_back((int)nodes.size());
        nodes[b].has_children = true;
        nodes.push_back(a);
        return true;
    };
    // one deferred path (most recently visited node first); false when nothing is left to do
    auto complete_some = [&]() -> bool {
        while (!incomplete.empty()) {
            const int b = incomplete.back();
            if (nodes[b].g_next >= nodes[b].n_goals) { incomplete.pop_back(); continue; }
            expand_one(b);
            return true;
        }
        return false;
    };

    const bool belief_tree = P.tree_type == 1;
    const size_t z_zero = pool.size();                                            // BeliefTree: MLE observation of a belief without a GPy model = zeros
    pool.resize(pool.size() + IPP_APPEND_MAX_K, 0.0);
    PathSet ps;
    std::vector<std::pair<size_t, int>> acts;
    // BeliefTree.random_rollouts (mcts_library.py:499-525) from belief node b: uniformly random actions (np.random.randint,
    // whose upper bound excludes the last action, Q12) down to the horizon, each recorded as a level with a zero observation
    auto random_rollouts = [&](int b) {
        int cur_depth = nodes[b].depth;
        acts.clear();
        for (int ci : nodes[b].children) acts.push_back({nodes[ci].data, nodes[ci].k});
        while (cur_depth <= P.max_depth) {
            const int n_act = (int)acts.size();
            if (n_act <= 1) return;
            const int a = np_randint(gnp, n_act - 1);
            if (acts[a].second > IPP_APPEND_MAX_K) { overflow = true; return; }
            levels.push_back({acts[a].first, z_zero, acts[a].second, 2});
            double pose[3];
            memcpy(pose, &pool[acts[a].first + 3 * (size_t)(acts[a].second - 1)], sizeof(pose));
            ++cur_depth;
            if (cur_depth > P.max_depth) return;
            build_path_set(P, pose, ps);
            acts.clear();
            for (size_t i = 0; i < ps.goal.size(); ++i) {
                const size_t off = pool.size();
                const double* sp = &ps.samp[i * (size_t)ps.max_samp * 3];
                pool.insert(pool.end(), sp, sp + 3 * ps.count[i]);
                acts.push_back({off, ps.count[i]});
            }
        }
    };

    double walk_us = 0.0, device_us = 0.0, launch_us = 0.0;
    for (int it = 0; it < P.budget; ++it) {
        const double t_walk = now_us();
        // ---- leaf_helper: walk down, recording the (points, observation, appends) of every action step
        levels.clear();
        int cur = 0;
        while (true) {
            if (!nodes[cur].action) {
                if (nodes[cur].depth == P.max_depth) break;
                // build_action_children + get_next_child.  The reference builds all candidate paths of a node at once and
                // then takes the first unvisited child; here the paths are tried one goal at a time, in the same order, so a
                // fresh node costs one Dubins path on the walk instead of sixteen -- the rest are completed while the host
                // waits for the device (complete_some), and the choice is the same: every built child before the first
                // unvisited one has been visited, and UCT only runs once all goals have been tried.
                if (nodes[cur].n_goals < 0) {
                    nodes[cur].goals = pool.size();
                    pool.resize(pool.size() + (size_t)(P.frontier_size + 1) * 3);
                    nodes[cur].n_goals = frontier_goals(P, nodes[cur].pose, &pool[nodes[cur].goals]);
                    incomplete.push_back(cur);
                }
                int pick = -1;
                for (int ci : nodes[cur].children)
                    if (nodes[ci].nq == 0) { pick = ci; break; }
                while (pick < 0 && nodes[cur].g_next < nodes[cur].n_goals)
                    if (expand_one(cur)) pick = nodes[cur].children.back();
                if (pick < 0 && nodes[cur].children.empty()) break;                // no feasible action: python children None
                if (pick >= 0 && belief_tree) {
                    // BeliefTree.leaf_helper (:540-547): an unvisited action ends the tree walk; the rest of the horizon is a
                    // random rollout from THIS belief node (it needs the node's complete action set), the leaf is the action
                    while (nodes[cur].g_next < nodes[cur].n_goals) expand_one(cur);
                    random_rollouts(cur);
                    cur = pick;
                    break;
                }
                if (pick < 0) {
                    // DPW-UCT (Tree) or UCB1 (BeliefTree :624-632) over all children, exact-equality ties -> random.choice
                    const Node& b = nodes[cur];
                    const double e_d = 0.5 * (1.0 - (3.0 / (10.0 * (P.max_depth - b.depth))));
                    double top = 0.0;
                    ties.clear();
                    for (int ci : b.children) {
                        const Node& ch = nodes[ci];
                        const double v = belief_tree
                            ? ch.reward / (double)ch.nq + P.c * sqrt(2.0 * P.log_table[b.nq] / (double)ch.nq)
                            : ch.reward / (double)ch.nq + P.c * sqrt(pow((double)b.nq, e_d) / (double)ch.nq);
                        if (ties.empty() || v > top) { top = v; ties.clear(); ties.push_back(ci); }
                        else if (v == top) ties.push_back(ci);
                    }
                }
                if (pick < 0) pick = ties[py_choice(gpy, (int)ties.size())];
                cur = pick;
            } else {
                // _record_action: re-use a child (progressive widening) or simulate a new observation
                Node& a = nodes[cur];
                int child = -1;
                if (belief_tree) {
                    // BeliefTree (:549-562): every traversal of an action creates a new belief child with the MLE observation
                    if (a.k > IPP_APPEND_MAX_K) { *status = 1; return IPP_OK; }
                    Node bnode;
                    bnode.depth = a.depth + 1;
                    bnode.parent = cur;
                    bnode.k = a.k;
                    bnode.data = z_zero;
                    memcpy(bnode.pose, a.end, sizeof(double) * 3);
                    child = (int)nodes.size();
                    levels.push_back({a.data, z_zero, a.k, 2});
                    nodes.push_back(bnode);                                        // may move `a`: not used below
                    nodes[cur].children.push_back(child);
                    nodes[cur].has_children = true;
                    cur = child;
                    continue;
                }
                if (a.has_children) {
                    const double al = 3.0 / (10.0 * (P.max_depth - a.depth) - 3.0);
                    const int nchild = (int)a.children.size();
                    if (a.depth < P.max_depth - 1 && floor(pow((double)nchild, al)) == floor(pow((double)(nchild - 1), al))) {
                        py_choice(gpy, nchild);                                    // draw consumed, result unused (:408)
                        int64_t fewest = nodes[a.children[0]].nq;
                        for (int ci : a.children) if (nodes[ci].nq < fewest) fewest = nodes[ci].nq;
                        ties.clear();
                        for (int ci : a.children) if (nodes[ci].nq == fewest) ties.push_back(ci);
                        child = ties[py_choice(gpy, (int)ties.size())];
                        levels.push_back({a.data, nodes[child].data, a.k, 1});
                    }
                }
                if (child < 0) {
                    if (a.k > IPP_APPEND_MAX_K) { *status = 1; return IPP_OK; }
                    Node bnode;
                    bnode.depth = a.depth + 1;
                    bnode.parent = cur;
                    bnode.k = a.k;
                    bnode.data = pool.size();
                    memcpy(bnode.pose, a.end, sizeof(double) * 3);
                    const double* sc = P.prior_scale + (size_t)a.k * IPP_APPEND_MAX_K;
                    for (int i = 0; i < a.k; ++i) pool.push_back(np_gauss(gnp, &np) * sc[i]);
                    child = (int)nodes.size();
                    levels.push_back({a.data, bnode.data, a.k, 2});
                    nodes.push_back(bnode);                                        // may move `a`: not used below
                    nodes[cur].children.push_back(child);
                    nodes[cur].has_children = true;
                }
                cur = child;
            }
        }
        if (overflow) { *status = 1; return IPP_OK; }
        // ---- settle: all rewards of the rollout in one device pass
        const double t_dev = now_us();
        walk_us += t_dev - t_walk;
        double reward = 0.0;
        const int L = (int)levels.size();
        if (L > 0 && value_reward) {
            // naive_value (aq_library.py:369-413): share of the sampled functions whose value at the point lies within the
            // tolerance of that function's maximum; the functions are evaluated for all points of the rollout in one pass
            int M = 0;
            for (const Level& lv : levels) M += lv.k;
            if (M > IPP_ROLLOUT_MAX_POINTS) { *status = 1; return IPP_OK; }
            int row = 0;
            for (int l = 0; l < L; ++l) {
                const Level& a = levels[l];
                if (a.k > IPP_APPEND_MAX_K) { *status = 1; return IPP_OK; }
                for (int i = 0; i < a.k; ++i, ++row) {
                    h_in[(size_t)row * D] = pool[a.data + 3 * i];
                    h_in[(size_t)row * D + 1] = pool[a.data + 3 * i + 1];
                    if (D == 3) h_in[(size_t)row * D + 2] = P.time;
                }
            }
            const int S_ = P.n_naive;
            if (zero_copy && m_vals && M <= ipp::RO_INLINE_POINTS) {
                memcpy(inl.v, h_in, sizeof(double) * (size_t)M * D);
                ++seq;
                rff_points_kernel<<<1, 1024, 0, stream>>>(P.rff_W, P.rff_b, P.rff_theta, P.rff_features, S_, D, inl, M, m_vals,
                                                          m_flag, seq);
                IPP_LAUNCH_CHECK();
                const double t_spin = now_us();
                for (unsigned spins = 0; *reinterpret_cast<volatile unsigned int*>(h_flag) != seq; ++spins) {
                    if (complete_some()) continue;
                    if ((spins & 0xffff) == 0xffff && now_us() - t_spin > 5e6) {
                        IPP_CUDA(cudaStreamSynchronize(stream));
                        if (*reinterpret_cast<volatile unsigned int*>(h_flag) == seq) break;
                        ::ipp::set_error("%s: rollout %d never completed", __func__, it);
                        return IPP_E_CUDA;
                    }
                }
                __atomic_thread_fence(__ATOMIC_ACQUIRE);
            } else {
                double* d_max = P.rff_vals_dev + (size_t)S_ * IPP_ROLLOUT_MAX_POINTS;
                int64_t* d_arg = reinterpret_cast<int64_t*>(d_max + S_);
                IPP_CUDA(cudaMemcpyAsync(d_in, h_in, sizeof(double) * (size_t)M * D, cudaMemcpyHostToDevice, stream));
                const int rc = ipp_rff_eval_argmax(P.rff_W, P.rff_b, P.rff_theta, P.rff_features, S_, D, 0, d_in, M,
                                                   P.rff_vals_dev, d_max, d_arg, P.rff_work, stream_);
                if (rc != IPP_OK) return rc;
                IPP_CUDA(cudaMemcpyAsync(P.rff_vals_host, P.rff_vals_dev, sizeof(double) * (size_t)S_ * M, cudaMemcpyDeviceToHost, stream));
                IPP_CUDA(cudaStreamSynchronize(stream));
            }
            row = 0;
            for (int l = 0; l < L; ++l) {
                const int k = levels[l].k;
                double f[IPP_APPEND_MAX_K];
                for (int i = 0; i < k; ++i) f[i] = 0.0;
                for (int sidx = 0; sidx < S_; ++sidx)
                    for (int i = 0; i < k; ++i)
                        f[i] += (fabs(P.rff_vals_host[(size_t)sidx * M + row + i] - P.naive_maxes[sidx]) <= P.reward_param) ? 1.0 : 0.0;
                for (int i = 0; i < k; ++i) f[i] = f[i] / (double)S_;
                reward = reward + np_pairwise_sum(f, k);
                row += k;
            }
            evals += L;
        } else if (L > 0 && host_reward) {
            // naive (aq_library.py:339-367): share of the sampled maximisers within `radius` of each point, summed over the
            // level's points the way np.sum does (pairwise, np_pairwise_sum); no belief, no device
            for (int l = 0; l < L; ++l) {
                const Level& a = levels[l];
                double f[IPP_APPEND_MAX_K];
                if (a.k > IPP_APPEND_MAX_K) { *status = 1; return IPP_OK; }
                for (int i = 0; i < a.k; ++i) f[i] = 0.0;
                for (int sidx = 0; sidx < P.n_naive; ++sidx) {
                    const double lx = P.naive_locs[2 * sidx], ly = P.naive_locs[2 * sidx + 1];
                    for (int i = 0; i < a.k; ++i) {
                        const double dx = pool[a.data + 3 * i] - lx, dy = pool[a.data + 3 * i + 1] - ly;
                        const double dist = sqrt(dx * dx + dy * dy);
                        f[i] += (dist <= P.reward_param) ? 1.0 : 0.0;
                    }
                }
                for (int i = 0; i < a.k; ++i) f[i] = f[i] / (double)P.n_naive;
                reward = reward + np_pairwise_sum(f, a.k);
            }
            evals += L;
        } else if (L > 0) {
            int M = 0;
            for (const Level& lv : levels) M += lv.k;
            if (L > IPP_ROLLOUT_MAX_LEVELS || M > IPP_ROLLOUT_MAX_POINTS) { *status = 1; return IPP_OK; }
            if ((P.kind == IPP_REWARD_INFO || hotspot) && M > 32) { *status = 1; return IPP_OK; }     // one-launch form only (ipp_b200.h)
            int32_t offsets[IPP_ROLLOUT_MAX_LEVELS + 1], reps[IPP_ROLLOUT_MAX_LEVELS];
            const bool fast = zero_copy && M <= ipp::RO_INLINE_POINTS;
            double* in = fast ? inl.v : h_in;
            int row = 0;
            for (int l = 0; l < L; ++l) {
                const Level& a = levels[l];
                offsets[l] = row;
                reps[l] = levels[l].reps;
                for (int i = 0; i < a.k; ++i, ++row) {
                    in[(size_t)row * D] = pool[a.data + 3 * i];
                    in[(size_t)row * D + 1] = pool[a.data + 3 * i + 1];
                    if (D == 3) in[(size_t)row * D + 2] = P.time;
                    in[(size_t)M * D + row] = pool[levels[l].z + i];
                }
            }
            offsets[L] = M;
            int32_t info;
            if (hotspot) {
                // hotspot_info_UCB = info_gain + mean_UCB (aq_library.py:117-135): two one-launch passes over the same
                // points, back to back on the stream -- UCB on the regression belief, information gain on (I + vK)^-1 --
                // each publishing its own tagged records; the sum per level is formed here, in the order ig + ucb
                if (!fast || reinterpret_cast<uintptr_t>(m_rec) % 16 != 0) { *status = 1; return IPP_OK; }
                const double t_launch = now_us();
                static unsigned long long hs_tag_counter = 1ull << 62;              // disjoint from the single-pass tags
                const unsigned long long tag_u = ++hs_tag_counter, tag_i = ++hs_tag_counter;
                bool tagged_u = false, tagged_i = false;
                const double info_params[1] = {P.info_param};
                int rc = ipp::rollout_launch(kern, X, N, alpha, winv, ldw, noise_pred, diag_add, IPP_REWARD_UCB, params_host, 1,
                                             nullptr, 0, d_in, d_in + (size_t)M * D, &inl, offsets, reps, L, m_out,
                                             reinterpret_cast<int32_t*>(m_out + L), m_flag, seq, work, stream, m_rec, tag_u, &tagged_u);
                if (rc != IPP_OK) return rc;
                rc = ipp::rollout_launch(P.info_kern, X, N, alpha, P.info_winv, P.info_ldw, 0.0, 1.0, IPP_REWARD_INFO, info_params, 1,
                                         nullptr, 0, d_in, d_in + (size_t)M * D, &inl, offsets, reps, L, m_out,
                                         reinterpret_cast<int32_t*>(m_out + L), m_flag, seq, work, stream, m_rec2, tag_i, &tagged_i);
                if (rc != IPP_OK) return rc;
                if (!tagged_u || !tagged_i) {                                       // (the four-kernel form has no information-gain pass)
                    IPP_CUDA(cudaStreamSynchronize(stream));
                    *status = 1;
                    return IPP_OK;
                }
                const double t_spin = now_us();
                launch_us += t_spin - t_launch;
                auto done = [&]() {
                    const volatile unsigned long long* tu = reinterpret_cast<const volatile unsigned long long*>(h_rec);
                    const volatile unsigned long long* ti = reinterpret_cast<const volatile unsigned long long*>(h_rec2);
                    for (int l = L; l >= 0; --l)
                        if (ti[2 * l + 1] != tag_i || tu[2 * l + 1] != tag_u) return false;
                    return true;
                };
                for (unsigned spins = 0; !done(); ++spins) {
                    if (complete_some()) continue;
                    if ((spins & 0xffff) == 0xffff && now_us() - t_spin > 5e6) {
                        IPP_CUDA(cudaStreamSynchronize(stream));
                        if (done()) break;
                        ::ipp::set_error("%s: rollout %d never completed", __func__, it);
                        return IPP_E_CUDA;
                    }
                }
                __atomic_thread_fence(__ATOMIC_ACQUIRE);
                for (int l = 0; l < L; ++l) h_out[l] = h_rec2[2 * l] + h_rec[2 * l];
                const int32_t bad = ((int32_t)h_rec[2 * L] != 0 || (int32_t)h_rec2[2 * L] != 0) ? 1 : 0;
                memcpy(h_out + L, &bad, sizeof(int32_t));
            } else if (fast) {
                // points by value in the kernel parameters, rewards + info written straight into the pinned buffer, completion
                // word polled by this thread: no copies and no driver synchronisation on the rollout's critical path
                ++seq;
                const double t_launch = now_us();
                bool tagged = false;
                static unsigned long long tag_counter = 0;                          // 64 bits: a tag never repeats
                const unsigned long long tag64 = ++tag_counter;
                const int rc = ipp::rollout_launch(kern, X, N, alpha, winv, ldw, noise_pred, diag_add, P.kind, params_host, 1,
                                                   maxes_dev, S, d_in, d_in + (size_t)M * D, &inl, offsets, reps, L, m_out,
                                                   reinterpret_cast<int32_t*>(m_out + L), m_flag, seq, work, stream,
                                                   (reinterpret_cast<uintptr_t>(m_rec) % 16 == 0) ? m_rec : nullptr, tag64, &tagged);
                if (rc != IPP_OK) return rc;
                const double t_spin = now_us();
                launch_us += t_spin - t_launch;
                // done = the completion word (four-kernel form) or, for the one-launch form, every {value, tag} record
                // carrying this rollout's tag: each record is one 16-byte store, so a visible tag means a visible value
                auto done = [&]() {
                    if (!tagged) return *reinterpret_cast<volatile unsigned int*>(h_flag) == seq;
                    const volatile unsigned long long* tags = reinterpret_cast<const volatile unsigned long long*>(h_rec);
                    for (int l = L; l >= 0; --l)
                        if (tags[2 * l + 1] != tag64) return false;
                    return true;
                };
                for (unsigned spins = 0; !done(); ++spins) {
                    if (complete_some()) continue;                                  // useful work while the device runs
                    if ((spins & 0xffff) == 0xffff && now_us() - t_spin > 5e6) {          // 5 s: something is wrong on the device
                        IPP_CUDA(cudaStreamSynchronize(stream));
                        if (done()) break;
                        ::ipp::set_error("%s: rollout %d never completed", __func__, it);
                        return IPP_E_CUDA;
                    }
                }
                __atomic_thread_fence(__ATOMIC_ACQUIRE);
                if (tagged) {                                                       // unpack into the plain layout read below
                    for (int l = 0; l < L; ++l) h_out[l] = h_rec[2 * l];
                    const int32_t bad = (int32_t)h_rec[2 * L];
                    memcpy(h_out + L, &bad, sizeof(int32_t));
                }
            } else {
                IPP_CUDA(cudaMemcpyAsync(d_in, h_in, sizeof(double) * (size_t)M * (D + 1), cudaMemcpyHostToDevice, stream));
                const int rc = ipp::rollout_launch(kern, X, N, alpha, winv, ldw, noise_pred, diag_add, P.kind, params_host, 1,
                                                   maxes_dev, S, d_in, d_in + (size_t)M * D, nullptr, offsets, reps, L, d_out,
                                                   reinterpret_cast<int32_t*>(d_out + L), nullptr, 0u, work, stream, nullptr, 0ull, nullptr);
                if (rc != IPP_OK) return rc;
                IPP_CUDA(cudaMemcpyAsync(h_out, d_out, sizeof(double) * (size_t)(L + 1), cudaMemcpyDeviceToHost, stream));
                IPP_CUDA(cudaStreamSynchronize(stream));
            }
            memcpy(&info, h_out + L, sizeof(int32_t));
            if (info != 0) { *status = 1; return IPP_OK; }
            for (int l = 0; l < L; ++l) reward = reward + h_out[l];
            evals += L;
        }
        device_us += now_us() - t_dev;
        // ---- backprop
        for (int n = cur; n >= 0; n = nodes[n].parent) {
            nodes[n].nq += 1;
            nodes[n].reward += reward;
        }
    }

    const Node& root = nodes[0];
    *n_root = (int32_t)root.children.size();
    for (size_t i = 0; i < root.children.size(); ++i) {
        const Node& ch = nodes[root.children[i]];
        root_goal[i] = ch.goal;
        root_visits[i] = ch.nq;
        root_reward[i] = ch.reward;
    }
    *reward_evals = evals;
    g_walk_us = walk_us;
    g_device_us = device_us;
    g_launch_us = launch_us;
    *py_rng = py;
    *np_rng = np;
    return IPP_OK;
}
