// Dubins candidate paths.  (1) Host-side (CPU, no kernel) path-set generation for one pose; (2) the same generator as a
// CUDA kernel for large batches of independent (pose, goal) pairs (BASELINE config 5: 64k candidate paths), one
// thread per path, sharing the closed forms below through __host__ __device__ functions.
// (1): the native counterpart of the `dubins` C library
// the reference calls from paths_library.py:141-175 (dubins.shortest_path + path.sample_many), plus the
// border / obstacle trimming of Dubins_Path_Generator.buffered_paths (obstacle worlds as box-with-hole lists).
// Classical six-word closed form (LSL, RSR, LSR, RSL, RLR, LRL; Shkel & Lumelsky 2001).
#include <math.h>
#include "internal.cuh"
#include "dubins.cuh"

namespace {

using namespace ipp_dubins;

__global__ void dubins_paths_kernel(const double* __restrict__ poses, const double* __restrict__ goals, int64_t P,
                                    double rho, double dense_step, int keep_every, double e0, double e1, double e2,
                                    double e3, double border, const __grid_constant__ ObstacleSet obs,
                                    double* __restrict__ samp, int32_t* __restrict__ counts,
                                    int max_samp, double* __restrict__ end_pose) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const double extent[4] = {e0, e1, e2, e3};
    const double pose[3] = {poses[3 * p], poses[3 * p + 1], poses[3 * p + 2]};
    const double goal[3] = {goals[3 * p], goals[3 * p + 1], goals[3 * p + 2]};
    int dense_n = 0;
    double end[3] = {pose[0], pose[1], pose[2]};
    const int n = dubins_one(pose, goal, rho, dense_step, keep_every, extent, border, obs.box, obs.n,
                             samp + p * max_samp * 3, max_samp, nullptr, 1 << 30, &dense_n, end);
    counts[p] = n;
    if (end_pose) { end_pose[3 * p] = end[0]; end_pose[3 * p + 1] = end[1]; end_pose[3 * p + 2] = end[2]; }
}

}  // namespace

// For each goal g: dense[g][0..dense_count[g]) = configurations every `dense_step` along the shortest Dubins
// path from `pose`, cut at the first one outside the open extent; samp = every `keep_every`-th of them, trimmed
// where a kept point lies within `border` of the extent exactly as the reference's loop does (including its
// `ttemp[0:m-1]` slicing quirk).  samp_count[g] = 0 when fewer than 2 points survive (path dropped, :167-171);
// dense_count[g] then is meaningless.  Arrays are host memory: samp [G][max_samp][3], dense [G][max_dense][3].
extern "C" int ipp_dubins_path_set(const double* pose, const double* goals, int G, double rho, double dense_step,
                                   int keep_every, const double* extent, double border,
                                   const ipp_obstacle_box* obstacles, int n_obstacles,
                                   double* samp, int* samp_count, int max_samp,
                                   double* dense, int* dense_count, int max_dense) {
    IPP_CHECK_ARG(pose && goals && extent && samp && samp_count && dense_count, "null pointer");
    IPP_CHECK_ARG(G >= 0 && rho > 0 && dense_step > 0 && keep_every >= 1 && max_samp >= 2 && max_dense >= 2, "bad argument");
    IPP_CHECK_ARG(n_obstacles >= 0 && (n_obstacles == 0 || obstacles), "bad obstacle list");
    for (int gi = 0; gi < G; ++gi) {
        int dn = 0;
        samp_count[gi] = dubins_one(pose, goals + 3 * gi, rho, dense_step, keep_every, extent, border,
                                    obstacles, n_obstacles, samp + (size_t)gi * max_samp * 3, max_samp,
                                    dense ? dense + (size_t)gi * max_dense * 3 : nullptr, max_dense, &dn, nullptr);
        dense_count[gi] = dn;
    }
    return IPP_OK;
}

// ---- 8f rank 2: the same generator on the device, one thread per (pose, goal) pair.
extern "C" int ipp_dubins_paths_device(const double* poses, const double* goals, int64_t P, double rho, double dense_step,
                                       int keep_every, const double* extent_host, double border,
                                       const ipp_obstacle_box* obstacles_host, int n_obstacles,
                                       double* samp, int32_t* counts, int max_samp, double* end_pose, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    IPP_CHECK_ARG(poses && goals && extent_host && samp && counts, "null pointer");
    IPP_CHECK_ARG(P >= 0 && rho > 0 && dense_step > 0 && keep_every >= 1 && max_samp >= 2, "bad argument");
    IPP_CHECK_ARG(n_obstacles >= 0 && n_obstacles <= IPP_MAX_OBSTACLES && (n_obstacles == 0 || obstacles_host),
                  "bad obstacle list (at most IPP_MAX_OBSTACLES boxes)");
    if (P == 0) return IPP_OK;
    ObstacleSet obs;
    obs.n = n_obstacles;
    for (int i = 0; i < n_obstacles; ++i) obs.box[i] = obstacles_host[i];
    dubins_paths_kernel<<<(unsigned)((P + 127) / 128), 128, 0, s>>>(poses, goals, P, rho, dense_step, keep_every,
                                                                   extent_host[0], extent_host[1], extent_host[2],
                                                                   extent_host[3], border, obs, samp, counts, max_samp, end_pose);
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}
