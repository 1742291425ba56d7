// a12 batching seam: every reward of ONE tree-search rollout in a single call.
//
// Reference flow (mcts_library.py:354-442 Tree.leaf_helper, :499-621 BeliefTree): at level l the rollout
// scores the k_l sample points of its action under the belief = base belief + the simulated observations
// of levels < l, then appends its own observation block once (re-used child, :413) or twice (new child,
// :426 & :430).  Each append is an O(N^2 k) block-inverse update of the N x N matrix and each reward an
// O(N^2 k) predict -- although the rollout only ever looks at <= ~40 points.
//
// Here the N x N work is done ONCE per rollout against the (immutable) base belief:
//   m0 = K(P, X) alpha,   C0 = K(P, P) - K(P, X) W^-1 K(X, P)          P = all sample points of the rollout
// and the per-level beliefs are obtained by conditioning the small joint Gaussian (m, C) over P on the noisy
// block observations, which is the same posterior the reference reaches through update_model
// (gpmodel_library.py:230-279: Schur block M = (S - R P^-1 Q + (noise + 1e-8) I)^-1 is exactly
// (C[B,B] + (noise + 1e-8) I)^-1 here):
//   A = C[B,B] + (diag_add / reps[l]) I = L L^T;  H = C[:,B] L^-T;  u = L^-1 (y_B - m_B);  m += H u;  C -= H H^T
// (the same block appended reps[l] times = one conditioning with the noise divided by reps[l]).  One CTA, everything
// in shared memory.
#include <stdlib.h>
#include <string.h>

#include "internal.cuh"
#include "dgemm_dmma.cuh"

namespace ipp {

// shared with rewards.cu (same closed forms; kept textually identical so both paths agree bit for bit)
__device__ __forceinline__ double ro_norm_cdf(double a) {
    const double x = a * 0.70710678118654752440;
    const double z = fabs(x);
    if (z < 0.70710678118654752440) return 0.5 + 0.5 * erf(x);
    double y = 0.5 * erfc(z);
    if (x > 0) y = 1.0 - y;
    return y;
}
__device__ __forceinline__ double ro_norm_pdf(double x) { return exp(-(x * x) / 2.0) / 2.50662827463100050242; }
__device__ __forceinline__ double ro_mes_utility(double ymax, double mu, double var) {
    const double gamma = (ymax - mu) / var;
    const double pdf = ro_norm_pdf(gamma), cdf = ro_norm_cdf(gamma);
    return gamma * pdf / (2.0 * cdf) - log(cdf);
}

constexpr int RO_MAX_POINTS = IPP_ROLLOUT_MAX_POINTS;
constexpr int RO_MAX_BLOCK = IPP_APPEND_MAX_K;
constexpr int RO_MAX_LEVELS = IPP_ROLLOUT_MAX_LEVELS;
constexpr int RO_THREADS = 256;

struct RolloutArgs {
    const double* mean0;        // [M]
    const double* cov0;         // [M][ldc]
    int64_t ldc;
    const double* zobs;         // [M]
    int M, L;
    int offsets[RO_MAX_LEVELS + 1];
    int reps[RO_MAX_LEVELS];
    int kind;
    double p0;                  // UCB: sqrt(beta); EI: eta
    const double* maxes; int S; // MES
    double noise_pred;          // added to the predictive variance (predict_value include_noise=True)
    double diag_add;            // noise + 1e-8 on the Schur block
    double* reward;             // [L]
    int32_t* info;              // 0 ok; l + 1: level l's block was not positive definite
    unsigned int* done;         // optional (host-mapped): set to done_seq after reward / info are visible system-wide
    unsigned int done_seq;
};

// Kernels 2-4 of a rollout are launched with programmatic stream serialisation: their blocks are scheduled while the
// previous kernel still runs and wait here until it has completed and its writes are visible (a no-op otherwise).
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Level-by-level rewards and conditioning of the small joint Gaussian (m, C) over the rollout's M points, by one CTA of
// RO_THREADS threads.  C [M][M+1], m [M], H [(M+1)][RO_MAX_BLOCK+1], red [2 * RO_THREADS / 32] live in shared memory and
// are already filled (C, m) when this is called; zobs may be shared or global.  Ends by publishing rewards, info and the
// completion word.
__device__ __forceinline__ void ro_condition(const RolloutArgs& p, const double* zobs, double* C, double* m, double* H,
                                             double* red, int& s_bad) {
    const int M = p.M, ldc = M + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int HL = RO_MAX_BLOCK + 1;
    if (tid == 0) s_bad = 0;
    __syncthreads();

    for (int l = 0; l < p.L; ++l) {
        const int b0 = p.offsets[l], kb = p.offsets[l + 1] - b0;
        // ---- reward of level l from (m, diag C + noise) of its points
        double s0 = 0.0, s1 = 0.0;
        if (p.kind == IPP_REWARD_MES) {
            for (int idx = tid; idx < kb * p.S; idx += RO_THREADS) {
                const int q = b0 + idx / p.S;
                s0 += ro_mes_utility(p.maxes[idx % p.S], m[q], C[q * ldc + q] + p.noise_pred);
            }
        } else {
            for (int q = b0 + tid; q < b0 + kb; q += RO_THREADS) {
                s0 += (p.kind == IPP_REWARD_EI) ? (m[q] - p.p0) : m[q];
                s1 += fabs(C[q * ldc + q] + p.noise_pred);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        if (lane == 0) { red[2 * warp] = s0; red[2 * warp + 1] = s1; }
        __syncthreads();
        if (tid == 0) {
            s0 = s1 = 0.0;
            for (int w = 0; w < RO_THREADS / 32; ++w) { s0 += red[2 * w]; s1 += red[2 * w + 1]; }
            double r;
            if (p.kind == IPP_REWARD_UCB) {
                r = s0 + p.p0 * s1;
            } else if (p.kind == IPP_REWARD_EI) {
                const double z = s0 / s1;
                const double big_phi = 0.5 * (1.0 + erf(z / 1.41421356237309504880));
                const double small_phi = 1.0 / 2.50662827463100050242 * exp(-(z * z) / 2.0);
                r = s0 * big_phi + s1 * small_phi;
            } else {
                r = 2.0 * s0 / (double)p.S;
            }
            p.reward[l] = r;
        }
        if (l == p.L - 1) break;                  // the last level's observation is never looked at again
        // ---- condition (m, C) on the block observation; only rows >= b0 are read later.  Appending the SAME noisy
        // block r times (the reference's double add_data, Q4) is one conditioning on that observation with the noise
        // divided by r: r independent readings that all equal y are one reading of y with variance diag_add / r.
        if (p.reps[l] > 0 && kb <= 4) {
            // Blocks of up to 4 points (the usual candidate path: horizon / sample step = 3-4 samples): EVERY thread
            // factors the 4 x 4 block redundantly in registers (identity-padded, fully unrolled) and substitutes its own
            // row and the residual -- two barrier-separated phases per level instead of 3 kb + 3.  ncu of the generic sweep
            // below at M = 20: 20 us, barrier stalls 6.5 per issued instruction (profiles/r1j_ncu_full_rollout_condition.csv).
            const double dadd = p.diag_add / (double)p.reps[l];
            const int R = M - b0;
            __syncthreads();
            double Lf[4][4], rd[4], u[4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b <= a; ++b)
                    Lf[a][b] = (a < kb) ? C[(b0 + a) * ldc + b0 + b] + (a == b ? dadd : 0.0) : (a == b ? 1.0 : 0.0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double d = Lf[j][j];
                if (!(d > 0.0)) { if (tid == 0) s_bad = l + 1; d = 1.0; }
                const double r = rsqrt(d);
                rd[j] = r;
#pragma unroll
                for (int i = j + 1; i < 4; ++i) Lf[i][j] *= r;
#pragma unroll
                for (int i = j + 1; i < 4; ++i)
#pragma unroll
                    for (int c = j + 1; c <= i; ++c) Lf[i][c] = fma(-Lf[i][j], Lf[c][j], Lf[i][c]);
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) {                     // u = L^-1 (y_B - m_B), in every thread
                double v = (a < kb) ? zobs[b0 + a] - m[b0 + a] : 0.0;
#pragma unroll
                for (int c = 0; c < a; ++c) v = fma(-Lf[a][c], u[c], v);
                u[a] = v * rd[a];
            }
            const int i = b0 + tid;                           // this thread's point (M <= 128 < RO_THREADS: at most one)
            double dm = 0.0;
            if (i < M) {
                double h[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    double v = (a < kb) ? C[i * ldc + b0 + a] : 0.0;
#pragma unroll
                    for (int c = 0; c < a; ++c) v = fma(-Lf[a][c], h[c], v);
                    h[a] = v * rd[a];
                    H[i * HL + a] = h[a];
                    dm = fma(h[a], u[a], dm);
                }
            }
            __syncthreads();                                  // H complete; everybody is done reading m[B] and C[:, B]
            if (i < M) m[i] += dm;
            for (int idx = tid; idx < R * R; idx += RO_THREADS) {
                const int ii = b0 + idx / R, jj = b0 + idx % R;
                double sacc = 0.0;
#pragma unroll
                for (int a = 0; a < 4; ++a) sacc = fma(H[ii * HL + a], H[jj * HL + a], sacc);
                C[ii * ldc + jj] -= sacc;
            }
        } else if (p.reps[l] > 0) {
            // Right-looking elimination of the panel [C[b0:, B] + dadd I ; (y_B - m_B)^T] by all threads: after column a,
            // H[i][a] = (L^-1 C[B, i])_a for every remaining point i and H[M][a] = u_a = (L^-1 (y_B - m_B))_a -- the
            // Cholesky factor of the block, the forward substitutions of every row and of the residual in one sweep
            // (kb steps of rsqrt + scale + rank-1 update instead of three dependent triangular loops).
            const double dadd = p.diag_add / (double)p.reps[l];
            const int R = M - b0;
            __syncthreads();
            for (int idx = tid; idx < R * kb; idx += RO_THREADS) {
                const int i = idx / kb, a = idx % kb;
                H[(b0 + i) * HL + a] = C[(b0 + i) * ldc + b0 + a] + (i == a ? dadd : 0.0);
            }
            if (tid < kb) H[M * HL + tid] = zobs[b0 + tid] - m[b0 + tid];
            __syncthreads();
            for (int a = 0; a < kb; ++a) {
                const double piv = H[(b0 + a) * HL + a];
                __syncthreads();
                if (!(piv > 0.0) && tid == 0) s_bad = l + 1;
                const double d = piv > 0.0 ? rsqrt(piv) : 1.0;
                for (int i = b0 + tid; i <= M; i += RO_THREADS)
                    H[i * HL + a] = (i < b0 + a) ? 0.0 : H[i * HL + a] * d;
                __syncthreads();
                const int rest = kb - a - 1;
                for (int idx = tid; idx < (R + 1) * rest; idx += RO_THREADS) {
                    const int i = b0 + idx / rest, c = a + 1 + idx % rest;
                    H[i * HL + c] = fma(-H[i * HL + a], H[(b0 + c) * HL + a], H[i * HL + c]);
                }
                __syncthreads();
            }
            for (int i = b0 + tid; i < M; i += RO_THREADS) {      // m += H u
                double dm = 0.0;
                for (int a = 0; a < kb; ++a) dm = fma(H[i * HL + a], H[M * HL + a], dm);
                m[i] += dm;
            }
            for (int idx = tid; idx < R * R; idx += RO_THREADS) {  // C -= H H^T
                const int i = b0 + idx / R, j = b0 + idx % R;
                double s = 0.0;
                for (int a = 0; a < kb; ++a) s = fma(H[i * HL + a], H[j * HL + a], s);
                C[i * ldc + j] -= s;
            }
        }
        __syncthreads();
    }
    __syncthreads();
    if (tid == 0) {
        *p.info = s_bad;
        if (p.done) {
            __threadfence_system();
            *reinterpret_cast<volatile unsigned int*>(p.done) = p.done_seq;
        }
    }
}

__global__ void __launch_bounds__(RO_THREADS, 1) rollout_condition_kernel(RolloutArgs p) {
    extern __shared__ __align__(16) double sm[];
    grid_dependency_wait();
    const int M = p.M, ldc = M + 1;
    double* C = sm;                              // [M][M+1]
    double* m = C + (size_t)M * ldc;             // [M]
    double* H = m + M;                           // [M + 1][RO_MAX_BLOCK + 1]; row M = the block's residual y_B - m_B
    double* red = H + (size_t)(M + 1) * (RO_MAX_BLOCK + 1);   // [RO_THREADS / 32][2]
    __shared__ int s_bad;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < M * M; idx += RO_THREADS) C[(idx / M) * ldc + idx % M] = p.cov0[(int64_t)(idx / M) * p.ldc + idx % M];
    for (int i = tid; i < M; i += RO_THREADS) m[i] = p.mean0[i];
    ro_condition(p, p.zobs, C, m, H, red, s_bad);
}

// ---- base posterior over the rollout's M <= 128 points.  A 128-row DMMA tile would be > 80 % padding here and only
// N / 128 CTAs would stream the N x N operand (measured 130 us per GEMM at N = 900), so the two contractions are
// bandwidth-shaped SIMT kernels over all SMs instead:
//   T1[m][n] = sum_j W[n][j] Kxt[m][j]          two rows n of W^-1 per warp (the only N^2 traffic), <= 32 m per pass
//   C0[m][m'] = Kqq[m][m'] - sum_n T1[m][n] Kxt[m'][n]   one warp per (m, m' <= m), mirrored
// Cross-covariance rows of the rollout's M <= 128 points: thread = observation n (coalesced stores along n), loop over
// the points; the general query kernel maps 4 queries to a warp and would keep 5 warps busy for M = 20.
template <int D>
__global__ void __launch_bounds__(256) rollout_rbf_kernel(RbfDev k, const double* __restrict__ X, int N,
                                                          const double* __restrict__ pts, int M,
                                                          double* __restrict__ Kxt, int64_t ldk) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double x[D];
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] = X[(int64_t)n * D + d];
    for (int m = 0; m < M; ++m) {
        double q[D];
#pragma unroll
        for (int d = 0; d < D; ++d) q[d] = pts[m * D + d];
        Kxt[(int64_t)m * ldk + n] = rbf_eval<D>(k, q, x);
    }
}

// Same, with the rollout's points and observations arriving BY VALUE in the kernel parameters (no host-to-device copy
// on the critical path of a rollout): block 0 also stores them to device scratch for the kernels that follow.
template <int D>
__global__ void __launch_bounds__(256) rollout_rbf_inline_kernel(RbfDev k, const double* __restrict__ X, int N,
                                                                 const __grid_constant__ RolloutInline pp, int M,
                                                                 double* __restrict__ pts_out, double* __restrict__ zobs_out,
                                                                 double* __restrict__ Kxt, int64_t ldk) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < M * D; i += blockDim.x) pts_out[i] = pp.v[i];
        for (int i = threadIdx.x; i < M; i += blockDim.x) zobs_out[i] = pp.v[M * D + i];
    }
    if (n >= N) return;
    double x[D];
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] = X[(int64_t)n * D + d];
    for (int m = 0; m < M; ++m) {
        double q[D];
#pragma unroll
        for (int d = 0; d < D; ++d) q[d] = pp.v[m * D + d];
        Kxt[(int64_t)m * ldk + n] = rbf_eval<D>(k, q, x);
    }
}

constexpr int TW_ROWS_PER_WARP = 4, TW_JC = 256, TW_WARPS = 8, TW_PTS = 16;
constexpr int TW_ROWS_PER_BLOCK = (TW_WARPS / 2) * TW_ROWS_PER_WARP;       // two warps (point halves) share a row group
constexpr int TW_SMEM = 2 * 32 * TW_JC * 8;                 // two tile buffers of [32 points][256 j] doubles = 128 KB

__device__ __forceinline__ void tw_cp16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}

__global__ void __launch_bounds__(32 * TW_WARPS, 1) rollout_tw_kernel(const double* __restrict__ W, int64_t ldw, int N,
                                                                      const double* __restrict__ Kxt, int64_t ldk, int M,
                                                                      double* __restrict__ T1) {
    // Block = 16 rows of W^-1; the cross-covariance rows of <= 32 points are staged in shared memory in j-chunks shared
    // by all 16 rows (at N = 4096 the M x N panel does not fit L1: the untiled version re-streamed it from L2 once per
    // row -- 2.7 GB per call, 1.8 ms).  The chunks arrive by cp.async, double buffered: a plain load-store loop
    // serialised ~20 L2 latencies per chunk and kept the kernel at 0.45 ms.
    // Warp = 4 rows x one HALF of the points: every staged value read from shared memory feeds 4 FMAs (the 2-row x
    // all-points layout fed 2 and was bound by shared-memory bandwidth, not by the FP64 pipe or HBM).
    extern __shared__ __align__(16) double tw_smem[];
    const uint32_t s_base = static_cast<uint32_t>(__cvta_generic_to_shared(tw_smem));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half_id = warp & 1;
    const int n0 = (blockIdx.x * (TW_WARPS / 2) + (warp >> 1)) * TW_ROWS_PER_WARP;
    for (int m0 = 0; m0 < M; m0 += 32) {
        const int mm = (M - m0) < 32 ? (M - m0) : 32;
        const int half = (mm + 1) >> 1;                                   // points of this warp: [i0, i1)
        const int i0 = half_id * half, i1 = (i0 + half) < mm ? (i0 + half) : mm;
        auto load_tile = [&](int buf, int j0) {
            for (int idx = threadIdx.x; idx < mm * (TW_JC / 2); idx += 32 * TW_WARPS) {
                const int m = idx / (TW_JC / 2), jj = 2 * (idx % (TW_JC / 2));
                const int left = N - (j0 + jj);
                const int bytes = left >= 2 ? 16 : (left == 1 ? 8 : 0);             // zero-fill beyond the N valid columns
                const double* src = Kxt + (int64_t)(m0 + m) * ldk + (bytes ? j0 + jj : 0);
                tw_cp16(s_base + (uint32_t)(((buf * 32 + m) * TW_JC + jj) * 8), src, bytes);
            }
            asm volatile("cp.async.commit_group;\n" ::);
        };
        double acc[TW_ROWS_PER_WARP][TW_PTS];
#pragma unroll
        for (int r = 0; r < TW_ROWS_PER_WARP; ++r)
#pragma unroll
            for (int i = 0; i < TW_PTS; ++i) acc[r][i] = 0.0;
        int buf = 0;
        load_tile(0, 0);
        for (int j0 = 0; j0 < N; j0 += TW_JC) {
            // this chunk's rows of W^-1 first (independent of the tile), then wait for the tile
            double w[TW_JC / 32][TW_ROWS_PER_WARP];
#pragma unroll
            for (int c = 0; c < TW_JC / 32; ++c) {
                const int j = j0 + lane + 32 * c;
#pragma unroll
                for (int r = 0; r < TW_ROWS_PER_WARP; ++r)
                    w[c][r] = (j < N && n0 + r < N) ? W[(int64_t)(n0 + r) * ldw + j] : 0.0;
            }
            if (j0 + TW_JC < N) {
                load_tile(buf ^ 1, j0 + TW_JC);
                asm volatile("cp.async.wait_group 1;\n" ::);
            } else {
                asm volatile("cp.async.wait_group 0;\n" ::);
            }
            __syncthreads();
            const double* sK = tw_smem + (size_t)buf * 32 * TW_JC + (size_t)i0 * TW_JC;
#pragma unroll
            for (int c = 0; c < TW_JC / 32; ++c) {
                const int jj = lane + 32 * c;
#pragma unroll
                for (int i = 0; i < TW_PTS; ++i) {
                    if (i0 + i < i1) {
                        const double kv = sK[i * TW_JC + jj];
#pragma unroll
                        for (int r = 0; r < TW_ROWS_PER_WARP; ++r) acc[r][i] = fma(w[c][r], kv, acc[r][i]);
                    }
                }
            }
            __syncthreads();
            buf ^= 1;
        }
#pragma unroll
        for (int r = 0; r < TW_ROWS_PER_WARP; ++r)
#pragma unroll
            for (int i = 0; i < TW_PTS; ++i) {
                if (i0 + i < i1) {
                    double t = acc[r][i];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                    if (lane == 0 && n0 + r < N) T1[(int64_t)(m0 + i0 + i) * ldk + n0 + r] = t;
                }
            }
        __syncthreads();
    }
}

// ---- the same pass on the FP64 tensor pipe.  ncu of the SIMT version at N = 4096 (profiles/r1j_ncu_full_rollout_tw.csv):
// 90 us, FP64 pipe 41 % busy with 8 warps per SM, DRAM at 19 % -- bound by FP64 FMA issue, not by the 134 MB of W^-1.
// DMMA.8x8x4 does the same multiply-adds at the tensor rate and takes BOTH operands straight from global memory:
//   block = 16 rows of W^-1 (2 DMMA row tiles), 8 warps = 8 contiguous slices of the j range (split-K inside the block)
//   warp, per 32 j: A = W^-1[16 rows][32 j], B = Kxt[<= 32 points][32 j], 8 k-steps x (2 x 4) DMMA tiles
//   k permutation: thread t (= lane % 4) of a quad supplies j = j0 + 8 t + s in k-step s for A and for B alike, so
//   every thread reads 8 CONTIGUOUS doubles per row / point (four 16-byte loads) and nothing is staged in shared memory
// The 8 partial 16 x 32 tiles are summed through shared memory in a fixed order.
constexpr int TD_ROWS = 16, TD_WARPS = 8, TD_PTS = 32, TD_PREFETCH = 128;

__device__ __forceinline__ void td_load8(double (&dst)[8], const double* __restrict__ row, int j, int N) {
    if (j + 8 <= N) {
        const double2* p = reinterpret_cast<const double2*>(row + j);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double2 v = __ldg(p + q);
            dst[2 * q] = v.x;
            dst[2 * q + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) dst[q] = (j + q < N) ? __ldg(row + j + q) : 0.0;     // pad columns are not initialised
    }
}

__global__ void __launch_bounds__(32 * TD_WARPS, 2) rollout_tw_dmma_kernel(const double* __restrict__ W, int64_t ldw, int N,
                                                                           const double* __restrict__ Kxt, int64_t ldk,
                                                                           int M, double* __restrict__ T1) {
    __shared__ double part[TD_WARPS][TD_ROWS][TD_PTS + 1];
    grid_dependency_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int n0 = blockIdx.x * TD_ROWS;
    const int slice = (((N + TD_WARPS - 1) / TD_WARPS) + 31) / 32 * 32;
    const int j_lo = warp * slice, j_hi = (j_lo + slice) < N ? (j_lo + slice) : N;
    const int r0 = (n0 + g) < N ? (n0 + g) : (N - 1), r1 = (n0 + 8 + g) < N ? (n0 + 8 + g) : (N - 1);
    const double* w0 = W + (int64_t)r0 * ldw;
    const double* w1 = W + (int64_t)r1 * ldw;
    for (int m0 = 0; m0 < M; m0 += TD_PTS) {
        const int mm = (M - m0) < TD_PTS ? (M - m0) : TD_PTS;
        const int ntiles = (mm + 7) >> 3;
        const double* kx[4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int pnt = m0 + nt * 8 + g;
            kx[nt] = Kxt + (int64_t)(pnt < M ? pnt : (M - 1)) * ldk;
        }
        double c[2][4][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) c[mt][nt][0] = c[mt][nt][1] = 0.0;
        for (int j0 = j_lo; j0 < j_hi; j0 += 32) {
            const int j = j0 + 8 * t;
            double a0[8], a1[8];
            td_load8(a0, w0, j, N);
            td_load8(a1, w1, j, N);
            if (j0 + TD_PREFETCH < j_hi) {        // W^-1 is the only HBM stream: pull the lines of a later chunk into L2 now
                asm volatile("prefetch.global.L2 [%0];" ::"l"(w0 + j + TD_PREFETCH));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(w1 + j + TD_PREFETCH));
            }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
                if (nt < ntiles) {
                    double b[8];
                    td_load8(b, kx[nt], j, N);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        dmma884(c[0][nt][0], c[0][nt][1], a0[ks], b[ks]);
                        dmma884(c[1][nt][0], c[1][nt][1], a1[ks], b[ks]);
                    }
                }
        }
        // C fragment: row g (+ 8 mt), columns (points) 2 t, 2 t + 1 of tile nt
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                part[warp][mt * 8 + g][nt * 8 + 2 * t] = c[mt][nt][0];
                part[warp][mt * 8 + g][nt * 8 + 2 * t + 1] = c[mt][nt][1];
            }
        __syncthreads();
        for (int idx = threadIdx.x; idx < TD_ROWS * TD_PTS; idx += 32 * TD_WARPS) {
            const int pnt = idx / TD_ROWS, row = idx % TD_ROWS;            // consecutive threads -> consecutive n: coalesced
            if (pnt < mm && n0 + row < N) {
                double sum = 0.0;
#pragma unroll
                for (int w = 0; w < TD_WARPS; ++w) sum += part[w][row][pnt];
                T1[(int64_t)(m0 + pnt) * ldk + n0 + row] = sum;
            }
        }
        __syncthreads();
    }
}

// One lane's share of a length-N dot product with 8 independent accumulators: 16 loads in flight per lane instead of a
// dependent load -> fma chain (one L2 round trip per 32 elements: 34 us for the covariance pass at N = 4096).
__device__ __forceinline__ double warp_dot(const double* __restrict__ a, const double* __restrict__ b, int N, int lane) {
    double s[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    int n = lane;
    for (; n + 7 * 32 < N; n += 8 * 32) {
#pragma unroll
        for (int u = 0; u < 8; ++u) s[u] = fma(a[n + 32 * u], b[n + 32 * u], s[u]);
    }
#pragma unroll
    for (int u = 0; u < 7; ++u)
        if (n + 32 * u < N) s[u] = fma(a[n + 32 * u], b[n + 32 * u], s[u]);
    return ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}

// C0[a][b] for the M (M + 1) / 2 pairs b <= a, then mean0[m] = sum_n Kxt[m][n] alpha[n] for M more "pairs".  Each dot
// product of length N is split over the 4 warps of a quad (contiguous quarters, combined in a fixed order), two dot
// products per block: with one warp per dot product only ~30 SMs had work at M = 20.
__global__ void __launch_bounds__(256) rollout_cov_kernel(RbfDev k, int D, const double* __restrict__ pts,
                                                          const double* __restrict__ T1, const double* __restrict__ Kxt,
                                                          int64_t ldk, int N, int M, double* __restrict__ cov, int64_t ldc,
                                                          const double* __restrict__ alpha, double* __restrict__ mean0) {
    __shared__ double part[8];
    grid_dependency_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = blockIdx.x * 2 + (warp >> 2), seg = warp & 3;
    const int pairs = M * (M + 1) / 2;
    const bool live = pair < pairs + M;
    int a = 0, b = 0;
    const double* u = nullptr;
    const double* v = nullptr;
    if (live) {
        if (pair >= pairs) {
            u = Kxt + (int64_t)(pair - pairs) * ldk;
            v = alpha;
        } else {
            a = (int)((sqrt(8.0 * pair + 1.0) - 1.0) * 0.5);       // pair = a (a + 1) / 2 + b, b <= a
            while (a * (a + 1) / 2 > pair) --a;
            while ((a + 1) * (a + 2) / 2 <= pair) ++a;
            b = pair - a * (a + 1) / 2;
            u = T1 + (int64_t)a * ldk;
            v = Kxt + (int64_t)b * ldk;
        }
        const int quarter = (((N + 3) / 4) + 31) / 32 * 32;        // multiple of 32: every warp keeps lane-strided loads
        const int lo = seg * quarter, hi = (lo + quarter) < N ? (lo + quarter) : N;
        double s = (lo < hi) ? warp_dot(u + lo, v + lo, hi - lo, lane) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) part[warp] = s;
    }
    __syncthreads();
    if (live && seg == 0 && lane == 0) {
        const double s = (part[warp] + part[warp + 1]) + (part[warp + 2] + part[warp + 3]);
        if (pair >= pairs) {
            mean0[pair - pairs] = s;
        } else {
            const double kqq = (D == 2) ? rbf_eval<2>(k, pts + 2 * a, pts + 2 * b) : rbf_eval<3>(k, pts + 3 * a, pts + 3 * b);
            const double val = kqq - s;
            cov[(int64_t)a * ldc + b] = val;
            cov[(int64_t)b * ldc + a] = val;
        }
    }
}

}  // namespace ipp

#include "tc05.cuh"
#include "rollout_fused.cuh"

namespace ipp {

// IPP_ROLLOUT_FUSED=0 / ipp_rollout_mode(0) select the four-kernel sequence (A/B measurements and cross-checks; it is
// always used for more than 32 points)
static int64_t rf_part_doubles() {      // partials (one per CTA or cluster) + the streaming form's group sums, even count
    return ((int64_t)sm_count() + RF_MAX_GROUPS) * (32 * 33 / 2 + 32 + 2) / 2 * 2;
}
static int g_rollout_fused = -1;
static bool rollout_use_fused() {
    if (g_rollout_fused < 0) { const char* e = getenv("IPP_ROLLOUT_FUSED"); g_rollout_fused = (e && e[0] == '0') ? 0 : 1; }
    return g_rollout_fused != 0;
}

constexpr int RF_STREAM_MIN_N = 2048;       // from here on the images come from a pre-pass and W^-1 arrives by bulk copies

// W^-1 ([N][ldw] doubles) as a tensor map for the streaming form: boxes of 16 columns x 64 rows (one panel of a staged
// tile), 128-byte swizzle, rows / columns beyond N read as zeros.  Encoding costs ~1 us on the host; the search asks for
// the same belief hundreds of times in a row, so the last one is kept.
static int rollout_w_map(CUtensorMap* out, const double* W, int64_t N, int64_t ldw) {
    static thread_local struct { const double* W; int64_t N, ldw; CUtensorMap map; } last = {nullptr, 0, 0, {}};
    if (last.W != W || last.N != N || last.ldw != ldw) {
        EncodeTiledFn fn = encode_tiled_fn();
        if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return IPP_E_CUDA; }
        const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)N};
        const cuuint64_t strides[1] = {(cuuint64_t)ldw * sizeof(double)};
        const cuuint32_t box[2] = {16u, (cuuint32_t)RF_B};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = fn(&last.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(W), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { last.W = nullptr; set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return IPP_E_CUDA; }
        last.W = W; last.N = N; last.ldw = ldw;
    }
    *out = last.map;
    return IPP_OK;
}

template <int D, int NT, bool STREAM>
static int rollout_fused_launch(const RolloutFusedArgs& fa_in, const RolloutInline* inl, const double* pts_dev, double* img,
                                int grid, cudaStream_t stream) {
    constexpr int bytes = rf_smem_doubles<NT, STREAM>() * (int)sizeof(double);
    IPP_OPT_IN_SMEM((rollout_fused_kernel<D, NT, STREAM>), bytes);
    RolloutFusedArgs fa = fa_in;
    RolloutInlineSmall small;
    if (inl) memcpy(small.v, inl->v, sizeof(double) * (size_t)fa.r.M * (D + 1));      // [M][D] points, then [M] observations
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = RF_CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(STREAM ? RF_THREADS + 32 : RF_THREADS);      // streaming form: + the warp that feeds the ring
    cfg.dynamicSmemBytes = bytes;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // clusters of RF_CLUSTER CTAs only while all of them are resident at once (fewer whole clusters than single CTAs fit:
    // each needs RF_CLUSTER free SMs in one GPC); otherwise every CTA is its own cluster
    static int max_clusters[64];                          // per device; 0 = not asked yet
    int dev = 0;
    IPP_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && max_clusters[dev] == 0) {
        cfg.gridDim = dim3((unsigned)(sm_count() / RF_CLUSTER * RF_CLUSTER));
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, rollout_fused_kernel<D, NT, STREAM>, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = -1; }
        max_clusters[dev] = n;
    }
    const int fit = (dev >= 0 && dev < 64) ? max_clusters[dev] : -1;
    int g = grid;
    if (!STREAM && fit > 0 && (grid + RF_CLUSTER - 1) / RF_CLUSTER <= fit) {
        g = (grid + RF_CLUSTER - 1) / RF_CLUSTER * RF_CLUSTER;         // whole clusters; surplus CTAs get no tiles
    } else {
        attr[0].val.clusterDim.x = 1;
    }
    cfg.gridDim = dim3((unsigned)g);
    alignas(64) CUtensorMap wmap;
    memset(&wmap, 0, sizeof(wmap));
    if (STREAM) {
        const int rc = rollout_w_map(&wmap, fa.W, fa.N, fa.ldw);
        if (rc != IPP_OK) return rc;
        fa.img = img;
        rollout_images_kernel<D, NT><<<(unsigned)fa.T, RF_THREADS, 0, stream>>>(fa.kern, fa.X, fa.N, fa.alpha, pts_dev, inl ? 1 : 0, fa.r.M, small, img);
        IPP_LAUNCH_CHECK();
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;       // scheduled while the pre-pass runs; waits for
        attr[1].val.programmaticStreamSerializationAllowed = 1;                // it in griddepcontrol.wait
        cfg.numAttrs = 2;
    }
    IPP_CUDA(cudaLaunchKernelEx(&cfg, rollout_fused_kernel<D, NT, STREAM>, fa, small, wmap));
    return IPP_OK;
}

template <int D, int NT>
static int rollout_fused_dispatch(const RolloutFusedArgs& fa, const RolloutInline* inl, const double* pts_dev, double* img,
                                  int grid, cudaStream_t stream) {
    if (fa.N >= RF_STREAM_MIN_N) return rollout_fused_launch<D, NT, true>(fa, inl, pts_dev, img, grid, stream);
    return rollout_fused_launch<D, NT, false>(fa, inl, pts_dev, img, grid, stream);
}

// IPP_ROLLOUT_TW=simt selects the SIMT pass (A/B measurements); the tensor pass is the default
static bool tw_use_dmma() {
    static const int use = [] { const char* e = getenv("IPP_ROLLOUT_TW"); return (e && e[0] == 's') ? 0 : 1; }();
    return use != 0;
}

static cudaLaunchConfig_t dependent_launch(unsigned grid, unsigned block, size_t smem, cudaStream_t stream,
                                           cudaLaunchAttribute* attr) {
    attr->id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr->val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cfg;
}

static size_t rollout_smem(int M) {
    return sizeof(double) * ((size_t)M * (M + 1) + M + (size_t)(M + 1) * (RO_MAX_BLOCK + 1) + 2 * (RO_THREADS / 32));
}

}  // namespace ipp

using namespace ipp;

static int rf_ntiles(int M) { return M <= 16 ? 2 : (M <= 24 ? 3 : 4); }     // P = 8 NT padded points of the fused kernel

extern "C" int64_t ipp_rollout_workspace(int64_t N, int64_t M) {
    if (M <= 0) return 16;
    const int64_t ldc = round_up(M, 2);
    const int64_t four_kernels = 2 * M * round_up(N, 16) + M * ldc + round_up(M, 2) + 16;
    // one-launch form: one partial per cluster, and for N >= RF_STREAM_MIN_N four P x 64 images per 64-block of observations
    const int64_t T = (N + 63) / 64, P = 8 * rf_ntiles((int)(M < 32 ? M : 32));
    const int64_t fused = ipp::rf_part_doubles() + (N >= ipp::RF_STREAM_MIN_N ? T * 4 * P * 64 : 0) + 16;
    return four_kernels > fused ? four_kernels : fused;
}

namespace ipp {

// One rollout's launch sequence.  inline_pts != NULL: the [M][D] points followed by the [M] observations are passed by
// value to the first kernel, which also writes them to pts / zobs (device scratch); else pts / zobs already hold them.
// done != NULL: a host-mapped word the last kernel sets to done_seq once reward[] and info[] (then usually host-mapped
// too) are visible to the host.
int rollout_launch(const ipp_rbf* kern, const double* X, int64_t N, const double* alpha,
                   const double* winv, int64_t ldw, double noise_pred, double diag_add,
                   int kind, const double* params_host, int n_params, const double* maxes_dev, int64_t S,
                   double* pts, double* zobs, const RolloutInline* inline_pts, const int32_t* offsets_host,
                   const int32_t* reps_host, int L, double* reward, int32_t* info,
                   unsigned int* done, unsigned int done_seq, double* work, cudaStream_t stream,
                   double* tagged_out, unsigned long long tag, bool* used_tagged) {
    IPP_CHECK_ARG(kern && pts && zobs && offsets_host && reps_host && reward && info && work, "null pointer");
    if (used_tagged) *used_tagged = false;
    IPP_CHECK_ARG(L >= 1 && L <= RO_MAX_LEVELS, "1 <= levels <= IPP_ROLLOUT_MAX_LEVELS");
    IPP_CHECK_ARG(kind == IPP_REWARD_UCB || kind == IPP_REWARD_EI || kind == IPP_REWARD_MES || kind == IPP_REWARD_INFO,
                  "unknown reward kind");
    IPP_CHECK_ARG(N >= 1 && X && alpha && winv, "needs a belief with data (the reference's no-data early-outs stay with the caller)");
    const int M = offsets_host[L];
    IPP_CHECK_ARG(offsets_host[0] == 0 && M >= 1 && M <= RO_MAX_POINTS, "1 <= points <= IPP_ROLLOUT_MAX_POINTS");
    IPP_CHECK_ARG(!inline_pts || M <= RO_INLINE_POINTS, "too many points for the by-value form");
    RolloutArgs p;
    for (int l = 0; l < L; ++l) {
        const int kb = offsets_host[l + 1] - offsets_host[l];
        IPP_CHECK_ARG(kb >= 1 && kb <= RO_MAX_BLOCK, "each level needs 1..IPP_APPEND_MAX_K points");
        IPP_CHECK_ARG(reps_host[l] >= 0 && reps_host[l] <= 4, "reps must be 0..4");
        p.offsets[l] = offsets_host[l];
        p.reps[l] = reps_host[l];
    }
    p.offsets[L] = M;
    p.p0 = 0.0;
    if (kind == IPP_REWARD_MES) IPP_CHECK_ARG(maxes_dev && S >= 1, "MES needs the sampled maxima");
    else { IPP_CHECK_ARG(params_host && n_params >= 1, "need params[0]"); p.p0 = params_host[0]; }

    p.M = M; p.L = L; p.kind = kind; p.maxes = maxes_dev; p.S = (int)S;
    p.noise_pred = noise_pred; p.diag_add = diag_add; p.reward = reward; p.info = info;
    p.done = done; p.done_seq = done_seq;
    p.mean0 = nullptr; p.cov0 = nullptr; p.ldc = 0; p.zobs = zobs;
    if (rollout_use_fused() && M <= 32 && (kern->dim == 2 || kern->dim == 3)) {
        static unsigned int seq = 0;                      // a fresh ticket slot per launch (a slot is zero again when its launch ends)
        RolloutFusedArgs fa;
        fa.kern = to_dev(*kern);
        fa.X = X; fa.alpha = alpha; fa.W = winv; fa.ldw = ldw; fa.N = (int)N;
        fa.pts = pts; fa.zobs = zobs; fa.use_inline = inline_pts ? 1 : 0;
        fa.part = work;
        fa.part2 = work + (int64_t)sm_count() * (32 * 33 / 2 + 32 + 2);
        fa.slot = (int)(__atomic_fetch_add(&seq, 1u, __ATOMIC_RELAXED) % RF_SLOTS);
        fa.tag = 0;
        if (tagged_out && tag) {                          // results as self-validating records, no completion word
            fa.tag = tag;
            p.reward = tagged_out;
        }
        fa.T = (int)((N + RF_B - 1) / RF_B);
        fa.tiles = fa.T * (fa.T + 1) / 2;
        fa.r = p;
        if (used_tagged) *used_tagged = fa.tag != 0;
        // small beliefs: a CTA per tpc tiles (IPP_ROLLOUT_TPC, default 1: shortest tile phase); large ones: one CTA per SM
        static const int tpc = [] { const char* e = getenv("IPP_ROLLOUT_TPC"); const int v = e ? atoi(e) : 1; return v >= 1 ? v : 1; }();
        int grid = (fa.tiles + tpc - 1) / tpc;
        if (grid > sm_count()) grid = sm_count();
        const int nt = rf_ntiles(M);
        const bool d2 = kern->dim == 2;
        fa.img = nullptr;
        double* img = work + rf_part_doubles();            // the streaming form's images, behind the partials
        if (nt == 2) return d2 ? rollout_fused_dispatch<2, 2>(fa, inline_pts, pts, img, grid, stream) : rollout_fused_dispatch<3, 2>(fa, inline_pts, pts, img, grid, stream);
        if (nt == 3) return d2 ? rollout_fused_dispatch<2, 3>(fa, inline_pts, pts, img, grid, stream) : rollout_fused_dispatch<3, 3>(fa, inline_pts, pts, img, grid, stream);
        return d2 ? rollout_fused_dispatch<2, 4>(fa, inline_pts, pts, img, grid, stream) : rollout_fused_dispatch<3, 4>(fa, inline_pts, pts, img, grid, stream);
    }
    IPP_CHECK_ARG(kind != IPP_REWARD_INFO, "information-gain rollouts take at most 32 points (one-launch form only)");
    const int64_t ldc = round_up(M, 2);
    double* cov0 = work;
    double* mean0 = cov0 + (int64_t)M * ldc;
    double* sub = mean0 + round_up(M, 2);
    cudaLaunchAttribute pdl;
    {
        const int64_t ldk = round_up(N, 16);
        double* Kxt = sub;
        double* T1 = sub + (int64_t)M * ldk;
        const RbfDev kd = to_dev(*kern);
        const unsigned nb = (unsigned)((N + 255) / 256);
        if (inline_pts) {
            if (kern->dim == 2) rollout_rbf_inline_kernel<2><<<nb, 256, 0, stream>>>(kd, X, (int)N, *inline_pts, M, pts, zobs, Kxt, ldk);
            else rollout_rbf_inline_kernel<3><<<nb, 256, 0, stream>>>(kd, X, (int)N, *inline_pts, M, pts, zobs, Kxt, ldk);
        } else {
            if (kern->dim == 2) rollout_rbf_kernel<2><<<nb, 256, 0, stream>>>(kd, X, (int)N, pts, M, Kxt, ldk);
            else rollout_rbf_kernel<3><<<nb, 256, 0, stream>>>(kd, X, (int)N, pts, M, Kxt, ldk);
        }
        IPP_LAUNCH_CHECK();
        if (tw_use_dmma()) {
            cudaLaunchConfig_t cfg = dependent_launch((unsigned)((N + TD_ROWS - 1) / TD_ROWS), 32 * TD_WARPS, 0, stream, &pdl);
            IPP_CUDA(cudaLaunchKernelEx(&cfg, rollout_tw_dmma_kernel, winv, ldw, (int)N, (const double*)Kxt, ldk, M, T1));
        } else {
            IPP_OPT_IN_SMEM(rollout_tw_kernel, TW_SMEM);
            constexpr int rows_per_block = TW_ROWS_PER_BLOCK;
            rollout_tw_kernel<<<(unsigned)((N + rows_per_block - 1) / rows_per_block), 32 * TW_WARPS, TW_SMEM, stream>>>(
                winv, ldw, (int)N, Kxt, ldk, M, T1);
        }
        IPP_LAUNCH_CHECK();
        const int pairs = M * (M + 1) / 2 + M;                  // covariance pairs + the M mean dot products
        {
            cudaLaunchConfig_t cfg = dependent_launch((unsigned)((pairs + 1) / 2), 256, 0, stream, &pdl);
            IPP_CUDA(cudaLaunchKernelEx(&cfg, rollout_cov_kernel, to_dev(*kern), (int)kern->dim, (const double*)pts,
                                        (const double*)T1, (const double*)Kxt, ldk, (int)N, M, cov0, ldc, alpha, mean0));
        }
        IPP_LAUNCH_CHECK();
    }
    p.mean0 = mean0; p.cov0 = cov0; p.ldc = ldc;
    IPP_OPT_IN_SMEM(rollout_condition_kernel, (int)rollout_smem(RO_MAX_POINTS));
    {
        cudaLaunchConfig_t cfg = dependent_launch(1, RO_THREADS, rollout_smem(M), stream, &pdl);
        IPP_CUDA(cudaLaunchKernelEx(&cfg, rollout_condition_kernel, p));
    }
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

}  // namespace ipp

extern "C" int ipp_rollout_debug_clocks(int64_t out[32]) {
    long long v[32];
    IPP_CUDA(cudaDeviceSynchronize());
    IPP_CUDA(cudaMemcpyFromSymbol(v, ipp::g_rf_clock, sizeof(v)));
    for (int i = 0; i < 32; ++i) out[i] = v[i];
    return IPP_OK;
}

extern "C" int ipp_rollout_mode(int fused) {
    const int before = ipp::rollout_use_fused() ? 1 : 0;
    if (fused == 0 || fused == 1) ipp::g_rollout_fused = fused;
    return before;
}

extern "C" int ipp_rollout_rewards(const ipp_rbf* kern, const double* X, int64_t N, const double* alpha,
                                   const double* winv, int64_t ldw, double noise_pred, double diag_add,
                                   int kind, const double* params_host, int n_params, const double* maxes_dev, int64_t S,
                                   const double* pts, const double* zobs, const int32_t* offsets_host,
                                   const int32_t* reps_host, int L, double* reward, int32_t* info,
                                   double* work, void* stream_) {
    return rollout_launch(kern, X, N, alpha, winv, ldw, noise_pred, diag_add, kind, params_host, n_params, maxes_dev, S,
                          const_cast<double*>(pts), const_cast<double*>(zobs), nullptr, offsets_host, reps_host, L,
                          reward, info, nullptr, 0u, work, static_cast<cudaStream_t>(stream_), nullptr, 0ull, nullptr);
}
