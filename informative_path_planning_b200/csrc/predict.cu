// a3 / a5: GP posterior mean and variance for a batch of query points.
//
//   mean[m] = sum_n k(q_m, x_n) alpha[n]                      (rbf_queries_kernel, fused with generation)
//   var[m]  = variance - quad(m) + noise_add
//     IPP_VAR_WINV_FULL  quad = sum_n (sum_j W[n][j] k_j) k_n            every 128x128 tile   (2 N^2 flop / query)
//     IPP_VAR_WINV_SYM   same products, W symmetric: strictly-lower tiles counted twice,
//                        diagonal tiles once                                                  (  N^2 flop / query)
//     IPP_VAR_CHOL       quad = sum_n (sum_{j<=n} Linv[n][j] k_j)^2      lower tiles only     (  N^2 flop / query)
//
// The cross-covariance of one chunk of queries is materialised query-major (Kxt[m][n]) once, so the
// exp is evaluated once per (m, n); the quadratic form is a DMMA GEMM  C[m][n] = sum_j Kxt[m][j] mat[n][j]
// whose 128x128 output tile is never stored: the epilogue multiplies it with Kxt[m][n] (or squares it),
// reduces over n with quad shuffles + a 4-warp shared-memory tree, and writes one partial per
// (n-tile, m).  A persistent grid (one CTA per SM) pulls (n-tile, m-tile) items from an atomic
// counter, heaviest n-tiles first.  A last pass adds the partials in a fixed order (deterministic).
#include "internal.cuh"
#include "dgemm_dmma.cuh"

namespace ipp {

struct PredictArgs {
    const double* Kxt; int64_t ldk; int Mc;
    const double* mat; int64_t ldm; int N;
    int mode;
    double* partial; int64_t ldp;
    int* counter;
    int T, Mtiles;
    int GM;            // m-tiles per L2 group: a group's Kxt panel (GM * 128 rows) stays L2-resident during its n-tile sweep
};

__global__ void __launch_bounds__(GEMM_THREADS, 1) predict_var_kernel(PredictArgs p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double s_red[Cfg::WARPS_N][BM];
    __shared__ int s_item;
    const uint32_t smem_base = smem_u32(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp / Cfg::WARPS_N) * WM, wn0 = (warp % Cfg::WARPS_N) * WN;
    const int total = p.T * p.Mtiles;

    for (;;) {
        if (tid == 0) s_item = atomicAdd(p.counter, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= total) break;
        // item -> (m-group, n-tile I heaviest first, m-tile within the group)
        const int grp = item / (p.T * p.GM);
        const int r = item - grp * p.T * p.GM;
        const int gm = (p.Mtiles - grp * p.GM) < p.GM ? (p.Mtiles - grp * p.GM) : p.GM;
        const int I = p.T - 1 - r / gm;
        const int m0 = (grp * p.GM + r % gm) * BM, n0 = I * BN;
        const int ndiag_end = (n0 + BN < p.N) ? (n0 + BN) : p.N;

        double acc[MI][NI][2];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        if (p.mode == IPP_VAR_WINV_SYM) {
            if (n0 > 0) {
                gemm_mainloop(smem_base, p.Kxt, p.ldk, m0, p.Mc, p.mat, p.ldm, n0, p.N, 0, n0, acc, tid, wm0, wn0, g, t);
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NI; ++j) { acc[i][j][0] *= 2.0; acc[i][j][1] *= 2.0; }
            }
            gemm_mainloop(smem_base, p.Kxt, p.ldk, m0, p.Mc, p.mat, p.ldm, n0, p.N, n0, ndiag_end, acc, tid, wm0, wn0, g, t);
        } else if (p.mode == IPP_VAR_CHOL) {
            gemm_mainloop(smem_base, p.Kxt, p.ldk, m0, p.Mc, p.mat, p.ldm, n0, p.N, 0, ndiag_end, acc, tid, wm0, wn0, g, t);
        } else {
            gemm_mainloop(smem_base, p.Kxt, p.ldk, m0, p.Mc, p.mat, p.ldm, n0, p.N, 0, p.N, acc, tid, wm0, wn0, g, t);
        }

        // fused epilogue: reduce the tile against Kxt[m][n0 .. n0+127] (or its own square) over n
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const int row = m0 + wm0 + 8 * i + g;
            double s = 0.0;
            if (row < p.Mc) {
                const double* kr = p.Kxt + (int64_t)row * p.ldk;
#pragma unroll
                for (int j = 0; j < NI; ++j) {
                    const int col = n0 + wn0 + 8 * j + 2 * t;
                    if (p.mode == IPP_VAR_CHOL) {
                        s = fma(acc[i][j][0], acc[i][j][0], s);
                        s = fma(acc[i][j][1], acc[i][j][1], s);
                    } else if (col + 1 < p.N) {
                        const double2 kv = *reinterpret_cast<const double2*>(kr + col);
                        s = fma(acc[i][j][0], kv.x, s);
                        s = fma(acc[i][j][1], kv.y, s);
                    } else if (col < p.N) {
                        s = fma(acc[i][j][0], kr[col], s);
                    }
                }
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (t == 0) s_red[warp % Cfg::WARPS_N][wm0 + 8 * i + g] = s;
        }
        __syncthreads();
        if (tid < BM && m0 + tid < p.Mc) {
            double q = 0.0;
#pragma unroll
            for (int w = 0; w < Cfg::WARPS_N; ++w) q += s_red[w][tid];
            p.partial[(int64_t)I * p.ldp + m0 + tid] = q;
        }
        __syncthreads();
    }
}

__global__ void predict_finalize_kernel(const double* __restrict__ partial, int64_t ldp, int T, int Mc,
                                        double variance, double noise_add, double* __restrict__ var) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= Mc) return;
    double q = 0.0;
    for (int I = 0; I < T; ++I) q += partial[(int64_t)I * ldp + m];
    var[m] = (variance - q) + noise_add;
}

__global__ void prior_fill_kernel(double* mean, double* var, int64_t M, double v) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m < M) { mean[m] = 0.0; var[m] = v; }
}

__global__ void add_scalar_kernel(double* A, int64_t rows, int64_t cols, int64_t ld, double v) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < rows * cols) A[(idx / cols) * ld + idx % cols] += v;
}

// ---------------------------------------------------------------------------------------------
// Small-batch path (M <= SMALL_M queries per pass): the planner's real call shape is k ~ 4 sample
// points per path (paths_library.py:152-171), for which a 128-row DMMA tile is 97 % padding.  Here the
// N x N operand is streamed ONCE across all SMs (HBM/L2-bound, N^2 * 8 B, half of it in the
// symmetric / triangular forms): one warp per matrix row n, the <= 32 queries' cross-covariances are
// accumulated in registers, t[n][m] = sum_j mat[n][j] Kxt[m][j], then reduced against Kxt[m][n] into
// one partial per (CTA, m); a fixed-order finalize adds the partials.
constexpr int SMALL_M = 32;
constexpr int SMALL_ROWS_PER_CTA = 32;     // 8 warps x 4 rows

template <int MT>      // MT = register tier (number of query accumulators per lane), M <= MT <= SMALL_M
__global__ void __launch_bounds__(256) predict_var_small_kernel(const double* __restrict__ Kxt, int64_t ldk, int M,
                                                                const double* __restrict__ mat, int64_t ldm, int N,
                                                                int mode, double* __restrict__ partial) {
    __shared__ double s_part[8][SMALL_M];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double cta_acc[MT];                 // lane 0 of each warp accumulates its rows' contributions
#pragma unroll
    for (int m = 0; m < MT; ++m) cta_acc[m] = 0.0;
    // rows are dealt round-robin to CTAs so that the triangular forms (row n has n+1 terms) balance
    for (int rr = warp; rr < SMALL_ROWS_PER_CTA; rr += 8) {
        const int n = (int)blockIdx.x + rr * (int)gridDim.x;
        if (n >= N) break;
        const double* row = mat + (int64_t)n * ldm;
        const int jend = (mode == IPP_VAR_WINV_FULL) ? N : (mode == IPP_VAR_CHOL ? n + 1 : n);
        double acc[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) acc[m] = 0.0;
        for (int j = lane; j < jend; j += 32) {
            const double w = row[j];
#pragma unroll
            for (int m = 0; m < MT; ++m)
                if (m < M) acc[m] = fma(w, Kxt[(int64_t)m * ldk + j], acc[m]);
        }
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            if (m < M) {
                double t = acc[m];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                if (lane == 0) {
                    const double kn = Kxt[(int64_t)m * ldk + n];
                    double c;
                    if (mode == IPP_VAR_CHOL) c = t * t;
                    else if (mode == IPP_VAR_WINV_FULL) c = t * kn;
                    else c = kn * fma(2.0, t, row[n] * kn);          // symmetric: 2 * strictly-lower + diagonal
                    cta_acc[m] += c;
                }
            }
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int m = 0; m < MT; ++m) s_part[warp][m] = cta_acc[m];
    }
    __syncthreads();
    if (threadIdx.x < M) {
        double q = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) q += s_part[w][threadIdx.x];
        partial[(int64_t)blockIdx.x * SMALL_M + threadIdx.x] = q;
    }
}

__global__ void predict_finalize_small_kernel(const double* __restrict__ partial, int nblocks, int M,
                                              double variance, double noise_add, double* __restrict__ var) {
    const int m = threadIdx.x;
    if (m >= M) return;
    double q = 0.0;
    for (int b = 0; b < nblocks; ++b) q += partial[(int64_t)b * SMALL_M + m];
    var[m] = (variance - q) + noise_add;
}

constexpr int64_t PREDICT_CHUNK = 32768;

static int64_t chunk_rows(int64_t M) { return M < PREDICT_CHUNK ? round_up(M, BM) : PREDICT_CHUNK; }

}  // namespace ipp

using namespace ipp;

extern "C" int64_t ipp_gp_predict_workspace(int64_t N, int64_t M) {
    if (N <= 0 || M <= 0) return 16;
    const int64_t ldk = round_up(N, 16), mc = chunk_rows(M), T = (N + BN - 1) / BN;
    return mc * ldk + T * mc + N + 64;     // (+N: partials of the small-batch path)
}

extern "C" int ipp_gp_predict(const ipp_rbf* kern, const double* X, int64_t N, const double* alpha,
                              const double* mat, int64_t ldm, int var_mode, double noise_add,
                              const double* Xq, int64_t M, double* mean, double* var,
                              double* work, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    IPP_CHECK_ARG(kern && Xq && mean && var, "null pointer");
    IPP_CHECK_ARG(M >= 1, "need at least one query point");
    IPP_CHECK_ARG(kern->dim == 2 || kern->dim == 3, "dimension must be 2 or 3");
    if (N == 0) {   // no observations: prior (gpmodel_library.py:310-311; noise is NOT added there)
        prior_fill_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(mean, var, M, kern->variance);
        IPP_LAUNCH_CHECK();
        return IPP_OK;
    }
    IPP_CHECK_ARG(X && alpha && mat && work, "null pointer");
    IPP_CHECK_ARG(var_mode >= 0 && var_mode <= 2, "bad var_mode");
    IPP_CHECK_ARG(ldm >= N && ldm % 2 == 0 && reinterpret_cast<uintptr_t>(mat) % 16 == 0, "mat must be 16-byte aligned, ldm even");
    IPP_CHECK_ARG(N < (1LL << 30), "N too large");
    IPP_OPT_IN_SMEM(predict_var_kernel, GEMM_SMEM);
    const int64_t ldk = round_up(N, 16), mc_max = chunk_rows(M);
    const int T = (int)((N + BN - 1) / BN);
    if (M <= 2 * SMALL_M) {                       // planner-sized batches: stream the operand once
        double* Kxt_s = work;
        double* part_s = work + (int64_t)round_up(M, BM) * ldk;
        const int nblocks = (int)((N + SMALL_ROWS_PER_CTA - 1) / SMALL_ROWS_PER_CTA);
        IPP_TRY(launch_rbf_queries(*kern, X, N, alpha, Xq, M, Kxt_s, ldk, mean, stream));
        for (int64_t s0 = 0; s0 < M; s0 += SMALL_M) {
            const int mm = (int)((M - s0) < SMALL_M ? (M - s0) : SMALL_M);
            if (mm <= 8)
                predict_var_small_kernel<8><<<nblocks, 256, 0, stream>>>(Kxt_s + s0 * ldk, ldk, mm, mat, ldm, (int)N, var_mode, part_s);
            else
                predict_var_small_kernel<SMALL_M><<<nblocks, 256, 0, stream>>>(Kxt_s + s0 * ldk, ldk, mm, mat, ldm, (int)N, var_mode, part_s);
            IPP_LAUNCH_CHECK();
            predict_finalize_small_kernel<<<1, SMALL_M, 0, stream>>>(part_s, nblocks, mm, kern->variance, noise_add, var + s0);
            IPP_LAUNCH_CHECK();
        }
        return IPP_OK;
    }
    double* Kxt = work;
    double* partial = work + mc_max * ldk;
    int* counter = reinterpret_cast<int*>(partial + (int64_t)T * mc_max);
    const int grid = sm_count();
    for (int64_t s = 0; s < M; s += mc_max) {
        const int64_t mc = (M - s) < mc_max ? (M - s) : mc_max;
        IPP_TRY(launch_rbf_queries(*kern, X, N, alpha, Xq + s * kern->dim, mc, Kxt, ldk, mean + s, stream));
        IPP_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), stream));
        PredictArgs p;
        p.Kxt = Kxt; p.ldk = ldk; p.Mc = (int)mc;
        p.mat = mat; p.ldm = ldm; p.N = (int)N;
        p.mode = var_mode;
        p.partial = partial; p.ldp = mc_max;
        p.counter = counter;
        p.T = T; p.Mtiles = (int)((mc + BM - 1) / BM);
        {
            const int64_t panel = (int64_t)BM * ldk * 8;                 // bytes of one m-tile's Kxt panel
            int64_t gm = (48LL << 20) / panel;                           // ~48 MB of the 126 MB L2 per group
            if (gm < 1) gm = 1;
            if (gm > p.Mtiles) gm = p.Mtiles;
            p.GM = (int)gm;
        }
        const int items = p.T * p.Mtiles;
        profile_mark(0, stream);
        predict_var_kernel<<<items < grid ? items : grid, GEMM_THREADS, GEMM_SMEM, stream>>>(p);
        profile_mark(1, stream);
        IPP_LAUNCH_CHECK();
        predict_finalize_kernel<<<(unsigned)((mc + 255) / 256), 256, 0, stream>>>(partial, mc_max, T, (int)mc,
                                                                                 kern->variance, noise_add, var + s);
        IPP_LAUNCH_CHECK();
    }
    return IPP_OK;
}

extern "C" int64_t ipp_gp_predict_cov_workspace(int64_t N, int64_t M) {
    if (N <= 0 || M <= 0) return 16;
    return 2 * round_up(M, 2) * round_up(N, 16) + 16;
}

extern "C" int ipp_gp_predict_cov(const ipp_rbf* kern, const double* X, int64_t N, const double* alpha,
                                  const double* winv, int64_t ldm, double noise_add,
                                  const double* Xq, int64_t M, double* mean, double* cov, int64_t ldc,
                                  double* work, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    IPP_CHECK_ARG(kern && X && alpha && winv && Xq && mean && cov && work, "null pointer");
    IPP_CHECK_ARG(N >= 1 && M >= 1 && M < (1 << 20), "bad shape");
    IPP_CHECK_ARG(ldc >= M && ldc % 2 == 0, "ldc must be even and >= M");
    const int64_t ldk = round_up(N, 16);
    double* Kxt = work;
    double* T1 = work + round_up(M, 2) * ldk;
    IPP_TRY(launch_rbf_queries(*kern, X, N, alpha, Xq, M, Kxt, ldk, mean, stream));
    GemmArgs g{};
    g.A = Kxt; g.lda = ldk; g.B = winv; g.ldb = ldm; g.C = T1; g.ldc = ldk;
    g.M = (int)M; g.N = (int)N; g.K = (int)N; g.alpha = 1.0; g.beta = 0.0;
    IPP_TRY(launch_gemm(g, stream));                                   // T1 = Kx^T W^-1
    IPP_TRY(launch_rbf_cross(*kern, Xq, M, Xq, M, cov, ldc, stream));  // K(q, q)
    GemmArgs h{};
    h.A = T1; h.lda = ldk; h.B = Kxt; h.ldb = ldk; h.C = cov; h.ldc = ldc;
    h.M = (int)M; h.N = (int)M; h.K = (int)N; h.alpha = -1.0; h.beta = 1.0;
    IPP_TRY(launch_gemm(h, stream));                                   // cov -= T1 Kx
    if (noise_add != 0.0) {
        add_scalar_kernel<<<(unsigned)((M * M + 255) / 256), 256, 0, stream>>>(cov, M, M, ldc, noise_add);
        IPP_LAUNCH_CHECK();
    }
    return IPP_OK;
}
