// a11: random-Fourier-feature posterior function samples evaluated on a grid, with a per-sample
// arg-max (general_target aq_library.py:138-139; grid stage of global_maximization :475-479), and
// the feature matrix Z = scale * cos(W X^T + b) (:171).
//   f_s(x_g) = sum_f theta[s][f] cos(W_s[f] . x_g + b_s[f])
// shared_w (S >= RFF_GEMM_MIN_S): values = Phi(grid) Theta^T is a dense FP64 contraction.  Phi[g][f] =
//   cos(W_f . x_g + b_f) is materialised once per chunk of grid points (one cos per (g, f)) and the
//   contraction runs on the DMMA tensor path (dgemm_dmma.cuh) with a fused epilogue: the 128 x 128
//   tile of function values is never stored (unless `values` is asked for); each sample column is
//   reduced to (max, first arg-max) with quad/warp shuffles + a 2-warp shared-memory step.
// shared_w with few samples: one cos per (g, f) reused by up to 32 samples held in registers (SIMT).
// per-sample W (the reference's sampling scheme) has no reuse and is cos-bound.
#include "internal.cuh"
#include "dgemm_dmma.cuh"

namespace ipp {

template <int D, int SC>
__global__ void __launch_bounds__(256) rff_eval_kernel(const double* __restrict__ W, const double* __restrict__ b,
                                                       const double* __restrict__ theta, int F, int S, int shared_w,
                                                       const double* __restrict__ grid, int64_t G,
                                                       double* __restrict__ values, double* __restrict__ part_val,
                                                       int64_t* __restrict__ part_idx) {
    __shared__ double sv[8];
    __shared__ int64_t si[8];
    const int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = gidx < G;
    double x[D];
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] = live ? grid[gidx * D + d] : 0.0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int s0 = 0; s0 < S; s0 += SC) {
        double acc[SC];
#pragma unroll
        for (int j = 0; j < SC; ++j) acc[j] = 0.0;
        if (shared_w) {
            for (int f = 0; f < F; ++f) {
                double ph = b[f];
#pragma unroll
                for (int d = 0; d < D; ++d) ph = fma(W[f * D + d], x[d], ph);
                const double c = cos(ph);
#pragma unroll
                for (int j = 0; j < SC; ++j)
                    if (s0 + j < S) acc[j] = fma(theta[(int64_t)(s0 + j) * F + f], c, acc[j]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < SC; ++j) {
                if (s0 + j < S) {
                    const double* Ws = W + (int64_t)(s0 + j) * F * D;
                    const double* bs = b + (int64_t)(s0 + j) * F;
                    const double* ts = theta + (int64_t)(s0 + j) * F;
                    double a = 0.0;
                    for (int f = 0; f < F; ++f) {
                        double ph = bs[f];
#pragma unroll
                        for (int d = 0; d < D; ++d) ph = fma(Ws[f * D + d], x[d], ph);
                        a = fma(ts[f], cos(ph), a);
                    }
                    acc[j] = a;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < SC; ++j) {
            if (s0 + j >= S) break;
            if (values && live) values[(int64_t)(s0 + j) * G + gidx] = acc[j];
            double v = live ? acc[j] : -INFINITY;
            int64_t ix = live ? gidx : INT64_MAX;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, v, o);
                const int64_t oi = __shfl_xor_sync(0xffffffffu, ix, o);
                if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
            }
            if (lane == 0) { sv[warp] = v; si[warp] = ix; }
            __syncthreads();
            if (threadIdx.x == 0) {
                double bv = sv[0];
                int64_t bi = si[0];
                for (int w = 1; w < 8; ++w)
                    if (sv[w] > bv || (sv[w] == bv && si[w] < bi)) { bv = sv[w]; bi = si[w]; }
                part_val[(int64_t)(s0 + j) * gridDim.x + blockIdx.x] = bv;
                part_idx[(int64_t)(s0 + j) * gridDim.x + blockIdx.x] = bi;
            }
            __syncthreads();
        }
    }
}

// one block per sample: first index of the maximum, like np.argmax
__global__ void __launch_bounds__(256) rff_argmax_finalize_kernel(const double* __restrict__ part_val,
                                                                  const int64_t* __restrict__ part_idx, int nblocks,
                                                                  double* __restrict__ max_val,
                                                                  int64_t* __restrict__ arg_max) {
    __shared__ double sv[256];
    __shared__ int64_t si[256];
    const int s = blockIdx.x;
    double v = -INFINITY;
    int64_t ix = INT64_MAX;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) {
        const double ov = part_val[(int64_t)s * nblocks + i];
        const int64_t oi = part_idx[(int64_t)s * nblocks + i];
        if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
    }
    sv[threadIdx.x] = v;
    si[threadIdx.x] = ix;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double ov = sv[threadIdx.x + o];
            const int64_t oi = si[threadIdx.x + o];
            if (ov > sv[threadIdx.x] || (ov == sv[threadIdx.x] && oi < si[threadIdx.x])) {
                sv[threadIdx.x] = ov;
                si[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { max_val[s] = sv[0]; arg_max[s] = si[0]; }
}

template <int D>
__global__ void rff_features_kernel(const double* __restrict__ W, const double* __restrict__ b, int64_t F,
                                    const double* __restrict__ X, int64_t N, double scale, double* __restrict__ Z,
                                    int64_t ldz) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t f = blockIdx.y;
    if (n >= N) return;
    double ph = b[f];
#pragma unroll
    for (int d = 0; d < D; ++d) ph = fma(W[f * D + d], X[n * D + d], ph);
    Z[f * ldz + n] = scale * cos(ph);
}

// ---- shared-feature GEMM path -------------------------------------------------------------------
constexpr int RFF_GEMM_MIN_S = 8;
constexpr int64_t RFF_CHUNK = 148 * 128 * 2;      // grid points per Phi chunk: two full waves of 128-row tiles

// Phi[g][f] = cos(W_f . x_g + b_f) for grid points [g0, g0 + Gc), row-major [Gc][ldf]; lanes run along f
template <int D>
__global__ void __launch_bounds__(256) rff_phi_kernel(const double* __restrict__ W, const double* __restrict__ b, int F,
                                                      const double* __restrict__ grid, int64_t g0, int Gc,
                                                      double* __restrict__ Phi, int64_t ldf) {
    constexpr int ROWS = 16;
    __shared__ double sx[ROWS][D];
    const int r0 = blockIdx.x * ROWS;
    if (threadIdx.x < ROWS * D) {
        const int r = threadIdx.x / D, d = threadIdx.x % D;
        sx[r][d] = (r0 + r < Gc) ? grid[(g0 + r0 + r) * D + d] : 0.0;
    }
    __syncthreads();
    for (int f = threadIdx.x; f < F; f += 256) {
        double w[D];
#pragma unroll
        for (int d = 0; d < D; ++d) w[d] = W[f * D + d];
        const double bf = b[f];
#pragma unroll 4
        for (int r = 0; r < ROWS; ++r) {
            if (r0 + r >= Gc) break;
            double ph = bf;
#pragma unroll
            for (int d = 0; d < D; ++d) ph = fma(w[d], sx[r][d], ph);
            Phi[(int64_t)(r0 + r) * ldf + f] = cos(ph);
        }
    }
}

__global__ void rff_pad_theta_kernel(const double* __restrict__ theta, int S, int F, double* __restrict__ out, int64_t ldf) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)S * F) out[(i / F) * ldf + i % F] = theta[i];
}

struct RffGemmArgs {
    const double* Phi; int64_t ldf; int Gc;      // A operand: [Gc][ldf]
    const double* Th; int S; int F;              // B operand: [S][ldf]
    int64_t g0, G;                               // global index of the chunk's first grid point, total grid size
    double* values;                              // optional [S][G]
    double* part_val; int64_t* part_idx;         // [S][nparts]
    int nparts, part0;                           // this chunk's m-tiles write columns part0 + blockIdx.x
};

__global__ void __launch_bounds__(GEMM_THREADS, 1) rff_gemm_argmax_kernel(RffGemmArgs p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double s_val[Cfg::WARPS_M][BN];
    __shared__ int s_row[Cfg::WARPS_M][BN];
    const uint32_t smem_base = smem_u32(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wmi = warp / Cfg::WARPS_N;
    const int wm0 = wmi * WM, wn0 = (warp % Cfg::WARPS_N) * WN;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    gemm_mainloop(smem_base, p.Phi, p.ldf, m0, p.Gc, p.Th, p.ldf, n0, p.S, 0, p.F, acc, tid, wm0, wn0, g, t);

    if (p.values) {
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const int row = m0 + wm0 + 8 * i + g;
            if (row >= p.Gc) continue;
#pragma unroll
            for (int j = 0; j < NI; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = n0 + wn0 + 8 * j + 2 * t + e;
                    if (col < p.S) p.values[(int64_t)col * p.G + p.g0 + row] = acc[i][j][e];
                }
        }
    }
    // per sample column: (max, first row attaining it) over the tile's 128 grid points
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double v = -INFINITY;
            int row = 0x7fffffff;
#pragma unroll
            for (int i = 0; i < MI; ++i) {          // rows ascend with i: strict > keeps the first maximum
                const int r = m0 + wm0 + 8 * i + g;
                if (r < p.Gc && acc[i][j][e] > v) { v = acc[i][j][e]; row = r; }
            }
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, v, o);
                const int orow = __shfl_xor_sync(0xffffffffu, row, o);
                if (ov > v || (ov == v && orow < row)) { v = ov; row = orow; }
            }
            if (g == 0) {
                s_val[wmi][wn0 + 8 * j + 2 * t + e] = v;
                s_row[wmi][wn0 + 8 * j + 2 * t + e] = row;
            }
        }
    __syncthreads();
    if (tid < BN && n0 + tid < p.S) {
        double v = s_val[0][tid];
        int row = s_row[0][tid];
#pragma unroll
        for (int w = 1; w < Cfg::WARPS_M; ++w)
            if (s_val[w][tid] > v || (s_val[w][tid] == v && s_row[w][tid] < row)) { v = s_val[w][tid]; row = s_row[w][tid]; }
        const int64_t slot = (int64_t)(n0 + tid) * p.nparts + p.part0 + blockIdx.x;
        p.part_val[slot] = v;
        p.part_idx[slot] = (row == 0x7fffffff) ? INT64_MAX : p.g0 + row;
    }
}

static int64_t rff_chunk_rows(int64_t G) { return G < RFF_CHUNK ? round_up(G, BM) : RFF_CHUNK; }
static bool rff_use_gemm(int64_t S, int shared_w) { return shared_w && S >= RFF_GEMM_MIN_S; }

}  // namespace ipp

using namespace ipp;

extern "C" int64_t ipp_rff_workspace(int64_t F, int64_t S, int64_t G, int shared_w) {
    if (F <= 0 || S <= 0 || G <= 0) return 16;
    if (!rff_use_gemm(S, shared_w)) return 2 * S * ((G + 255) / 256) + 16;
    const int64_t ldf = round_up(F, 16);
    return rff_chunk_rows(G) * ldf + S * ldf + 2 * S * ((G + BM - 1) / BM) + 16;
}

extern "C" int ipp_rff_eval_argmax(const double* W, const double* b, const double* theta, int64_t F, int64_t S, int D,
                                   int shared_w, const double* grid, int64_t G, double* values,
                                   double* max_val, int64_t* arg_max, double* work, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    IPP_CHECK_ARG(W && b && theta && grid && max_val && arg_max && work, "null pointer");
    IPP_CHECK_ARG(F >= 1 && S >= 1 && G >= 1 && F < (1 << 30) && S < (1 << 30), "bad shape");
    IPP_CHECK_ARG(D == 2 || D == 3, "dimension must be 2 or 3");
    if (rff_use_gemm(S, shared_w)) {
        IPP_OPT_IN_SMEM(rff_gemm_argmax_kernel, GEMM_SMEM);
        const int64_t ldf = round_up(F, 16), gc_max = rff_chunk_rows(G);
        const int nparts = (int)((G + BM - 1) / BM);
        double* Phi = work;
        double* Th = Phi + gc_max * ldf;
        double* part_val = Th + S * ldf;
        int64_t* part_idx = reinterpret_cast<int64_t*>(part_val + (int64_t)S * nparts);
        rff_pad_theta_kernel<<<(unsigned)((S * F + 255) / 256), 256, 0, s>>>(theta, (int)S, (int)F, Th, ldf);
        IPP_LAUNCH_CHECK();
        for (int64_t g0 = 0; g0 < G; g0 += gc_max) {
            const int gc = (int)((G - g0) < gc_max ? (G - g0) : gc_max);
            if (D == 2) rff_phi_kernel<2><<<(gc + 15) / 16, 256, 0, s>>>(W, b, (int)F, grid, g0, gc, Phi, ldf);
            else rff_phi_kernel<3><<<(gc + 15) / 16, 256, 0, s>>>(W, b, (int)F, grid, g0, gc, Phi, ldf);
            IPP_LAUNCH_CHECK();
            RffGemmArgs p;
            p.Phi = Phi; p.ldf = ldf; p.Gc = gc;
            p.Th = Th; p.S = (int)S; p.F = (int)F;
            p.g0 = g0; p.G = G; p.values = values;
            p.part_val = part_val; p.part_idx = part_idx;
            p.nparts = nparts; p.part0 = (int)(g0 / BM);
            dim3 tiles((gc + BM - 1) / BM, (unsigned)((S + BN - 1) / BN));
            rff_gemm_argmax_kernel<<<tiles, GEMM_THREADS, GEMM_SMEM, s>>>(p);
            IPP_LAUNCH_CHECK();
        }
        rff_argmax_finalize_kernel<<<(unsigned)S, 256, 0, s>>>(part_val, part_idx, nparts, max_val, arg_max);
        IPP_LAUNCH_CHECK();
        return IPP_OK;
    }
    const int nblocks = (int)((G + 255) / 256);
    double* part_val = work;
    int64_t* part_idx = reinterpret_cast<int64_t*>(work + (int64_t)S * nblocks);
    if (shared_w) {
        if (D == 2) rff_eval_kernel<2, 32><<<nblocks, 256, 0, s>>>(W, b, theta, (int)F, (int)S, 1, grid, G, values, part_val, part_idx);
        else rff_eval_kernel<3, 32><<<nblocks, 256, 0, s>>>(W, b, theta, (int)F, (int)S, 1, grid, G, values, part_val, part_idx);
    } else {
        if (D == 2) rff_eval_kernel<2, 4><<<nblocks, 256, 0, s>>>(W, b, theta, (int)F, (int)S, 0, grid, G, values, part_val, part_idx);
        else rff_eval_kernel<3, 4><<<nblocks, 256, 0, s>>>(W, b, theta, (int)F, (int)S, 0, grid, G, values, part_val, part_idx);
    }
    IPP_LAUNCH_CHECK();
    rff_argmax_finalize_kernel<<<(unsigned)S, 256, 0, s>>>(part_val, part_idx, nblocks, max_val, arg_max);
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

extern "C" int ipp_rff_features(const double* W, const double* b, int64_t F, int D, const double* X, int64_t N,
                                double scale, double* Z, int64_t ldz, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    IPP_CHECK_ARG(W && b && X && Z && F >= 1 && N >= 1 && ldz >= N && F <= 65535, "bad arguments");
    IPP_CHECK_ARG(D == 2 || D == 3, "dimension must be 2 or 3");
    dim3 grid((unsigned)((N + 255) / 256), (unsigned)F);
    if (D == 2) rff_features_kernel<2><<<grid, 256, 0, s>>>(W, b, F, X, N, scale, Z, ldz);
    else rff_features_kernel<3><<<grid, 256, 0, s>>>(W, b, F, X, N, scale, Z, ldz);
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}
