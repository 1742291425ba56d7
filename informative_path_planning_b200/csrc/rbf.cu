// a1: RBF cross-covariance kernels (GPy kern.K restated: variance * exp(-0.5 * |(x - y)/l|^2)).
// FP64 SIMT distance + exp; the difference form (x - y)^2 is used instead of GPy's expanded
// |x|^2 + |y|^2 - 2 x.y (same value to ~1e-14 relative, exact zero on coincident points).
#include "internal.cuh"

namespace ipp {

// out[i][j] = k(X1[i], X2[j]);  thread = column j, block row-slab of ROWS rows staged in shared memory
template <int D>
__global__ void __launch_bounds__(256) rbf_cross_kernel(RbfDev k, const double* __restrict__ X1, int64_t n1,
                                                        const double* __restrict__ X2, int64_t n2,
                                                        double* __restrict__ out, int64_t ldo) {
    constexpr int ROWS = 32;
    __shared__ double xs[ROWS][D];
    const int64_t i0 = (int64_t)blockIdx.y * ROWS;
    for (int idx = threadIdx.x; idx < ROWS * D; idx += blockDim.x) {
        const int64_t r = i0 + idx / D;
        xs[idx / D][idx % D] = r < n1 ? X1[r * D + idx % D] : 0.0;
    }
    __syncthreads();
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n2) return;
    double y[D];
#pragma unroll
    for (int d = 0; d < D; ++d) y[d] = X2[j * D + d];
    const int rows = (int)((n1 - i0) < ROWS ? (n1 - i0) : ROWS);
    for (int r = 0; r < rows; ++r) out[(i0 + r) * ldo + j] = rbf_eval<D>(k, xs[r], y);
}

int launch_rbf_cross(const ipp_rbf& kern, const double* X1, int64_t n1, const double* X2, int64_t n2,
                     double* out, int64_t ldo, cudaStream_t s) {
    if (n1 <= 0 || n2 <= 0) return IPP_OK;
    IPP_CHECK_ARG(kern.dim == 2 || kern.dim == 3, "dimension must be 2 or 3");
    const RbfDev k = to_dev(kern);
    dim3 grid((unsigned)((n2 + 255) / 256), (unsigned)((n1 + 31) / 32));
    if (kern.dim == 2) rbf_cross_kernel<2><<<grid, 256, 0, s>>>(k, X1, n1, X2, n2, out, ldo);
    else rbf_cross_kernel<3><<<grid, 256, 0, s>>>(k, X1, n1, X2, n2, out, ldo);
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

// Query-major cross-covariance for predict: Kxt[m][n] = k(q_m, x_n), mean[m] = sum_n Kxt[m][n] alpha[n].
// A warp owns a contiguous run of `per_warp` queries and sweeps the observations lane-strided for up to QPW of them at
// a time (queries in registers), so every store is a coalesced 256-byte row segment; X and alpha are tiny
// (N*(D+1)*8 B) and stay in L1/L2 -- the sweep fetches the operands of the next observation before it evaluates the
// current one.  The launch is ONE wave: per_warp = ceil(M / resident warps), so every warp does the same number of
// sweeps (a fixed 4 queries per warp ran 2.3 waves at M = 32768: the last one a third full, profiles/r2a_*).
// Values and summation order per query are those of the fixed-QPW form (lane-strided partial sums, xor tree).
constexpr int RBFQ_QPW = 4, RBFQ_CTAS_PER_SM = 3;
constexpr int RBFQ_SMEM_THREADS = 512, RBFQ_SMEM_MAX = 200 * 1024;   // staged form: one block of 16 warps per SM
// one sweep over the observations for Q queries (Q is a compile-time count: no guards inside the loop)
template <int D, int Q, bool SMEM>
__device__ __forceinline__ void rbf_queries_sweep(const RbfDev& k, const double* __restrict__ X, int64_t N,
                                                  const double* __restrict__ alpha, const double* __restrict__ Xq,
                                                  int64_t m0, double* __restrict__ Kxt, int64_t ldk,
                                                  double* __restrict__ mean, int lane) {
    double q[Q][D], acc[Q];
    double* row[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i) {
#pragma unroll
        for (int d = 0; d < D; ++d) q[i][d] = SMEM ? Xq[(m0 + i) * D + d] * k.il[d] : Xq[(m0 + i) * D + d];
        acc[i] = 0.0;
        row[i] = Kxt + (m0 + i) * ldk;
    }
    if constexpr (SMEM) {
        // X and alpha staged in shared memory by the block (alpha == nullptr: zeros were staged)
        for (int64_t n = lane; n < N; n += 32) {
            double x[D];
#pragma unroll
            for (int d = 0; d < D; ++d) x[d] = X[n * D + d];
            const double a = alpha[n];
#pragma unroll
            for (int i = 0; i < Q; ++i) {
                double r2 = 0.0;                           // both sides were scaled by 1 / lengthscale when they were loaded
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const double u = q[i][d] - x[d];
                    r2 = fma(u, u, r2);
                }
                const double kv = k.variance * exp_neg_exact(-0.5 * r2);
                row[i][n] = kv;
                acc[i] = fma(kv, a, acc[i]);
            }
        }
    } else {
    // the operands of observation n + 32 are fetched before observation n is evaluated (under the kernel's own store
    // traffic a global load takes longer than a warp's Q exponentials)
    double xn[D], an = 0.0;
    if (lane < N) {
#pragma unroll
        for (int d = 0; d < D; ++d) xn[d] = X[(int64_t)lane * D + d];
        if (alpha) an = alpha[lane];
    }
    for (int64_t n = lane; n < N; n += 32) {
        double x[D];
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = xn[d];
        const double a = an;
        const int64_t nn = (n + 32 < N) ? (n + 32) : n;  // (the last step re-reads its own operands)
#pragma unroll
        for (int d = 0; d < D; ++d) xn[d] = X[nn * D + d];
        if (alpha) an = alpha[nn];
#pragma unroll
        for (int i = 0; i < Q; ++i) {
            const double kv = rbf_eval_bf<D>(k, q[i], x);
            row[i][n] = kv;
            acc[i] = fma(kv, a, acc[i]);
        }
    }
    }
    if (mean) {
#pragma unroll
        for (int i = 0; i < Q; ++i) {
            double s = acc[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) mean[m0 + i] = s;
        }
    }
}

template <int D, bool SMEM>
__global__ void __launch_bounds__(SMEM ? RBFQ_SMEM_THREADS : 256, SMEM ? 1 : RBFQ_CTAS_PER_SM)
rbf_queries_kernel(RbfDev k, const double* __restrict__ X, int64_t N, const double* __restrict__ alpha,
                   const double* __restrict__ Xq, int64_t M, int64_t per_warp, double* __restrict__ Kxt, int64_t ldk,
                   double* __restrict__ mean) {
    constexpr int QPW = RBFQ_QPW;
    extern __shared__ double rbfq_smem[];                  // SMEM: [N][D] observations, [N] alpha
    const int lane = threadIdx.x & 31;
    if constexpr (SMEM) {
        for (int64_t i = threadIdx.x; i < N * D; i += blockDim.x) {
            const int d = (int)(i % D);                    // (selects, not an indexed read of the kernel parameters)
            rbfq_smem[i] = X[i] * (d == 0 ? k.il[0] : (d == 1 || D == 2) ? k.il[1] : k.il[D - 1]);
        }
        for (int64_t i = threadIdx.x; i < N; i += blockDim.x) rbfq_smem[N * D + i] = alpha ? alpha[i] : 0.0;
        __syncthreads();
        X = rbfq_smem;
        alpha = rbfq_smem + N * D;
    }
    const int64_t mb = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * per_warp;
    const int64_t me = (mb + per_warp < M) ? (mb + per_warp) : M;
    int64_t m0 = mb;
    for (; m0 + QPW <= me; m0 += QPW) rbf_queries_sweep<D, QPW, SMEM>(k, X, N, alpha, Xq, m0, Kxt, ldk, mean, lane);
    switch ((int)(me - m0)) {                              // warp-uniform remainder
        case 3: rbf_queries_sweep<D, 3, SMEM>(k, X, N, alpha, Xq, m0, Kxt, ldk, mean, lane); break;
        case 2: rbf_queries_sweep<D, 2, SMEM>(k, X, N, alpha, Xq, m0, Kxt, ldk, mean, lane); break;
        case 1: rbf_queries_sweep<D, 1, SMEM>(k, X, N, alpha, Xq, m0, Kxt, ldk, mean, lane); break;
        default: break;
    }
}

template <int D>
static int rbf_queries_launch(const RbfDev& k, const double* X, int64_t N, const double* alpha, const double* Xq, int64_t M,
                              double* Kxt, int64_t ldk, double* mean, cudaStream_t s) {
    static const int mode = [] { const char* e = getenv("IPP_RBFQ_MODE"); return e ? atoi(e) : 2; }();   // 0: round 1's grid, 1: one wave, 2: staged
    const int64_t smem = N * (D + 1) * (int64_t)sizeof(double);
    if (mode == 2 && smem <= RBFQ_SMEM_MAX) {
        const int wpb = RBFQ_SMEM_THREADS / 32;
        const int64_t resident_warps = (int64_t)sm_count() * wpb;
        const int64_t per_warp = (M + resident_warps - 1) / resident_warps;
        const int64_t warps = (M + per_warp - 1) / per_warp;
        IPP_OPT_IN_SMEM((rbf_queries_kernel<D, true>), RBFQ_SMEM_MAX);
        rbf_queries_kernel<D, true><<<(unsigned)((warps + wpb - 1) / wpb), RBFQ_SMEM_THREADS, (size_t)smem, s>>>(
            k, X, N, alpha, Xq, M, per_warp, Kxt, ldk, mean);
    } else {
        const int64_t resident_warps = (int64_t)sm_count() * RBFQ_CTAS_PER_SM * 8;
        const int64_t per_warp = mode == 0 ? RBFQ_QPW : (M + resident_warps - 1) / resident_warps;
        const int64_t warps = (M + per_warp - 1) / per_warp;
        rbf_queries_kernel<D, false><<<(unsigned)((warps + 7) / 8), 256, 0, s>>>(k, X, N, alpha, Xq, M, per_warp, Kxt, ldk, mean);
    }
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

int launch_rbf_queries(const ipp_rbf& kern, const double* X, int64_t N, const double* alpha,
                       const double* Xq, int64_t M, double* Kxt, int64_t ldk, double* mean, cudaStream_t s) {
    if (M <= 0) return IPP_OK;
    IPP_CHECK_ARG(kern.dim == 2 || kern.dim == 3, "dimension must be 2 or 3");
    const RbfDev k = to_dev(kern);
    return kern.dim == 2 ? rbf_queries_launch<2>(k, X, N, alpha, Xq, M, Kxt, ldk, mean, s)
                         : rbf_queries_launch<3>(k, X, N, alpha, Xq, M, Kxt, ldk, mean, s);
}

}  // namespace ipp

extern "C" int ipp_rbf_cross(const ipp_rbf* kern, const double* X1, int64_t n1, const double* X2, int64_t n2,
                             double* out, int64_t ldo, void* stream) {
    IPP_CHECK_ARG(kern && X1 && X2 && out, "null pointer");
    IPP_CHECK_ARG(ldo >= n2, "ldo < n2");
    return ipp::launch_rbf_cross(*kern, X1, n1, X2, n2, out, ldo, static_cast<cudaStream_t>(stream));
}
