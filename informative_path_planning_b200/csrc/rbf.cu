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
// One warp owns QPW queries held in registers and sweeps the observations lane-strided, so every
// store is a coalesced 256-byte row segment; X and alpha are tiny (N*(D+1)*8 B) and stay in L1/L2.
template <int D>
__global__ void __launch_bounds__(256) rbf_queries_kernel(RbfDev k, const double* __restrict__ X, int64_t N,
                                                          const double* __restrict__ alpha,
                                                          const double* __restrict__ Xq, int64_t M,
                                                          double* __restrict__ Kxt, int64_t ldk,
                                                          double* __restrict__ mean) {
    constexpr int QPW = 4;
    const int lane = threadIdx.x & 31;
    const int64_t m0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * QPW;
    if (m0 >= M) return;
    double q[QPW][D], acc[QPW];
#pragma unroll
    for (int i = 0; i < QPW; ++i) {
        const int64_t m = (m0 + i < M) ? (m0 + i) : (M - 1);
#pragma unroll
        for (int d = 0; d < D; ++d) q[i][d] = Xq[m * D + d];
        acc[i] = 0.0;
    }
    for (int64_t n = lane; n < N; n += 32) {
        double x[D];
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = X[n * D + d];
        const double a = alpha ? alpha[n] : 0.0;
#pragma unroll
        for (int i = 0; i < QPW; ++i) {
            const double kv = rbf_eval<D>(k, q[i], x);
            if (m0 + i < M) Kxt[(m0 + i) * ldk + n] = kv;
            acc[i] = fma(kv, a, acc[i]);
        }
    }
    if (mean) {
#pragma unroll
        for (int i = 0; i < QPW; ++i) {
            double s = acc[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0 && m0 + i < M) mean[m0 + i] = s;
        }
    }
}

int launch_rbf_queries(const ipp_rbf& kern, const double* X, int64_t N, const double* alpha,
                       const double* Xq, int64_t M, double* Kxt, int64_t ldk, double* mean, cudaStream_t s) {
    if (M <= 0) return IPP_OK;
    IPP_CHECK_ARG(kern.dim == 2 || kern.dim == 3, "dimension must be 2 or 3");
    const RbfDev k = to_dev(kern);
    const int64_t per_block = 8 * 4;
    const unsigned grid = (unsigned)((M + per_block - 1) / per_block);
    if (kern.dim == 2) rbf_queries_kernel<2><<<grid, 256, 0, s>>>(k, X, N, alpha, Xq, M, Kxt, ldk, mean);
    else rbf_queries_kernel<3><<<grid, 256, 0, s>>>(k, X, N, alpha, Xq, M, Kxt, ldk, mean);
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

}  // namespace ipp

extern "C" int ipp_rbf_cross(const ipp_rbf* kern, const double* X1, int64_t n1, const double* X2, int64_t n2,
                             double* out, int64_t ldo, void* stream) {
    IPP_CHECK_ARG(kern && X1 && X2 && out, "null pointer");
    IPP_CHECK_ARG(ldo >= n2, "ldo < n2");
    return ipp::launch_rbf_cross(*kern, X1, n1, X2, n2, out, ldo, static_cast<cudaStream_t>(stream));
}
