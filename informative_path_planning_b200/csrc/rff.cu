// a11: random-Fourier-feature posterior function samples evaluated on a grid, with a per-sample
// arg-max (general_target aq_library.py:138-139; grid stage of global_maximization :475-479), and
// the feature matrix Z = scale * cos(W X^T + b) (:171).
//   f_s(x_g) = sum_f theta[s][f] cos(W_s[f] . x_g + b_s[f])
// shared_w: the cosine of a (g, f) pair is evaluated once and reused by SC = 32 samples held in
// registers; per-sample W (the reference's sampling scheme) has no reuse and is cos-bound.
#include "internal.cuh"

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

}  // namespace ipp

using namespace ipp;

extern "C" int64_t ipp_rff_workspace(int64_t S, int64_t G) { return 2 * S * ((G + 255) / 256) + 16; }

extern "C" int ipp_rff_eval_argmax(const double* W, const double* b, const double* theta, int64_t F, int64_t S, int D,
                                   int shared_w, const double* grid, int64_t G, double* values,
                                   double* max_val, int64_t* arg_max, double* work, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    IPP_CHECK_ARG(W && b && theta && grid && max_val && arg_max && work, "null pointer");
    IPP_CHECK_ARG(F >= 1 && S >= 1 && G >= 1 && F < (1 << 30) && S < (1 << 30), "bad shape");
    IPP_CHECK_ARG(D == 2 || D == 3, "dimension must be 2 or 3");
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
