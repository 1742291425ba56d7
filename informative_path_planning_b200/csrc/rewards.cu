// a7 / a8 / a9: acquisition scoring applied to (mean, var) of every sample point, reduced per path.
//   UCB  aq_library.py:106-114   sum mu + sqrt(beta_t) * sum |var|        (variance, not std: quirk Q1)
//   EI   aq_library.py:569-580   x Phi(z) + sum|var| phi(z),  x = sum(mu - eta), z = x / sum|var|
//   MES  aq_library.py:309-337   2/S * sum_i sum_p u,  u = g phi(g) / (2 Phi(g)) - log Phi(g),  g = (y*_i - mu)/var  (Q2)
// Plain IEEE FP64 (no fast-math): Inf/NaN propagate exactly as numpy's do.
#include "internal.cuh"

namespace ipp {

// scipy.special.ndtr (cephes ndtr.c): the same branches, on CUDA's erf/erfc
__device__ __forceinline__ double norm_cdf(double a) {
    const double x = a * 0.70710678118654752440;
    const double z = fabs(x);
    if (z < 0.70710678118654752440) return 0.5 + 0.5 * erf(x);
    double y = 0.5 * erfc(z);
    if (x > 0) y = 1.0 - y;
    return y;
}
// scipy.stats.norm.pdf: exp(-x**2/2) / sqrt(2 pi)
__device__ __forceinline__ double norm_pdf(double x) { return exp(-(x * x) / 2.0) / 2.50662827463100050242; }

__device__ __forceinline__ double mes_utility(double ymax, double mu, double var) {
    const double gamma = (ymax - mu) / var;
    const double pdf = norm_pdf(gamma), cdf = norm_cdf(gamma);
    return gamma * pdf / (2.0 * cdf) - log(cdf);
}

// point_val[m] = sum_i u(y*_i, mu_m, var_m)
__global__ void mes_points_kernel(const double* __restrict__ mean, const double* __restrict__ var, int64_t M,
                                  const double* __restrict__ maxes, int64_t S, double* __restrict__ point_val) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const double mu = mean[m], v = var[m];
    double s = 0.0;
    for (int64_t i = 0; i < S; ++i) s += mes_utility(maxes[i], mu, v);
    point_val[m] = s;
}

__global__ void ucb_points_kernel(const double* __restrict__ mean, const double* __restrict__ var, int64_t M,
                                  double sqrt_beta, double* __restrict__ point_val) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m < M) point_val[m] = mean[m] + sqrt_beta * fabs(var[m]);
}

// one warp per path
__global__ void __launch_bounds__(256) reward_paths_kernel(int kind, const double* __restrict__ mean,
                                                           const double* __restrict__ var,
                                                           const int64_t* __restrict__ offsets, int64_t P, double p0,
                                                           const double* __restrict__ point_val, double inv_S,
                                                           double* __restrict__ reward) {
    const int64_t p = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (p >= P) return;
    const int64_t b = offsets[p], e = offsets[p + 1];
    double s0 = 0.0, s1 = 0.0;
    for (int64_t m = b + lane; m < e; m += 32) {
        if (kind == IPP_REWARD_MES) {
            s0 += point_val[m];
        } else {
            s0 += (kind == IPP_REWARD_EI) ? (mean[m] - p0) : mean[m];
            s1 += fabs(var[m]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (lane != 0) return;
    double r;
    if (kind == IPP_REWARD_UCB) {
        r = s0 + p0 * s1;
    } else if (kind == IPP_REWARD_EI) {
        const double z = s0 / s1;
        const double big_phi = 0.5 * (1.0 + erf(z / 1.41421356237309504880));
        const double small_phi = 1.0 / 2.50662827463100050242 * exp(-(z * z) / 2.0);
        r = s0 * big_phi + s1 * small_phi;
    } else {
        r = 2.0 * s0 * inv_S;
    }
    reward[p] = r;
}

}  // namespace ipp

using namespace ipp;

extern "C" int ipp_reward_paths(int kind, const double* mean, const double* var, const int64_t* offsets, int64_t P,
                                const double* params_host, int n_params, const double* maxes_dev, int64_t S,
                                double* reward, double* point_out, int64_t M, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    IPP_CHECK_ARG(mean && var && M >= 0 && P >= 0, "bad arguments");
    IPP_CHECK_ARG(kind == IPP_REWARD_UCB || kind == IPP_REWARD_EI || kind == IPP_REWARD_MES, "unknown reward kind");
    double p0 = 0.0, inv_S = 0.0;
    if (kind == IPP_REWARD_MES) {
        IPP_CHECK_ARG(maxes_dev && S >= 1, "MES needs the sampled maxima");
        IPP_CHECK_ARG(point_out, "MES needs point_out[M] as scratch");
        inv_S = 1.0 / (double)S;
        if (M > 0) {
            mes_points_kernel<<<(unsigned)((M + 127) / 128), 128, 0, s>>>(mean, var, M, maxes_dev, S, point_out);
            IPP_LAUNCH_CHECK();
        }
    } else {
        IPP_CHECK_ARG(params_host && n_params >= 1, "need params[0]");
        p0 = params_host[0];
        if (kind == IPP_REWARD_UCB && point_out && M > 0) {
            ucb_points_kernel<<<(unsigned)((M + 255) / 256), 256, 0, s>>>(mean, var, M, p0, point_out);
            IPP_LAUNCH_CHECK();
        }
    }
    if (P > 0) {
        IPP_CHECK_ARG(offsets && reward, "need offsets and reward");
        reward_paths_kernel<<<(unsigned)((P + 7) / 8), 256, 0, s>>>(kind, mean, var, offsets, P, p0, point_out, inv_S, reward);
        IPP_LAUNCH_CHECK();
    }
    return IPP_OK;
}
