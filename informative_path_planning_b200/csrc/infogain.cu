// a10: info_gain (aq_library.py:34-80) for a batch of paths, in Schur-complement form.
//   reference:  2 pi e * ( logdet(I + v K([X; q])) - logdet(I + v K(X)) )        two (N+k)^3 LUs per call
//   here:       logdet(I + v K([X;q])) - logdet(I + v K(X)) = logdet S_p,
//               S_p = I_k + v Kqq - (v K_qX) (I + v K_XX)^-1 (v K_Xq)            one factorisation per belief
// With K = variance * e(.,.) and v = variance, v K is an RBF with variance v^2 (`kern_v2`), so the
// a1/a2 kernels are reused unchanged: binv = (I + vK)^-1 comes from ipp_gp_factor(kern_v2, diag_add = 1).
#include "internal.cuh"

namespace ipp {

constexpr int IG_KMAX = 32;

// one warp per path: S_p (k x k) from the two N-long panels, Cholesky in shared memory, logdet
template <int D>
__global__ void __launch_bounds__(32) info_gain_schur_kernel(RbfDev kern, const double* __restrict__ Xq,
                                                             const int64_t* __restrict__ offsets,
                                                             const double* __restrict__ Ct, const double* __restrict__ Gm,
                                                             int64_t ldn, int64_t N, double* __restrict__ logdet_out,
                                                             int32_t* info) {
    __shared__ double a[IG_KMAX][IG_KMAX + 1];
    const int64_t p = blockIdx.x;
    const int64_t b = offsets[p];
    const int k = (int)(offsets[p + 1] - b);
    const int lane = threadIdx.x;
    if (k <= 0) { if (lane == 0) logdet_out[p] = 0.0; return; }
    if (k > IG_KMAX) {          // the caller routes such paths through the blocked factorisation; reaching here is an error
        if (lane == 0) { logdet_out[p] = nan(""); if (info) atomicMax(info, 0x7fffffff); }
        return;
    }
    for (int r = 0; r < k; ++r) {
        for (int c = 0; c <= r; ++c) {
            double s = 0.0;
            const double* gr = Gm + (b + r) * ldn;
            const double* cc = Ct + (b + c) * ldn;
            for (int64_t n = lane; n < N; n += 32) s = fma(gr[n], cc[n], s);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) {
                const double kqq = rbf_eval<D>(kern, Xq + (b + r) * D, Xq + (b + c) * D);
                a[r][c] = (r == c ? 1.0 : 0.0) + kqq - s;
            }
        }
    }
    __syncwarp();
    int bad = 0;
    for (int j = 0; j < k; ++j) {
        if (lane == 0) {
            double d = a[j][j];
            if (!(d > 0.0)) { bad = j + 1; d = 1.0; }
            a[j][j] = sqrt(d);
        }
        __syncwarp();
        if (lane > j && lane < k) a[lane][j] /= a[j][j];
        __syncwarp();
        if (lane > j && lane < k)
            for (int c = j + 1; c <= lane; ++c) a[lane][c] = fma(-a[lane][j], a[c][j], a[lane][c]);
        __syncwarp();
    }
    if (lane == 0) {
        double s = 0.0;
        for (int j = 0; j < k; ++j) s += log(a[j][j]);
        logdet_out[p] = bad ? nan("") : 2.0 * s;
        if (bad && info) atomicMax(info, (int)(p + 1));
    }
}

}  // namespace ipp

using namespace ipp;

extern "C" int64_t ipp_info_gain_workspace(int64_t N, int64_t M) {
    return 2 * round_up(M > 0 ? M : 1, 2) * round_up(N > 0 ? N : 1, 16) + 16;
}

extern "C" int ipp_info_gain_paths(const ipp_rbf* kern_v2, const double* X, int64_t N, const double* binv, int64_t ldb,
                                   const double* Xq, const int64_t* offsets, int64_t P, int64_t M,
                                   double* logdet_out, int32_t* info, double* work, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    IPP_CHECK_ARG(kern_v2 && X && binv && Xq && offsets && logdet_out && work, "null pointer");
    IPP_CHECK_ARG(N >= 1 && M >= 1 && P >= 1, "bad shape");
    IPP_CHECK_ARG(kern_v2->dim == 2 || kern_v2->dim == 3, "dimension must be 2 or 3");
    const int64_t ldn = round_up(N, 16);
    double* Ct = work;
    double* Gm = work + round_up(M, 2) * ldn;
    if (info) IPP_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), s));
    IPP_TRY(launch_rbf_queries(*kern_v2, X, N, nullptr, Xq, M, Ct, ldn, nullptr, s));
    GemmArgs g{};
    g.A = Ct; g.lda = ldn; g.B = binv; g.ldb = ldb; g.C = Gm; g.ldc = ldn;
    g.M = (int)M; g.N = (int)N; g.K = (int)N; g.alpha = 1.0; g.beta = 0.0;
    IPP_TRY(launch_gemm(g, s));
    const RbfDev k = to_dev(*kern_v2);
    if (kern_v2->dim == 2) info_gain_schur_kernel<2><<<(unsigned)P, 32, 0, s>>>(k, Xq, offsets, Ct, Gm, ldn, N, logdet_out, info);
    else info_gain_schur_kernel<3><<<(unsigned)P, 32, 0, s>>>(k, Xq, offsets, Ct, Gm, ldn, N, logdet_out, info);
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}
