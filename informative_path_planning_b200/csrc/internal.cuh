// Internal declarations shared by the translation units of libipp_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/ipp_b200.h"

namespace ipp {

void set_error(const char* fmt, ...);
int  sm_count();
void profile_mark(int which, cudaStream_t s);      // CUDA-event bracket of the dominant kernel (ipp_profile_*)

#define IPP_CHECK_ARG(cond, msg)                                                            \
    do { if (!(cond)) { ::ipp::set_error("%s: %s", __func__, msg); return IPP_E_INVALID; } } while (0)

#define IPP_CUDA(call)                                                                      \
    do { cudaError_t e__ = (call); if (e__ != cudaSuccess) {                                \
        ::ipp::set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e__));    \
        return IPP_E_CUDA; } } while (0)

#define IPP_LAUNCH_CHECK()                                                                  \
    do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) {                    \
        ::ipp::set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(e__));\
        return IPP_E_CUDA; } } while (0)

#define IPP_TRY(call) do { int rc__ = (call); if (rc__ != IPP_OK) return rc__; } while (0)

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device (per-context) setting: remember it per device so
// that one process driving several GPUs configures every one of them (a plain function-static flag would not).
#define IPP_OPT_IN_SMEM(kernel, bytes)                                                                   \
    do {                                                                                                 \
        static bool done__[64] = {};                                                                     \
        int dev__ = 0;                                                                                   \
        IPP_CUDA(cudaGetDevice(&dev__));                                                                 \
        if (dev__ < 0 || dev__ >= 64 || !done__[dev__]) {                                                \
            IPP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            if (dev__ >= 0 && dev__ < 64) done__[dev__] = true;                                          \
        }                                                                                                \
    } while (0)

// ----- generic FP64 DMMA GEMM:  C (and/or Ct) = alpha * A[M][K] * B[N][K]^T + beta * C
struct GemmArgs {
    const double* A; const double* B; double* C; double* Ct;
    int64_t lda, ldb, ldc, ldct;
    int M, N, K;
    double alpha, beta;
    int lower_only;      // skip output tiles that lie strictly above the block diagonal
    int k_from_max;      // contraction starts at max(m0, n0) (product of two upper-triangular-stored factors)
    int k_to_min;        // contraction ends at min(m0, n0) + tile (product with lower-triangular operands)
    int batch;           // > 1: blockIdx.z = problem index; operand / result z starts sA, sB, sC, sCt elements further on
    int64_t sA, sB, sC, sCt;
};
int launch_gemm(const GemmArgs& g, cudaStream_t stream);

// ----- kernels used across files
int launch_rbf_cross(const ipp_rbf& k, const double* X1, int64_t n1, const double* X2, int64_t n2,
                     double* out, int64_t ldo, cudaStream_t s);
// Kxt[m][n] = k(q_m, x_n) for a chunk of queries, plus mean[m] = sum_n Kxt[m][n] alpha[n] (alpha may be NULL)
int launch_rbf_queries(const ipp_rbf& k, const double* X, int64_t N, const double* alpha,
                       const double* Xq, int64_t M, double* Kxt, int64_t ldk, double* mean, cudaStream_t s);
int launch_gemv(const double* A, int64_t lda, int64_t rows, int64_t cols, const double* x, double* y, cudaStream_t s);
int potrf_lower(double* A, int64_t N, int64_t lda, double* dinv /* [ceil(N/64)][64*64] */, double* logdet,
                int32_t* info, cudaStream_t s);
// pdinv of a dense SPD matrix in device memory (factor.cu): winv = A^-1, linv (may be NULL) = chol(A)^-1; A is overwritten
// by its factor; work: ipp_pdinv_workspace(N) doubles
int pdinv_device(double* A, int64_t N, int64_t lda, double* winv, double* linv, int64_t ldw,
                 double* logdet, int32_t* info, double* work, cudaStream_t s);

struct RbfDev {
    int dim;
    double variance;
    double il[IPP_MAX_DIM];
};

static inline RbfDev to_dev(const ipp_rbf& k) {
    RbfDev d;
    d.dim = k.dim;
    d.variance = k.variance;
    for (int i = 0; i < IPP_MAX_DIM; ++i) d.il[i] = i < k.dim ? k.inv_lengthscale[i] : 0.0;
    return d;
}

template <int D>
__device__ __forceinline__ double rbf_eval(const RbfDev& k, const double* a, const double* b) {
    double r2 = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const double u = (a[d] - b[d]) * k.il[d];
        r2 = fma(u, u, r2);
    }
    return k.variance * exp(-0.5 * r2);
}


// exp(x), x <= 0, branch-free (several calls interleave in one thread; libdevice's exp is ~2x the instructions, half of
// them outside the FP64 pipe, with a branch per call): Cody-Waite reduction by ln 2 and the degree-13
// Taylor polynomial on |f| <= ln 2 / 2 (truncation 4e-18 relative), result scaled by 2^n through the exponent bits.
// The constants sit in constant memory: FP64 instructions read a constant-bank operand directly, whereas 64-bit
// immediates are rebuilt in (uniform) registers inside the loop -- 6 extra instructions per call in rbf_queries_kernel.
// No clamp of the argument: for x < -700 the polynomial may see garbage, the final select discards it.
static __constant__ double EXPC[18] = {
    1.4426950408889634074, 6755399441055744.0, -6.93147180369123816490e-01, -1.90821492927058770002e-10,
    1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07, 2.7557319223985893e-06,
    2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03, 8.333333333333333e-03, 4.1666666666666664e-02,
    1.6666666666666666e-01, 0.5, 1.0, -700.0};
__device__ __forceinline__ double exp_neg_exact(double x) {
    const double tt = fma(x, EXPC[0], EXPC[1]);                               // n = rint(x log2 e) in the low word
    const int n = __double2loint(tt);
    const double nf = tt - EXPC[1];
    double f = fma(nf, EXPC[2], x);
    f = fma(nf, EXPC[3], f);
    double p = EXPC[4];                                                       // 1/13!
#pragma unroll
    for (int i = 5; i <= 16; ++i) p = fma(p, f, EXPC[i]);                     // ... 1/2!, 1
    p = fma(p, f, EXPC[16]);
    const double r = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));   // p in [0.70, 1.42]
    return x < EXPC[17] ? 0.0 : r;
}

template <int D>
__device__ __forceinline__ double rbf_eval_bf(const RbfDev& k, const double* a, const double* b) {
    double r2 = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const double u = (a[d] - b[d]) * k.il[d];
        r2 = fma(u, u, r2);
    }
    return k.variance * exp_neg_exact(-0.5 * r2);
}

// exp(x) for x <= 0 to ~2^-36 relative (degree-9 Taylor after Cody-Waite reduction): used by the tensor-core modes
// (3xTF32: operands rounded to 22 bits; sliced INT8: 42-bit fixed point, tolerance 1e-6), where libdevice's < 1 ulp exp
// (about twice the FP64-pipe instructions) buys nothing.  The FP64 DMMA mode keeps the exact exp.
__device__ __forceinline__ double exp_neg_fast(double x) {
    if (x < -700.0) return 0.0;
    const double t = fma(x, 1.4426950408889634074, 6755399441055744.0);      // n = rint(x log2 e) in the low word
    const int n = __double2loint(t);
    const double nf = t - 6755399441055744.0;
    double f = fma(nf, -6.93147180369123816490e-01, x);
    f = fma(nf, -1.90821492927058770002e-10, f);
    double p = 2.7557319223985893e-06;                                       // 1/9!
    p = fma(p, f, 2.4801587301587302e-05);
    p = fma(p, f, 1.9841269841269841e-04);
    p = fma(p, f, 1.3888888888888889e-03);
    p = fma(p, f, 8.3333333333333332e-03);
    p = fma(p, f, 4.1666666666666664e-02);
    p = fma(p, f, 1.6666666666666666e-01);
    p = fma(p, f, 0.5);
    p = fma(p, f, 1.0);
    p = fma(p, f, 1.0);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));   // p in [0.70, 1.42]: exact scaling by 2^n
}

template <int D>
__device__ __forceinline__ double rbf_eval_fast(const RbfDev& k, const double* a, const double* b) {
    double r2 = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const double u = (a[d] - b[d]) * k.il[d];
        r2 = fma(u, u, r2);
    }
    return k.variance * exp_neg_fast(-0.5 * r2);
}


// 1 / sqrt(d), d > 0 and normal, to ~1 ulp in four dependent FP64 operations after the hardware approximation
// (MUFU.RSQ64H, 2^-22.9): one third-order step  y' = y (1 + e/2 + 3 e^2 / 8),  e = 1 - d y^2, leaves e^3 ~ 2^-68.
// libdevice's rsqrt() is ~20 dependent instructions; on the pivot chains of the small Cholesky factorisations
// (rollout conditioning, diagonal blocks of the refresh) that latency is the critical path.
__device__ __forceinline__ double rsqrt_fast(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double t = d * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double ye = y * e;
    return fma(ye, p, y);
}

// ---- rollout.cu <-> tree.cu: one rollout's launch sequence with the points passed by value and a host-visible
// completion word (see rollout_launch in rollout.cu)
constexpr int RO_INLINE_POINTS = 96;                                          // 96 * 4 doubles = 3 KB of kernel parameters
struct RolloutInline { double v[RO_INLINE_POINTS * (IPP_MAX_DIM + 1)]; };     // [M][D] points, then [M] observations
int rollout_launch(const ipp_rbf* kern, const double* X, int64_t N, const double* alpha,
                   const double* winv, int64_t ldw, double noise_pred, double diag_add,
                   int kind, const double* params_host, int n_params, const double* maxes_dev, int64_t S,
                   double* pts, double* zobs, const RolloutInline* inline_pts, const int32_t* offsets_host,
                   const int32_t* reps_host, int L, double* reward, int32_t* info,
                   unsigned int* done, unsigned int done_seq, double* work, cudaStream_t stream,
                   double* tagged_out, unsigned long long tag, bool* used_tagged);
// tagged_out != NULL (host-mapped, 16-byte aligned, 2 (L + 1) doubles) and tag != 0: when the one-launch form runs, the
// results are written there as {value, tag} records -- L rewards, then info -- and *used_tagged is set; `done` is then not
// signalled.  Tags must never repeat for a given buffer.

}  // namespace ipp
