// Generic FP64 DMMA GEMM (NT form) used by the factorisation, the full-covariance predict,
// info_gain's Schur blocks and exposed as ipp_dgemm_nt.
#include "internal.cuh"
#include "dgemm_dmma.cuh"

namespace ipp {

__global__ void __launch_bounds__(GEMM_THREADS, 1) dgemm_nt_kernel(GemmArgs p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (blockIdx.z) {                                    // batched: the same shape on operands a fixed stride apart
        p.A += (int64_t)blockIdx.z * p.sA;
        p.B += (int64_t)blockIdx.z * p.sB;
        if (p.C) p.C += (int64_t)blockIdx.z * p.sC;
        if (p.Ct) p.Ct += (int64_t)blockIdx.z * p.sCt;
    }
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    if (p.lower_only && n0 >= m0 + BM) return;
    const uint32_t smem_base = smem_u32(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp / Cfg::WARPS_N) * WM, wn0 = (warp % Cfg::WARPS_N) * WN;

    int kbeg = 0, kend = p.K;
    if (p.k_from_max) kbeg = ((m0 > n0 ? m0 : n0) / BK) * BK;
    if (p.k_to_min) { const int lim = (m0 < n0 ? m0 : n0) + BM; kend = lim < p.K ? lim : p.K; }

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    if (kend > kbeg)
        gemm_mainloop(smem_base, p.A, p.lda, m0, p.M, p.B, p.ldb, n0, p.N, kbeg, kend, acc, tid, wm0, wn0, g, t);

#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int row = m0 + wm0 + 8 * i + g;
        if (row >= p.M) continue;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int col = n0 + wn0 + 8 * j + 2 * t;
            if (col >= p.N) continue;
            double v0 = p.alpha * acc[i][j][0], v1 = p.alpha * acc[i][j][1];
            const bool two = col + 1 < p.N;
            if (p.beta != 0.0) {
                const double* c = p.C + (int64_t)row * p.ldc + col;
                v0 += p.beta * c[0];
                if (two) v1 += p.beta * c[1];
            }
            if (p.C) {
                double* c = p.C + (int64_t)row * p.ldc + col;
                if (two) *reinterpret_cast<double2*>(c) = make_double2(v0, v1);
                else c[0] = v0;
            }
            if (p.Ct) {
                p.Ct[(int64_t)col * p.ldct + row] = v0;
                if (two) p.Ct[(int64_t)(col + 1) * p.ldct + row] = v1;
            }
        }
    }
}

int launch_gemm(const GemmArgs& g, cudaStream_t stream) {
    if (g.M <= 0 || g.N <= 0) return IPP_OK;
    IPP_CHECK_ARG((g.lda % 2 == 0) && (g.ldb % 2 == 0), "lda/ldb must be even (16-byte rows)");
    IPP_CHECK_ARG((reinterpret_cast<uintptr_t>(g.A) % 16 == 0) && (reinterpret_cast<uintptr_t>(g.B) % 16 == 0),
                  "A and B must be 16-byte aligned");
    IPP_CHECK_ARG(g.C == nullptr || ((g.ldc % 2 == 0) && (reinterpret_cast<uintptr_t>(g.C) % 16 == 0)),
                  "C must be 16-byte aligned with even ldc");
    IPP_CHECK_ARG(g.C != nullptr || g.beta == 0.0, "beta != 0 needs C");
    IPP_OPT_IN_SMEM(dgemm_nt_kernel, GEMM_SMEM);
    IPP_CHECK_ARG(g.batch <= 1 || (g.sA % 2 == 0 && g.sB % 2 == 0 && g.sC % 2 == 0), "batch strides must keep 16-byte alignment");
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, g.batch > 1 ? g.batch : 1);
    dgemm_nt_kernel<<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(g);
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

}  // namespace ipp

extern "C" int ipp_dgemm_nt(int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
                            const double* B, int64_t ldb, double beta, double* C, int64_t ldc,
                            double* Ct, int64_t ldct, void* stream) {
    IPP_CHECK_ARG(M >= 0 && N >= 0 && K >= 0 && M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), "bad shape");
    IPP_CHECK_ARG(C != nullptr || Ct != nullptr, "need an output");
    ipp::GemmArgs g{};
    g.A = A; g.B = B; g.C = C; g.Ct = Ct;
    g.lda = lda; g.ldb = ldb; g.ldc = ldc; g.ldct = ldct;
    g.M = (int)M; g.N = (int)N; g.K = (int)K;
    g.alpha = alpha; g.beta = beta;
    return ipp::launch_gemm(g, static_cast<cudaStream_t>(stream));
}
