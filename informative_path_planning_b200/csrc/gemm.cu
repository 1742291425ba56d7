// Generic FP64 DMMA GEMM (NT form) used by the factorisation, the full-covariance predict,
// info_gain's Schur blocks and exposed as ipp_dgemm_nt.
#include "internal.cuh"
#include "dgemm_dmma.cuh"
#include <stdlib.h>

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

// ---- 64 x 64 tiles, four warps, for problems that do not fill the GPU with 128 x 128 tiles: the panel solve and the
// in-panel update of the blocked Cholesky (rows x 64 x 64 and rows x <= 192 x 64: at most 64 big tiles, each at least
// 8 us of DMMAs on its one SM), the small products of the triangular-inverse doubling, the F x F work of max-value
// sampling.  Same operand layout, swizzle, k-permutation and GemmArgs semantics as the big kernel (tile-relative rules
// use this kernel's tile size); three stages of 32 KB, two CTAs per SM.
constexpr int SB = 64, S_THREADS = 128, S_STAGES = 3;
constexpr int S_TILE_BYTES = SB * ROW_BYTES, S_STAGE_BYTES = 2 * S_TILE_BYTES, S_SMEM = S_STAGES * S_STAGE_BYTES;
constexpr int S_LOADS = (SB * CHUNKS_PER_ROW) / S_THREADS, S_RSTEP = S_THREADS / CHUNKS_PER_ROW;

struct SmallLoader {
    const double* base; int64_t rstep; uint32_t dst; int kcol; unsigned valid;
    __device__ __forceinline__ void init(const double* G, int64_t ld, int row0, int nrows, int tid) {
        const int r = tid / CHUNKS_PER_ROW, ch = tid % CHUNKS_PER_ROW;
        kcol = ch * 2;
        base = G + (int64_t)(row0 + r) * ld + kcol;
        rstep = (int64_t)S_RSTEP * ld;
        dst = r * ROW_BYTES + ((ch ^ ((r & 1) << 2)) << 4);      // S_RSTEP is even: (row & 1) == (r & 1) for all i
        valid = 0;
#pragma unroll
        for (int i = 0; i < S_LOADS; ++i)
            if (row0 + r + i * S_RSTEP < nrows) valid |= 1u << i;
    }
    __device__ __forceinline__ void load(uint32_t smem_tile, int k0, int kend) const {
        const int rem = kend - (k0 + kcol);
        const int bytes = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
        const double* src = base + k0;
#pragma unroll
        for (int i = 0; i < S_LOADS; ++i) {
            const bool ok = ((valid >> i) & 1u) && bytes > 0;
            cp_async16(smem_tile + dst + i * (S_RSTEP * ROW_BYTES), ok ? (const void*)(src + i * rstep) : (const void*)base,
                       ok ? bytes : 0);
        }
    }
};

__global__ void __launch_bounds__(S_THREADS, 2) dgemm_nt_small_kernel(GemmArgs p) {
    static_assert(S_RSTEP % 2 == 0 && S_LOADS * S_RSTEP == SB, "loader geometry");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (blockIdx.z) {
        p.A += (int64_t)blockIdx.z * p.sA;
        p.B += (int64_t)blockIdx.z * p.sB;
        if (p.C) p.C += (int64_t)blockIdx.z * p.sC;
        if (p.Ct) p.Ct += (int64_t)blockIdx.z * p.sCt;
    }
    const int m0 = blockIdx.y * SB, n0 = blockIdx.x * SB;
    if (p.lower_only && n0 >= m0 + SB) return;
    const uint32_t smem_base = smem_u32(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp >> 1) * 32, wn0 = (warp & 1) * 32;
    int kbeg = 0, kend = p.K;
    if (p.k_from_max) kbeg = ((m0 > n0 ? m0 : n0) / BK) * BK;
    if (p.k_to_min) { const int lim = (m0 < n0 ? m0 : n0) + SB; kend = lim < p.K ? lim : p.K; }
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    if (kend > kbeg) {
        const int KT = (kend - kbeg + BK - 1) / BK;
        SmallLoader la, lb;
        la.init(p.A, p.lda, m0, p.M, tid);
        lb.init(p.B, p.ldb, n0, p.N, tid);
#pragma unroll
        for (int s = 0; s < S_STAGES - 1; ++s) {
            if (s < KT) {
                la.load(smem_base + s * S_STAGE_BYTES, kbeg + s * BK, kend);
                lb.load(smem_base + s * S_STAGE_BYTES + S_TILE_BYTES, kbeg + s * BK, kend);
            }
            cp_async_commit();
        }
        for (int kt = 0; kt < KT; ++kt) {
            cp_async_wait<S_STAGES - 2>();
            __syncthreads();
            const int nk = kt + S_STAGES - 1;
            if (nk < KT) {
                la.load(smem_base + (nk % S_STAGES) * S_STAGE_BYTES, kbeg + nk * BK, kend);
                lb.load(smem_base + (nk % S_STAGES) * S_STAGE_BYTES + S_TILE_BYTES, kbeg + nk * BK, kend);
            }
            cp_async_commit();
            const uint32_t sA = smem_base + (kt % S_STAGES) * S_STAGE_BYTES, sB = sA + S_TILE_BYTES;
#pragma unroll
            for (int kg = 0; kg < BK / 8; ++kg) {
                const uint32_t chunk = (uint32_t)(((4 * kg + t) ^ ((g & 1) << 2)) << 4);
                double2 a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = lds128(sA + (wm0 + 8 * i + g) * ROW_BYTES + chunk);
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = lds128(sB + (wn0 + 8 * j + g) * ROW_BYTES + chunk);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
            }
        }
        cp_async_wait<0>();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = m0 + wm0 + 8 * i + g;
        if (row >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
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

// IPP_GEMM_SMALL=0: always the 128 x 128 kernel (A/B measurements)
static bool gemm_small_enabled() {
    static const int use = [] { const char* e = getenv("IPP_GEMM_SMALL"); return (e && e[0] == '0') ? 0 : 1; }();
    return use != 0;
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
    // big tiles for at most half the SMs (counting only the tiles a lower_only product computes): four times as many small ones
    const int64_t tm = (g.M + BM - 1) / BM, tn = (g.N + BN - 1) / BN;
    const int64_t big_tiles = (g.lower_only ? (tn >= tm ? tm * (tm + 1) / 2 : tm * tn - tn * (tn - 1) / 2) : tm * tn) *
                              (g.batch > 1 ? g.batch : 1);
    if (gemm_small_enabled() && 2 * big_tiles <= sm_count()) {
        IPP_OPT_IN_SMEM(dgemm_nt_small_kernel, S_SMEM);
        dim3 sgrid((g.N + SB - 1) / SB, (g.M + SB - 1) / SB, g.batch > 1 ? g.batch : 1);
        dgemm_nt_small_kernel<<<sgrid, S_THREADS, S_SMEM, stream>>>(g);
        IPP_LAUNCH_CHECK();
        return IPP_OK;
    }
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
