// a2: dense FP64 refresh of the belief (GPy pdinv restated on the device):
//   A = K(X,X) + diag_add I  ->  L = chol(A)  ->  L^-1  ->  W^-1 = L^-T L^-1 (exactly symmetric)  ->  alpha = W^-1 z
//
// Right-looking blocked Cholesky, block 64: a single-CTA kernel factors the 64x64 diagonal block in
// shared memory and also emits its triangular inverse; the panel solve (L21 = A21 * Ldd^-T) and the
// trailing update (A22 -= L21 L21^T, lower tiles) are DMMA GEMMs.  The triangular inverse is built by
// recursive doubling with GEMMs only (L^-1 and its transpose U are kept side by side so that every
// product is of the K-major "NT" form the DMMA kernel wants):
//   Linv21 = -Linv22 * (L21 * Linv11)
// and W^-1 = U U^T is a lower-tile GEMM whose contraction starts at max(m0, n0), mirrored on store.
#include "internal.cuh"
#include "dgemm_dmma.cuh"
#include <type_traits>
#include <utility>

namespace ipp {

constexpr int NB = 64;
constexpr int DIAG_LD = NB + 1;
__device__ long long g_potrf_clock[4];       // SM clocks of the last diagonal block: load, pivots + inverse, -, store (diagnostics)

// Cholesky of the jb x jb lower block at A (leading dimension lda) + its inverse into dinv[64][64].
//
// r1: 3 barriers, a division per column entry and two integer divisions per updated element for each of the 64 pivots,
// then 64 threads each running a serial O(n^2) substitution -- 100 us per block.  r2a: one barrier per pivot, the block
// updated in shared memory by 16 x 16 threads, inverse by recursive doubling -- 44 us (757 clocks per pivot against ~200 of
// dependency chain: every pivot re-read and re-wrote its thread's 16 elements through shared memory).  Now the block
// LIVES IN REGISTERS: warp w owns the columns w, w + 8, ... and lane l the rows l, l + 32 of each (16 elements per thread),
// the pivot loop is unrolled by column slot so every register index is a constant, and per pivot
//   * the owning warp takes the pivot from the lane that holds row j (one shuffle), forms r = 1 / sqrt(d) (rsqrt_fast),
//     scales its column and puts the 64 scaled entries (+ r) into shared memory -- the only thing that ever goes there;
//   * one barrier (the column buffer is double-buffered);
//   * every thread reads the two entries of its rows and the (at most eight) entries of its columns and applies the
//     rank-1 update to its registers;
//   * the INVERSE rides along: with B = I at the start, row j of L^-1 is B[j][:] r and B[i][:] -= l_ij (row j) for i > j
//     -- the same rank-1 pattern; row j of a warp's columns sits in one lane, so it is broadcast by shuffles, no barrier.
//     (The r2a version ran the inverse as a second phase, recursive doubling through shared memory.)
__global__ void __launch_bounds__(256, 1) potrf_diag_kernel(double* __restrict__ A, int64_t lda, int jb,
                                                         double* __restrict__ dinv, double* logdet,
                                                         int32_t* info, int j0) {
    __shared__ double tile[NB][DIAG_LD];          // staging: A on the way in, L and L^-1 on the way out (coalesced I/O)
    __shared__ double colbuf[2][NB];              // the scaled column of the current pivot
    __shared__ double rbuf[NB], dbuf[NB];         // 1 / l_jj, l_jj
    __shared__ int bad;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long t0 = clock64();
    {
        double v[16];                             // all sixteen loads of a thread in flight before the first store
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int r = (tid >> 6) + 4 * u, c = tid & 63;
            v[u] = (r == c) ? 1.0 : 0.0;          // rows / columns beyond jb: identity
            if (r < jb && c <= r) v[u] = A[(int64_t)r * lda + c];
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) tile[(tid >> 6) + 4 * u][tid & 63] = v[u];
    }
    if (tid == 0) bad = 0;
    __syncthreads();
    double a[2][8], B[2][8];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            a[i][m] = tile[lane + 32 * i][warp + 8 * m];
            B[i][m] = (lane + 32 * i == warp + 8 * m) ? 1.0 : 0.0;
        }
    const long long t1 = clock64();
    // The eight pivots j = 8 MJ + w8 of column slot MJ share one copy of the code (register indices depend on MJ only;
    // a fully unrolled loop -- 64 copies, ~160 KB of instructions -- ran at 470-520 clocks per pivot, bound by
    // instruction fetch: no line was ever executed twice).
    auto slot = [&](auto mjc) {
        constexpr int MJ = decltype(mjc)::value, IJ = MJ >> 2;
#pragma unroll 1
        for (int w8 = 0; w8 < 8; ++w8) {
            const int j = 8 * MJ + w8, lj = j & 31;
            if (warp == w8) {       // the owning warp takes d from the lane that holds row j, scales its column, publishes it
                const double d0 = __shfl_sync(0xffffffffu, a[IJ][MJ], lj);
                const bool ok = d0 > 0.0;                           // also catches NaN, like LAPACK's dpotrf
                const double d = ok ? d0 : 1.0;
                const double r = rsqrt_fast(d);
                if (!ok && lane == 0 && bad == 0) bad = j + 1;
#pragma unroll
                for (int i = IJ; i < 2; ++i) {                      // (row slot 0 lies above the pivots of slots 4 .. 7)
                    const int row = lane + 32 * i;
                    const double v = row > j ? a[i][MJ] * r : (row == j ? d * r : 0.0);
                    a[i][MJ] = v;
                    colbuf[j & 1][row] = v;
                }
                if (lane == 0) { rbuf[j] = r; dbuf[j] = d * r; }
            }
            __syncthreads();
            const double r = rbuf[j];
            double lr[2] = {0.0, 0.0};
#pragma unroll
            for (int i = IJ; i < 2; ++i) lr[i] = (lane + 32 * i > j) ? colbuf[j & 1][lane + 32 * i] : 0.0;
#pragma unroll
            for (int m = MJ; m < 8; ++m) {                          // trailing update: columns c > j of this warp
                const int c = warp + 8 * m;
                if (c > j) {
                    const double lc = colbuf[j & 1][c];
#pragma unroll
                    for (int i = IJ; i < 2; ++i) a[i][m] = fma(-lr[i], lc, a[i][m]);
                }
            }
#pragma unroll
            for (int m = 0; m <= MJ; ++m) {                         // inverse: row j becomes final, the rows below take their update
                const int c = warp + 8 * m;
                if (c <= j) {
                    const double x = __shfl_sync(0xffffffffu, B[IJ][m], lj) * r;
#pragma unroll
                    for (int i = IJ; i < 2; ++i) B[i][m] = (lane + 32 * i == j) ? x : fma(-lr[i], x, B[i][m]);
                }
            }
        }
    };
    [&]<int... Ms>(std::integer_sequence<int, Ms...>) { (slot(std::integral_constant<int, Ms>{}), ...); }(std::make_integer_sequence<int, 8>{});
    const long long t2 = clock64();
    __syncthreads();
    // ---- out: the factor into A (lower triangle of the jb x jb block), its inverse into dinv (zeros elsewhere)
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int m = 0; m < 8; ++m) tile[lane + 32 * i][warp + 8 * m] = a[i][m];
    __syncthreads();
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int r = idx >> 6, c = idx & 63;
        if (r < jb && c <= r) A[(int64_t)r * lda + c] = tile[r][c];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int m = 0; m < 8; ++m) tile[lane + 32 * i][warp + 8 * m] = B[i][m];
    __syncthreads();
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int r = idx >> 6, c = idx & 63;
        dinv[idx] = (r < jb && c <= r) ? tile[r][c] : 0.0;
    }
    if (tid < 32) {
        double sacc = 0.0;
        for (int j = tid; j < jb; j += 32) sacc += log(dbuf[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
        if (tid == 0) {
            if (logdet) *logdet += 2.0 * sacc;
            if (bad && info && *info == 0) *info = j0 + bad;
            g_potrf_clock[0] = t1 - t0; g_potrf_clock[1] = t2 - t1; g_potrf_clock[2] = 0; g_potrf_clock[3] = clock64() - t2;
        }
    }
}

__global__ void init_factor_scalars_kernel(double* logdet, int32_t* info) {
    if (logdet) *logdet = 0.0;
    if (info) *info = 0;
}

// Two-level blocking.  A rank-64 update of the whole trailing matrix after every diagonal block ran at 10 TFLOP/s (CTAs
// of two k-steps each: all prologue and epilogue; 4.5 of the 8.5 ms of a refresh at N = 4096): inside a panel of
// POTRF_PANEL columns the rank-64 updates now touch the rest of the PANEL only, and the trailing matrix takes one
// rank-POTRF_PANEL update per panel.
constexpr int64_t POTRF_PANEL = 256;
int potrf_lower(double* A, int64_t N, int64_t lda, double* dinv, double* logdet, int32_t* info, cudaStream_t s) {
    init_factor_scalars_kernel<<<1, 1, 0, s>>>(logdet, info);
    IPP_LAUNCH_CHECK();
    for (int64_t J0 = 0; J0 < N; J0 += POTRF_PANEL) {
        const int64_t J1 = (J0 + POTRF_PANEL < N) ? J0 + POTRF_PANEL : N;
        for (int64_t j0 = J0; j0 < J1; j0 += NB) {
            const int jb = (int)((J1 - j0) < NB ? (J1 - j0) : NB);
            double* Ajj = A + j0 * lda + j0;
            double* D = dinv + (j0 / NB) * NB * NB;
            potrf_diag_kernel<<<1, 256, 0, s>>>(Ajj, lda, jb, D, logdet, info, (int)j0);
            IPP_LAUNCH_CHECK();
            const int64_t rows = N - j0 - jb;
            if (rows <= 0) break;
            double* A21 = A + (j0 + jb) * lda + j0;
            GemmArgs g{};                                   // L21 = A21 * Ldd^-T (in place), all rows below the block
            g.A = A21; g.lda = lda; g.B = D; g.ldb = NB; g.C = A21; g.ldc = lda;
            g.M = (int)rows; g.N = jb; g.K = jb; g.alpha = 1.0; g.beta = 0.0;
            IPP_TRY(launch_gemm(g, s));
            const int64_t cols = J1 - (j0 + jb);            // the panel's columns still to be factored
            if (cols > 0) {
                GemmArgs h{};                               // A[j0 + jb :, j0 + jb : J1] -= L21 L21[: cols]^T on the lower tiles
                h.A = A21; h.lda = lda; h.B = A21; h.ldb = lda; h.C = A + (j0 + jb) * lda + (j0 + jb); h.ldc = lda;
                h.M = (int)rows; h.N = (int)cols; h.K = jb; h.alpha = -1.0; h.beta = 1.0; h.lower_only = 1;
                IPP_TRY(launch_gemm(h, s));
            }
        }
        const int64_t rest = N - J1;
        if (rest > 0) {
            GemmArgs t{};                                   // trailing matrix -= L[J1 :, J0 : J1] L[J1 :, J0 : J1]^T, lower tiles
            t.A = A + J1 * lda + J0; t.lda = lda; t.B = t.A; t.ldb = lda; t.C = A + J1 * lda + J1; t.ldc = lda;
            t.M = (int)rest; t.N = (int)rest; t.K = (int)(J1 - J0); t.alpha = -1.0; t.beta = 1.0; t.lower_only = 1;
            IPP_TRY(launch_gemm(t, s));
        }
    }
    return IPP_OK;
}

// A = K(X, X) + diag_add * I  (full matrix; the Cholesky reads the lower triangle)
__global__ void add_diag_kernel(double* A, int64_t N, int64_t lda, double v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) A[i * lda + i] += v;
}

// scatter the 64x64 inverse diagonal blocks into Linv (lower) and U = Linv^T (upper)
__global__ void place_diag_inverse_kernel(const double* __restrict__ dinv, int64_t N, double* __restrict__ Linv,
                                          int64_t ldl, double* __restrict__ U, int64_t ld) {
    const int64_t j0 = (int64_t)blockIdx.x * NB;
    const double* D = dinv + (int64_t)blockIdx.x * NB * NB;
    for (int idx = threadIdx.x; idx < NB * NB; idx += blockDim.x) {
        const int r = idx >> 6, c = idx & 63;
        if (j0 + r < N && j0 + c < N) {
            Linv[(j0 + r) * ldl + j0 + c] = D[idx];
            U[(j0 + c) * ld + j0 + r] = D[idx];
        }
    }
}

}  // namespace ipp

using namespace ipp;

extern "C" int ipp_potrf_debug_clocks(int64_t out[4]) {
    long long v[4];
    IPP_CUDA(cudaDeviceSynchronize());
    IPP_CUDA(cudaMemcpyFromSymbol(v, ipp::g_potrf_clock, sizeof(v)));
    for (int i = 0; i < 4; ++i) out[i] = v[i];
    return IPP_OK;
}

extern "C" int64_t ipp_potrf_workspace(int64_t N) { return ((N + NB - 1) / NB + 1) * NB * NB + 16; }

extern "C" int ipp_potrf_lower(double* A, int64_t N, int64_t lda, double* logdet, int32_t* info, double* work,
                               void* stream) {
    IPP_CHECK_ARG(A && work && N >= 1 && lda >= N && lda % 2 == 0, "bad arguments");
    IPP_CHECK_ARG(reinterpret_cast<uintptr_t>(A) % 16 == 0, "A must be 16-byte aligned");
    return potrf_lower(A, N, lda, work, logdet, info, static_cast<cudaStream_t>(stream));
}

// pdinv of a dense SPD matrix already in device memory (lower triangle read, A overwritten by its factor)
int ipp::pdinv_device(double* Lf, int64_t N, int64_t lda, double* winv, double* linv, int64_t ldw,
                      double* logdet, int32_t* info, double* work, cudaStream_t s) {
    const int64_t ld = round_up(N, 16);
    double* U = work;                     // Linv^T
    double* Tt = U + N * ld;              // scratch for (L21 Linv11)^T
    double* Li = Tt + N * ld;             // Linv when the caller does not want it
    double* dinv = Li + N * ld;
    double* Linv = linv ? linv : Li;
    const int64_t ldl = linv ? ldw : ld;

    IPP_TRY(potrf_lower(Lf, N, lda, dinv, logdet, info, s));
    IPP_CUDA(cudaMemsetAsync(Linv, 0, sizeof(double) * N * ldl, s));
    IPP_CUDA(cudaMemsetAsync(U, 0, sizeof(double) * N * ld, s));
    place_diag_inverse_kernel<<<(unsigned)((N + NB - 1) / NB), 256, 0, s>>>(dinv, N, Linv, ldl, U, ld);
    IPP_LAUNCH_CHECK();
    // Doubling: at block size b the pairs (rows r0 + b .., columns r0 ..), r0 = 0, 2b, 4b, ... are independent.  All FULL
    // pairs of a level go out as one batched launch per product (the r1 version launched every pair on its own: 126
    // launches at N = 4096, most of them far too small to fill the GPU, 4.9 of the 12.9 ms of a refresh); a ragged last
    // pair (N not a multiple of 2b) follows on its own.  Tt holds (L21 Linv11)^T of pair r0 in its rows r0 .. r0 + b.
    for (int64_t b = NB; b < N; b *= 2) {
        const int64_t pairs = (N - b + 2 * b - 1) / (2 * b);            // r0 + b < N
        const int64_t last_r0 = (pairs - 1) * 2 * b;
        const int64_t last_b2 = (N - last_r0 - b) < b ? (N - last_r0 - b) : b;
        const int64_t full = last_b2 == b ? pairs : pairs - 1;
        for (int part = 0; part < 2; ++part) {
            const int64_t r0 = part == 0 ? 0 : last_r0;
            const int64_t nb = part == 0 ? full : (full < pairs ? 1 : 0);
            const int64_t b2 = part == 0 ? b : last_b2;
            if (nb <= 0) continue;
            GemmArgs g{};                               // Tt = (L21 * Linv11)^T  [b][b2]
            g.A = Lf + (r0 + b) * lda + r0; g.lda = lda;
            g.B = U + r0 * ld + r0; g.ldb = ld;
            g.C = nullptr; g.Ct = Tt + r0 * ld; g.ldct = ld;
            g.M = (int)b2; g.N = (int)b; g.K = (int)b; g.alpha = 1.0; g.beta = 0.0;
            g.batch = (int)nb; g.sA = 2 * b * lda + 2 * b; g.sB = 2 * b * ld + 2 * b; g.sC = 0; g.sCt = 2 * b * ld;
            IPP_TRY(launch_gemm(g, s));
            GemmArgs h{};                               // Linv21 = -Linv22 * T, and its transpose into U
            h.A = Linv + (r0 + b) * ldl + (r0 + b); h.lda = ldl;
            h.B = Tt + r0 * ld; h.ldb = ld;
            h.C = Linv + (r0 + b) * ldl + r0; h.ldc = ldl;
            h.Ct = U + r0 * ld + (r0 + b); h.ldct = ld;
            h.M = (int)b2; h.N = (int)b; h.K = (int)b2; h.alpha = -1.0; h.beta = 0.0;
            h.batch = (int)nb; h.sA = 2 * b * ldl + 2 * b; h.sB = 2 * b * ld; h.sC = 2 * b * ldl + 2 * b; h.sCt = 2 * b * ld + 2 * b;
            IPP_TRY(launch_gemm(h, s));
        }
    }
    GemmArgs w{};                                       // W^-1 = U U^T, lower tiles, mirrored
    w.A = U; w.lda = ld; w.B = U; w.ldb = ld; w.C = winv; w.ldc = ldw; w.Ct = winv; w.ldct = ldw;
    w.M = (int)N; w.N = (int)N; w.K = (int)N; w.alpha = 1.0; w.beta = 0.0; w.lower_only = 1; w.k_from_max = 1;
    IPP_TRY(launch_gemm(w, s));
    return IPP_OK;
}

extern "C" int64_t ipp_pdinv_workspace(int64_t N) {
    const int64_t ld = round_up(N > 0 ? N : 1, 16);
    return 3 * N * ld + ipp_potrf_workspace(N) + 64;
}

extern "C" int ipp_pdinv(double* A, int64_t N, int64_t lda, double* winv, double* linv, int64_t ldw,
                         double* logdet, int32_t* info, double* work, void* stream_) {
    IPP_CHECK_ARG(A && winv && info && work, "null pointer");
    IPP_CHECK_ARG(N >= 1 && N < (1LL << 30) && lda >= N && lda % 2 == 0 && ldw >= N && ldw % 2 == 0, "bad shape");
    IPP_CHECK_ARG(reinterpret_cast<uintptr_t>(A) % 16 == 0 && reinterpret_cast<uintptr_t>(winv) % 16 == 0 &&
                  (linv == nullptr || reinterpret_cast<uintptr_t>(linv) % 16 == 0), "matrices must be 16-byte aligned");
    return pdinv_device(A, N, lda, winv, linv, ldw, logdet, info, work, static_cast<cudaStream_t>(stream_));
}

extern "C" int64_t ipp_gp_factor_workspace(int64_t N) {
    const int64_t ld = round_up(N > 0 ? N : 1, 16);
    return N * ld + ipp_pdinv_workspace(N);
}

extern "C" int ipp_gp_factor(const ipp_rbf* kern, const double* X, const double* z, int64_t N, double diag_add,
                             double* winv, double* linv, int64_t ldw, double* alpha, double* logdet, int32_t* info,
                             double* work, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    IPP_CHECK_ARG(kern && X && winv && work && info, "null pointer");
    IPP_CHECK_ARG(N >= 1 && N < (1LL << 30), "bad N");
    IPP_CHECK_ARG(ldw >= N && ldw % 2 == 0 && reinterpret_cast<uintptr_t>(winv) % 16 == 0, "winv: even ldw, 16-byte aligned");
    IPP_CHECK_ARG(linv == nullptr || reinterpret_cast<uintptr_t>(linv) % 16 == 0, "linv must be 16-byte aligned");
    const int64_t ld = round_up(N, 16);
    double* Lf = work;                    // K + diag_add I, then its Cholesky factor (lower)
    IPP_TRY(launch_rbf_cross(*kern, X, N, X, N, Lf, ld, s));
    add_diag_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(Lf, N, ld, diag_add);
    IPP_LAUNCH_CHECK();
    IPP_TRY(pdinv_device(Lf, N, ld, winv, linv, ldw, logdet, info, work + N * ld, s));
    if (alpha && z) IPP_TRY(launch_gemv(winv, ldw, N, N, z, alpha, s));
    return IPP_OK;
}
