// a4: OnlineGPModel.update_model (gpmodel_library.py:230-279) -- block-inverse append of k <= 32 points.
//
//   Q = K(X_old, X_new) (N x k),  S = K(X_new, X_new),  P = the current inverse (symmetric), R = Q^T
//   V  = P Q         (N x k)                                          (= (R P)^T because P is symmetric)
//   M  = pdinv(S - V^T Q + diag_add I) = Li^T Li                      (:252-255; Li = inverse Cholesky factor)
//   Pnew = P + V M V^T,  Qnew = -(V M),  Rnew = Qnew^T,  Snew = M     (:257-265)
//   alpha = Wnew z                                                    (:273, full recompute like the reference)
//
// The reference evaluates Pnew as ((((P Q) M) R) P), an O(N^3) product; here the same factors are associated
// as U U^T with U = V Li^T (N x k), O(N^2 k), which makes Pnew EXACTLY symmetric (fma(U_ia, U_ja, s) commutes),
// like pdinv's symmetrified output: the symmetric-half variance form of predict (IPP_VAR_WINV_SYM) relies on
// it -- an asymmetry of 1e-11 in W^-1 is amplified by |k*|^2 ~ 1e4 into the posterior variance.  Every pass over
// the N x N inverse is a coalesced HBM stream (read P for V, read P + write Pnew, read Wnew once for alpha).
#include "internal.cuh"

namespace ipp {

constexpr int KMAX = IPP_APPEND_MAX_K;

// V[i][a] = sum_j P[i][j] Q[j][a].  A block owns 16 rows (two per warp) and walks j in chunks of 256: the chunk of Q is
// staged in shared memory once per block (transposed, pitch 257: conflict-free both ways), every lane then has sixteen
// independent 8-byte loads of P in flight (two rows x eight strides of 32) before the first FMA.  (The r1 kernel -- one
// warp per row, Q re-read from L2 with a 256-byte lane stride, one load in flight per lane -- ran at 0.39 TB/s at
// N = 4096: 341 us for the 134 MB of W^-1, four times the whole rank-k update that follows.)
constexpr int PQ_JC = 256, PQ_PITCH = PQ_JC + 1;
template <int KT>
__global__ void __launch_bounds__(256) append_pq_kernel(const double* __restrict__ P, int64_t ldp, int64_t N, int k,
                                                        const double* __restrict__ Q, double* __restrict__ V) {
    extern __shared__ double sq[];                       // [KT][PQ_PITCH]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i0 = (int64_t)blockIdx.x * 16 + warp * 2;
    double acc[2][KT];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int a = 0; a < KT; ++a) acc[r][a] = 0.0;
    for (int64_t j0 = 0; j0 < N; j0 += PQ_JC) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < KT * PQ_JC; idx += 256) {
            const int jj = idx / KT, a = idx % KT;
            sq[a * PQ_PITCH + jj] = (a < k && j0 + jj < N) ? Q[(j0 + jj) * KMAX + a] : 0.0;
        }
        __syncthreads();
        double p[2][8];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int64_t j = j0 + lane + 32 * t;
                p[r][t] = (i0 + r < N && j < N) ? P[(i0 + r) * ldp + j] : 0.0;
            }
#pragma unroll
        for (int t = 0; t < 8; ++t)
#pragma unroll
            for (int a = 0; a < KT; ++a) {
                const double q = sq[a * PQ_PITCH + lane + 32 * t];
                acc[0][a] = fma(p[0][t], q, acc[0][a]);
                acc[1][a] = fma(p[1][t], q, acc[1][a]);
            }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int a = 0; a < KT; ++a) {
            double s = acc[r][a];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0 && a < k && i0 + r < N) V[(i0 + r) * KMAX + a] = s;
        }
}

template <int KT>
static int launch_append_pq(const double* P, int64_t ldp, int64_t N, int k, const double* Q, double* V, cudaStream_t s) {
    constexpr int bytes = KT * PQ_PITCH * (int)sizeof(double);
    if (bytes > 48 * 1024) IPP_OPT_IN_SMEM(append_pq_kernel<KT>, bytes);
    append_pq_kernel<KT><<<(unsigned)((N + 15) / 16), 256, bytes, s>>>(P, ldp, N, k, Q, V);
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

// G[a][b] = sum_j V[j][a] Q[j][b]: one block per (a, b), fixed-order tree over the block's 256 partial sums
__global__ void __launch_bounds__(256) append_gram_kernel(const double* __restrict__ Vt, const double* __restrict__ Q,
                                                          int64_t N, int k, double* __restrict__ G) {
    __shared__ double red[8];
    const int a = blockIdx.x / k, b = blockIdx.x % k;
    double s = 0.0;
    for (int64_t j = threadIdx.x; j < N; j += 256) s = fma(Vt[j * KMAX + a], Q[j * KMAX + b], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0)
        G[a * KMAX + b] = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
}

// M = pdinv(S - G + diag_add I) for a k x k block, with GPy's jitchol retries
// (linalg.py: up to 5 retries with jitter mean(diag) * 1e-6 * 10^i).  One warp; lane = row.
__global__ void __launch_bounds__(32) append_schur_inverse_kernel(const double* __restrict__ S, const double* __restrict__ G,
                                                                  int k, double diag_add, double* __restrict__ Mout,
                                                                  double* __restrict__ Liout, int32_t* info) {
    __shared__ double a0[KMAX][KMAX + 1], a[KMAX][KMAX + 1], li[KMAX][KMAX + 1];
    __shared__ int bad;
    const int r = threadIdx.x;
    for (int c = 0; c < k; ++c)
        if (r < k) a0[r][c] = S[r * KMAX + c] - G[r * KMAX + c] + (r == c ? diag_add : 0.0);
    __syncwarp();
    double mean_diag = 0.0;
    for (int c = 0; c < k; ++c) mean_diag += a0[c][c];
    mean_diag /= k;
    double jitter = 0.0;
    int tries = 0;
    for (;;) {
        for (int c = 0; c < k; ++c)
            if (r < k) a[r][c] = a0[r][c] + (r == c ? jitter : 0.0);
        if (r == 0) bad = 0;
        __syncwarp();
        for (int j = 0; j < k; ++j) {
            if (r == 0) {
                double d = a[j][j];
                if (!(d > 0.0)) { if (!bad) bad = j + 1; d = 1.0; }
                a[j][j] = sqrt(d);
            }
            __syncwarp();
            if (r > j && r < k) a[r][j] /= a[j][j];
            __syncwarp();
            if (r > j && r < k)
                for (int c = j + 1; c <= r; ++c) a[r][c] = fma(-a[r][j], a[c][j], a[r][c]);
            __syncwarp();
        }
        if (!bad) break;
        if (tries >= 5 || !(mean_diag > 0.0)) break;
        jitter = (tries == 0) ? mean_diag * 1e-6 : jitter * 10.0;
        ++tries;
        __syncwarp();
    }
    if (bad) {
        if (r == 0 && info && *info == 0) *info = bad;
        // still emit something finite-shaped so downstream kernels do not fault
    }
    // Linv by forward substitution (lane = column), then Ai = Linv^T Linv, lower triangle mirrored
    if (r < k) {
        const int c = r;
        for (int i = 0; i < c; ++i) li[i][c] = 0.0;
        li[c][c] = 1.0 / a[c][c];
        for (int i = c + 1; i < k; ++i) {
            double s = 0.0;
            for (int q = c; q < i; ++q) s = fma(a[i][q], li[q][c], s);
            li[i][c] = -s / a[i][i];
        }
    }
    __syncwarp();
    if (r < k) {
        for (int c = 0; c <= r; ++c) {
            double s = 0.0;
            for (int q = r; q < k; ++q) s = fma(li[q][r], li[q][c], s);
            Mout[r * KMAX + c] = s;
            Mout[c * KMAX + r] = s;
        }
        for (int c = 0; c < k; ++c) Liout[r * KMAX + c] = li[r][c];
    }
}

// U[i][a] = sum_b V[i][b] Li[a][b]  (U = V Li^T, so V M V^T = U U^T);  VM[i][b] = sum_a V[i][a] M[a][b]
__global__ void __launch_bounds__(256) append_apply_m_kernel(const double* __restrict__ V, const double* __restrict__ Li,
                                                             const double* __restrict__ Mm, int64_t N, int k,
                                                             double* __restrict__ U, double* __restrict__ VM) {
    __shared__ double m[KMAX][KMAX + 1], li[KMAX][KMAX + 1];
    for (int idx = threadIdx.x; idx < k * k; idx += blockDim.x) {
        m[idx / k][idx % k] = Mm[(idx / k) * KMAX + idx % k];
        li[idx / k][idx % k] = Li[(idx / k) * KMAX + idx % k];
    }
    __syncthreads();
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * k) return;
    const int64_t i = idx / k;
    const int b = (int)(idx % k);
    double s = 0.0, u = 0.0;
    for (int a = 0; a < k; ++a) s = fma(V[i * KMAX + a], m[a][b], s);
    for (int a = 0; a <= b; ++a) u = fma(V[i * KMAX + a], li[b][a], u);       // Li is lower triangular
    VM[i * KMAX + b] = s;
    U[i * KMAX + b] = u;
}

// Wnew: 64 x 64 output tile per block; rank-k term U U^T from shared-memory panels (bitwise symmetric)
__global__ void __launch_bounds__(256) append_update_kernel(const double* __restrict__ P, int64_t ldp, int64_t N, int k,
                                                            const double* __restrict__ U, const double* __restrict__ VM,
                                                            const double* __restrict__ Mm,
                                                            double* __restrict__ Wn, int64_t ldn) {
    __shared__ double sui[64][KMAX + 1], suj[64][KMAX + 1];
    const int64_t i0 = (int64_t)blockIdx.y * 64, j0 = (int64_t)blockIdx.x * 64;
    const int64_t NT = N + k;
    for (int idx = threadIdx.x; idx < 64 * k; idx += blockDim.x) {
        const int r = idx / k, a = idx % k;
        sui[r][a] = (i0 + r < N) ? U[(i0 + r) * KMAX + a] : 0.0;
        suj[r][a] = (j0 + r < N) ? U[(j0 + r) * KMAX + a] : 0.0;
    }
    __syncthreads();
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;   // 64 columns x 4 row phases
    const int64_t j = j0 + tx;
    if (j >= NT) return;
    for (int r = ty; r < 64; r += 4) {
        const int64_t i = i0 + r;
        if (i >= NT) break;
        double v;
        if (i < N && j < N) {
            double s = 0.0;
            for (int a = 0; a < k; ++a) s = fma(sui[r][a], suj[tx][a], s);
            v = P[i * ldp + j] + s;                            // P bitwise symmetric in => Wnew bitwise symmetric out
        } else if (i < N) {
            v = -VM[i * KMAX + (j - N)];                       // Qnew = -(P Q) M
        } else if (j < N) {
            v = -VM[j * KMAX + (i - N)];                       // Rnew = Qnew^T
        } else {
            v = Mm[(i - N) * KMAX + (j - N)];                  // Snew = M
        }
        Wn[i * ldn + j] = v;
    }
}

}  // namespace ipp

using namespace ipp;

extern "C" int64_t ipp_gp_append_workspace(int64_t N, int64_t k) {
    (void)k;
    return 5 * (N + KMAX) * KMAX + 4 * KMAX * KMAX + 64;
}

extern "C" int ipp_gp_append(const ipp_rbf* kern, const double* X_all, const double* z_all, int64_t N, int64_t k,
                             double diag_add, const double* winv_old, int64_t ldw_old,
                             double* winv_new, int64_t ldw_new, double* alpha_new, int32_t* info,
                             double* work, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    IPP_CHECK_ARG(kern && X_all && winv_old && winv_new && work && info, "null pointer");
    IPP_CHECK_ARG(N >= 1 && k >= 1 && k <= KMAX, "need N >= 1 and 1 <= k <= IPP_APPEND_MAX_K");
    IPP_CHECK_ARG(ldw_old >= N && ldw_new >= N + k, "leading dimensions too small");
    IPP_CHECK_ARG(kern->dim == 2 || kern->dim == 3, "dimension must be 2 or 3");
    const int D = kern->dim;
    const double* Xnew = X_all + N * D;
    double* Q = work;                          // [N][KMAX]
    double* V = Q + (N + KMAX) * KMAX;
    double* U = V + (N + KMAX) * KMAX;
    double* VM = U + (N + KMAX) * KMAX;
    double* MVt = VM + (N + KMAX) * KMAX;      // (spare panel)
    double* S = MVt + (N + KMAX) * KMAX;       // [KMAX][KMAX]
    double* G = S + KMAX * KMAX;
    double* Mm = G + KMAX * KMAX;
    double* Li = Mm + KMAX * KMAX;
    IPP_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), s));
    IPP_TRY(launch_rbf_cross(*kern, X_all, N, Xnew, k, Q, KMAX, s));
    IPP_TRY(launch_rbf_cross(*kern, Xnew, k, Xnew, k, S, KMAX, s));
    if (k <= 4) IPP_TRY(launch_append_pq<4>(winv_old, ldw_old, N, (int)k, Q, V, s));
    else if (k <= 8) IPP_TRY(launch_append_pq<8>(winv_old, ldw_old, N, (int)k, Q, V, s));
    else if (k <= 16) IPP_TRY(launch_append_pq<16>(winv_old, ldw_old, N, (int)k, Q, V, s));
    else IPP_TRY(launch_append_pq<32>(winv_old, ldw_old, N, (int)k, Q, V, s));
    append_gram_kernel<<<(unsigned)(k * k), 256, 0, s>>>(V, Q, N, (int)k, G);
    IPP_LAUNCH_CHECK();
    append_schur_inverse_kernel<<<1, 32, 0, s>>>(S, G, (int)k, diag_add, Mm, Li, info);
    IPP_LAUNCH_CHECK();
    append_apply_m_kernel<<<(unsigned)((N * k + 255) / 256), 256, 0, s>>>(V, Li, Mm, N, (int)k, U, VM);
    IPP_LAUNCH_CHECK();
    const int64_t NT = N + k;
    dim3 grid((unsigned)((NT + 63) / 64), (unsigned)((NT + 63) / 64));
    append_update_kernel<<<grid, 256, 0, s>>>(winv_old, ldw_old, N, (int)k, U, VM, Mm, winv_new, ldw_new);
    IPP_LAUNCH_CHECK();
    if (alpha_new && z_all) IPP_TRY(launch_gemv(winv_new, ldw_new, NT, NT, z_all, alpha_new, s));
    return IPP_OK;
}
