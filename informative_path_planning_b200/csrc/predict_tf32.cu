// a3, "3xTF32" mode: posterior variance in GPy's Cholesky form on the 5th-generation tensor cores.
//
//   var[m] = variance - sum_n (sum_{j<=n} Linv[n][j] k(q_m, x_j))^2 + noise_add
//
// Every FP64 operand is split once into two TF32 numbers, a = hi + lo (22 significant bits), and the
// contraction is evaluated as hi*hi + hi*lo + lo*hi with FP32 accumulation ("3xTF32"):
//   * ipp_tf32_split           Linv -> (hi, lo) float arrays, once per refresh of the belief;
//   * rbf_queries_tf32_kernel  cross-covariance chunk Kxt[m][j] generated in FP64 (one exp per pair), stored
//                              query-major as (hi, lo); the posterior mean is accumulated in FP64 there;
//   * predict_var_tf32_kernel  D[m][n] = sum_j Kxt[m][j] Linv[n][j] as tcgen05.mma.kind::tf32, 128 x 128
//                              tiles, operands staged by TMA (SWIZZLE_128B, 3 x 64 KB stages), accumulators
//                              in TMEM (4 x 128 columns).  Warp-specialised persistent CTA: warp 0 = TMA
//                              producer, warp 1 = MMA issuer (one thread), warps 2-9 = epilogue (tcgen05.ld
//                              32x32b: thread = query row x 64 columns; the tile is never stored).
//                              Lower-triangular Linv: n-tile I contracts k < 128 (I + 1) only.
// Accumulation: rows of Linv oscillate, so |terms| / |result| ~ 1e3 in this contraction and the tensor core's
// truncating FP32 accumulator is the dominant error when it runs over the whole K (measured at N = 4096: 1.2e-2 of
// the posterior variance).  The MMA therefore accumulates only `acc_kblocks` k-blocks of 32 per TMEM buffer and the
// epilogue sums the chunks in registers with round-to-nearest FP32 adds, overlapped with the next chunks' MMAs
// (acc_kblocks = 1: 7.7e-4 max / 2e-4 median of the posterior variance, 15 % slower than unchunked).
// Not a replacement for the FP64 mode where 1e-6 relative variance is required.
#include "internal.cuh"
#include "tc05.cuh"

namespace ipp {

constexpr int TBM = 128, TBN = 128, TBK = 32;        // tile in elements; TBK floats = 128 B = one swizzle row
constexpr int T_STAGES = 3;
constexpr int T_TILE_BYTES = TBM * TBK * 4;          // 16 KB
constexpr int T_STAGE_BYTES = 4 * T_TILE_BYTES;      // A hi, A lo, B hi, B lo
constexpr int T_SMEM = T_STAGES * T_STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
constexpr int T_EPI_WARPS = 8;                       // two per TMEM lane quarter, 64 accumulator columns each
constexpr int T_THREADS = 64 + 32 * T_EPI_WARPS;
constexpr int T_ACC_COLS = 128;                      // FP32 accumulator columns per tile
constexpr int T_NBUF = 4;                            // TMEM accumulator buffers (chunk pipeline MMA -> epilogue)
constexpr int T_TMEM_COLS = T_NBUF * T_ACC_COLS;     // 512 = all of TMEM (one CTA per SM)
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N = 128, M = 128
constexpr uint32_t T_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TBN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);

__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(T_IDESC), "r"(accumulate) : "memory");
}
// K-major operand tile [128 rows][128 B], SWIZZLE_128B: 8-row groups 1024 B apart (SBO), LBO unused (1),
// descriptor version 1 (Blackwell), layout type 2 = SWIZZLE_128B   (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
struct Tf32Args {
    double* partial; int64_t ldp;      // [2 T][ldp] per-(n-tile, column half, query) partial sums of squares
    int Mc, N, T, Mtiles, GM;
    int KC;                            // k-blocks (of 32) accumulated inside the tensor core before the epilogue takes over
};

// item -> (n-tile I heaviest first inside an L2-sized group of GM m-tiles, m-tile); identical in all three roles
__device__ __forceinline__ bool tf32_item(const Tf32Args& p, int it, int& m0, int& I, int& kblocks) {
    const int item = (int)blockIdx.x + it * (int)gridDim.x;
    if (item >= p.T * p.Mtiles) return false;
    const int grp = item / (p.T * p.GM);
    const int r = item - grp * p.T * p.GM;
    const int gm = (p.Mtiles - grp * p.GM) < p.GM ? (p.Mtiles - grp * p.GM) : p.GM;
    I = p.T - 1 - r / gm;
    m0 = (grp * p.GM + r % gm) * TBM;
    const int kend = ((I + 1) * TBN < p.N) ? (I + 1) * TBN : p.N;
    kblocks = (kend + TBK - 1) / TBK;
    return true;
}

__global__ void __launch_bounds__(T_THREADS, 1)
predict_var_tf32_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                        const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                        const Tf32Args p) {
    extern __shared__ unsigned char t_smem_raw[];
    const uint32_t raw = smem_addr(t_smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;                       // SWIZZLE_128B tiles need 1024-B alignment
    const uint32_t bars = tiles + T_STAGES * T_STAGE_BYTES;              // 8-byte mbarriers
    const uint32_t full0 = bars, empty0 = bars + 8 * T_STAGES;
    const uint32_t tfull0 = bars + 16 * T_STAGES, tempty0 = tfull0 + 8 * T_NBUF;
    const uint32_t tmem_slot = tempty0 + 8 * T_NBUF;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < T_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < T_NBUF; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 32 * T_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(T_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

    if (warp == 0) {                                                     // ===== TMA producer (one elected lane issues)
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_hi) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_lo) : "memory");
        }
        uint32_t stage = 0, phase = 0;
        int m0, I, kblocks;
        for (int it = 0; tf32_item(p, it, m0, I, kblocks); ++it) {
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                if (elect_one()) {
                    const uint32_t full = full0 + 8 * stage, dst = tiles + stage * T_STAGE_BYTES;
                    mbar_expect_tx(full, T_STAGE_BYTES);
                    tma_load_2d(dst, &map_a_hi, kb * TBK, m0, full);
                    tma_load_2d(dst + T_TILE_BYTES, &map_a_lo, kb * TBK, m0, full);
                    tma_load_2d(dst + 2 * T_TILE_BYTES, &map_b_hi, kb * TBK, I * TBN, full);
                    tma_load_2d(dst + 3 * T_TILE_BYTES, &map_b_lo, kb * TBK, I * TBN, full);
                }
                __syncwarp();
                if (++stage == T_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {                                              // ===== MMA issuer (one elected lane issues)
        uint32_t stage = 0, phase = 0, buf = 0, buf_phase = 0;
        int m0, I, kblocks;
        for (int it = 0; tf32_item(p, it, m0, I, kblocks); ++it) {
            for (int kb0 = 0; kb0 < kblocks; kb0 += p.KC) {              // one accumulation chunk per TMEM buffer
                const int kb1 = (kb0 + p.KC < kblocks) ? kb0 + p.KC : kblocks;
                mbar_wait(tempty0 + 8 * buf, buf_phase ^ 1);             // epilogue has drained this buffer
                tc_fence_after();
                const uint32_t d = tmem_base + buf * T_ACC_COLS;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full0 + 8 * stage, phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t base = tiles + stage * T_STAGE_BYTES;
                        const uint64_t a_hi = make_desc(base), a_lo = make_desc(base + T_TILE_BYTES);
                        const uint64_t b_hi = make_desc(base + 2 * T_TILE_BYTES), b_lo = make_desc(base + 3 * T_TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < TBK / 8; ++k) {              // UMMA_K = 8 TF32 = 32 B: +2 in 16-B units
                            const uint64_t adv = (uint64_t)(2 * k);
                            tc_mma_tf32(d, a_lo + adv, b_hi + adv, (kb != kb0) || k != 0);
                            tc_mma_tf32(d, a_hi + adv, b_lo + adv, 1);
                            tc_mma_tf32(d, a_hi + adv, b_hi + adv, 1);
                        }
                        tc_commit(empty0 + 8 * stage);                   // frees the stage when these MMAs retire
                        if (kb == kb1 - 1) tc_commit(tfull0 + 8 * buf);  // chunk complete -> epilogue
                    }
                    __syncwarp();
                    if (++stage == T_STAGES) { stage = 0; phase ^= 1; }
                }
                if (++buf == T_NBUF) { buf = 0; buf_phase ^= 1; }
            }
        }
    } else {                                                             // ===== epilogue warps 2..5
        // The tensor core accumulates in FP32 with truncation; over a long contraction with heavy cancellation
        // (rows of Linv oscillate) that error dominates.  So the MMA only accumulates KC k-blocks per TMEM buffer
        // and the chunks are summed here in registers with round-to-nearest FP32 adds (thread = query row,
        // 128 columns), overlapped with the MMAs of the next chunks.
        const int quarter = warp & 3;                                    // TMEM lanes 32 * (warp % 4) ...
        const int half = (warp - 2) >> 2;                                // accumulator columns [64 half, 64 half + 64)
        constexpr int HC = T_ACC_COLS / 2;
        uint32_t buf = 0, buf_phase = 0;
        int m0, I, kblocks;
        for (int it = 0; tf32_item(p, it, m0, I, kblocks); ++it) {
            float acc[HC];
#pragma unroll
            for (int i = 0; i < HC; ++i) acc[i] = 0.f;
            for (int kb0 = 0; kb0 < kblocks; kb0 += p.KC) {
                mbar_wait(tfull0 + 8 * buf, buf_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + buf * T_ACC_COLS + half * HC + ((uint32_t)(quarter * 32) << 16);
                uint32_t r0[32], r1[32];
                tmem_ld32(taddr, r0);
                tmem_ld32(taddr + 32, r1);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(tempty0 + 8 * buf);                          // buffer is in registers: hand it back early
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    acc[i] += __uint_as_float(r0[i]);
                    acc[32 + i] += __uint_as_float(r1[i]);
                }
                if (++buf == T_NBUF) { buf = 0; buf_phase ^= 1; }
            }
            double q = 0.0;
#pragma unroll
            for (int i = 0; i < HC; ++i) {
                const double v = (double)acc[i];
                q = fma(v, v, q);
            }
            const int row = m0 + quarter * 32 + lane;
            if (row < p.Mc) p.partial[(int64_t)(2 * I + half) * p.ldp + row] = q;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(T_TMEM_COLS) : "memory");
    }
}

__global__ void predict_finalize_tf32_kernel(const double* __restrict__ partial, int64_t ldp, int T, int Mc,
                                             double variance, double noise_add, double* __restrict__ var) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= Mc) return;
    double q = 0.0;
    for (int I = 0; I < T; ++I) q += partial[(int64_t)I * ldp + m];
    var[m] = (variance - q) + noise_add;
}

// a = hi + lo with hi, lo TF32 (10 explicit mantissa bits, round to nearest): hi carries the top 11 significant
// bits, lo the next 11; the MMA ignores the low 13 bits of its FP32 inputs, so both are pre-rounded.
__device__ __forceinline__ float tf32_rn(float f) {
    return __uint_as_float((__float_as_uint(f) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void split_tf32(double a, float& hi, float& lo) {
    hi = tf32_rn((float)a);
    lo = tf32_rn((float)(a - (double)hi));
}

__global__ void tf32_split_kernel(const double* __restrict__ A, int64_t rows, int64_t cols, int64_t lda,
                                  float* __restrict__ hi, float* __restrict__ lo, int64_t ldf) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * ldf) return;
    const int64_t r = idx / ldf, c = idx % ldf;
    float h = 0.f, l = 0.f;
    if (c < cols) split_tf32(A[r * lda + c], h, l);
    hi[idx] = h;
    lo[idx] = l;
}

// Query-major cross-covariance chunk as (hi, lo) + the FP64 posterior mean; warp owns 4 queries, lanes sweep n
template <int D>
__global__ void __launch_bounds__(256) rbf_queries_tf32_kernel(RbfDev k, const double* __restrict__ X, int64_t N,
                                                               const double* __restrict__ alpha,
                                                               const double* __restrict__ Xq, int64_t M,
                                                               float* __restrict__ Khi, float* __restrict__ Klo,
                                                               int64_t ldf, double* __restrict__ mean) {
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
        const double a = alpha[n];
#pragma unroll
        for (int i = 0; i < QPW; ++i) {
            const double kv = rbf_eval_fast<D>(k, q[i], x);
            if (m0 + i < M) {
                float h, l;
                split_tf32(kv, h, l);
                Khi[(m0 + i) * ldf + n] = h;
                Klo[(m0 + i) * ldf + n] = l;
            }
            acc[i] = fma(kv, a, acc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < QPW; ++i) {
        double s = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && m0 + i < M) mean[m0 + i] = s;
    }
}

// ---- host side
// [rows][ldf] float matrix, logical width `cols`: box = 32 floats x 128 rows, SWIZZLE_128B, out-of-bounds -> 0
static int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ldf) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return IPP_E_CUDA; }
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ldf * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)TBK, (cuuint32_t)TBM};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return IPP_E_CUDA; }
    return IPP_OK;
}

constexpr int64_t TF32_CHUNK = 32768;
static int64_t tf32_chunk_rows(int64_t M) { return M < TF32_CHUNK ? round_up(M, TBM) : TF32_CHUNK; }

}  // namespace ipp

using namespace ipp;

extern "C" int ipp_tf32_split(const double* A, int64_t rows, int64_t cols, int64_t lda, float* hi, float* lo,
                              int64_t ldf, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    IPP_CHECK_ARG(A && hi && lo && rows >= 1 && cols >= 1 && lda >= cols && ldf >= cols, "bad arguments");
    const int64_t total = rows * ldf;
    tf32_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(A, rows, cols, lda, hi, lo, ldf);
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

extern "C" int64_t ipp_gp_predict_tf32_workspace(int64_t N, int64_t M) {
    if (N <= 0 || M <= 0) return 16;
    const int64_t ldf = round_up(N, 32), mc = tf32_chunk_rows(M), T = (N + TBN - 1) / TBN;
    return mc * ldf /* hi + lo floats */ + 2 * T * mc + 64;
}

extern "C" int ipp_gp_predict_tf32(const ipp_rbf* kern, const double* X, int64_t N, const double* alpha,
                                   const float* linv_hi, const float* linv_lo, int64_t ldl, int acc_kblocks,
                                   double noise_add, const double* Xq, int64_t M, double* mean, double* var,
                                   double* work, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    IPP_CHECK_ARG(kern && X && alpha && linv_hi && linv_lo && Xq && mean && var && work, "null pointer");
    IPP_CHECK_ARG(N >= 1 && M >= 1 && N < (1LL << 30), "bad shape");
    IPP_CHECK_ARG(acc_kblocks >= 1, "acc_kblocks >= 1");
    IPP_CHECK_ARG(kern->dim == 2 || kern->dim == 3, "dimension must be 2 or 3");
    IPP_CHECK_ARG(ldl >= N && ldl % 4 == 0 && reinterpret_cast<uintptr_t>(linv_hi) % 16 == 0 &&
                  reinterpret_cast<uintptr_t>(linv_lo) % 16 == 0, "linv_hi/lo must be 16-byte aligned with ldl % 4 == 0");
    IPP_OPT_IN_SMEM(predict_var_tf32_kernel, T_SMEM);
    const int64_t ldf = round_up(N, 32), mc_max = tf32_chunk_rows(M);
    const int T = (int)((N + TBN - 1) / TBN);
    float* Khi = reinterpret_cast<float*>(work);
    float* Klo = Khi + mc_max * ldf;
    double* partial = work + mc_max * ldf;
    CUtensorMap map_b_hi, map_b_lo;
    IPP_TRY(make_map(&map_b_hi, linv_hi, N, N, ldl));
    IPP_TRY(make_map(&map_b_lo, linv_lo, N, N, ldl));
    const RbfDev kd = to_dev(*kern);
    const int grid = sm_count();
    for (int64_t s = 0; s < M; s += mc_max) {
        const int64_t mc = (M - s) < mc_max ? (M - s) : mc_max;
        const unsigned qblocks = (unsigned)((mc + 31) / 32);
        if (kern->dim == 2)
            rbf_queries_tf32_kernel<2><<<qblocks, 256, 0, stream>>>(kd, X, N, alpha, Xq + s * 2, mc, Khi, Klo, ldf, mean + s);
        else
            rbf_queries_tf32_kernel<3><<<qblocks, 256, 0, stream>>>(kd, X, N, alpha, Xq + s * 3, mc, Khi, Klo, ldf, mean + s);
        IPP_LAUNCH_CHECK();
        CUtensorMap map_a_hi, map_a_lo;
        IPP_TRY(make_map(&map_a_hi, Khi, mc, N, ldf));
        IPP_TRY(make_map(&map_a_lo, Klo, mc, N, ldf));
        Tf32Args p;
        p.partial = partial; p.ldp = mc_max;
        p.Mc = (int)mc; p.N = (int)N; p.T = T; p.Mtiles = (int)((mc + TBM - 1) / TBM);
        {
            const int64_t panel = (int64_t)TBM * ldf * 8;                // hi + lo bytes of one m-tile's panel
            int64_t gm = (48LL << 20) / panel;                           // ~48 MB of the 126 MB L2 per group
            if (gm < 1) gm = 1;
            if (gm > p.Mtiles) gm = p.Mtiles;
            p.GM = (int)gm;
        }
        p.KC = acc_kblocks;
        const int items = p.T * p.Mtiles;
        profile_mark(0, stream);
        predict_var_tf32_kernel<<<items < grid ? items : grid, T_THREADS, T_SMEM, stream>>>(map_a_hi, map_a_lo, map_b_hi,
                                                                                            map_b_lo, p);
        profile_mark(1, stream);
        IPP_LAUNCH_CHECK();
        predict_finalize_tf32_kernel<<<(unsigned)((mc + 255) / 256), 256, 0, stream>>>(partial, mc_max, 2 * T, (int)mc,
                                                                                      kern->variance, noise_add, var + s);
        IPP_LAUNCH_CHECK();
    }
    return IPP_OK;
}
