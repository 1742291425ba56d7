// a3, "sliced INT8" mode: FP64-grade posterior variance on the INT8 tensor cores (Ozaki-style error-free slicing).
//
//   var[m] = variance - sum_n (sum_{j<=n} Linv[n][j] k(q_m, x_j))^2 + noise_add          (GPy's Cholesky form)
//
// Both operands are written in fixed point against ONE power-of-two scale each (k <= variance for every entry of the
// cross-covariance; |Linv| <= its global maximum) and cut into NS = 6 signed 7-bit slices:
//       x = scale * sum_s x_s 2^(-6 - 7 s),   x_s integer in [-64, 64]                       (42 bits)
// The contraction is then 21 INT8 products (slice pairs with i + j <= 5), each accumulated EXACTLY in INT32 by
// tcgen05.mma.kind::i8 -- no floating-point accumulation error at all, which is what the 3xTF32 mode suffers from
// (rows of Linv oscillate: |terms| / |result| ~ 1e3) -- and the six weight levels w = i + j are combined in FP64 in
// the epilogue.  What is dropped is 2^-42 of the operand scales per entry plus the slice pairs with i + j > 5:
// measured variance error <= ~2e-8 relative at N = 2048..8192 (noise 0.5), i.e. inside the FP64 mode's 1e-6 parity
// tolerance with a wide margin (numpy emulation with exact integers agrees: 2.0e-8).
//
// Kernel: 128 x 64 tiles (6 levels x 64 columns of TMEM), 3 x 72 KB stages, 42 MMAs (128 x 64 x 32) per stage issued by
// one thread, persistent warp-specialised CTA.  Lower-triangular Linv: n-tile I contracts
// k < 64 (I + 1).
// Operand layout: the generators write the planes ALREADY in the shared-memory image of a stage -- tile-major
// [row tile][k-block of 64][plane][rows][64 B] with the SWIZZLE_64B chunk permutation applied (16-byte chunk c of row r
// lives at chunk c ^ ((r >> 1) & 3)) -- so a stage is two contiguous 1-D bulk copies (48 + 24 KB at 6 x 64) instead of
// 1152 64-byte tensor-map rows; with tensor maps the TMA request rate (~2.2 clk per 64-byte row) capped the kernel.
#include "internal.cuh"
#include "tc05.cuh"

namespace ipp {

#ifndef IPP_I8_COLLECTOR
#define IPP_I8_COLLECTOR 1
#endif
constexpr int I_BM = 128, I_BK = 64;                   // tile rows; I_BK int8 = 64 B = one SWIZZLE_64B row
constexpr int I_THREADS = 192;                         // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int I_TMEM_COLS = 512;
constexpr int I_MAX_NS = 7;

// Three instantiations: NS slices per operand (= number of weight levels i + j = 0..NS-1, products NS (NS + 1) / 2), the
// widest tile whose NS INT32 accumulators fit the 512 TMEM columns, and as many stages as fit shared memory:
//   <7, 64, 2>  49-bit operands, 28 products, 7 x 64 = 448 columns   ("i8x7": ~3e-10 relative variance error -- below the
//               FP64 DMMA mode's own deviation from an extended-precision solve; FP64-equivalent for this path)
//   <6, 64, 3>  42-bit operands, 21 products, 6 x 64 = 384 columns   ("i8x6": ~4e-8, FP64-grade: inside the 1e-6 tolerance)
//   <5, 96, 3>  35-bit operands, 15 products, 5 x 96 = 480 columns   ("i8x5": ~4e-6; fewer products AND a wider MMA --
//               faster and 200x more accurate than the 3xTF32 mode)
template <int NS_, int BN_, int STAGES_>
struct I8Cfg {
    static constexpr int NS = NS_, BN = BN_, STAGES = STAGES_;
    static constexpr int A_PLANE = I_BM * I_BK, B_PLANE = BN_ * I_BK;
    static constexpr int A_BYTES = NS_ * A_PLANE, B_BYTES = NS_ * B_PLANE;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int SMEM = STAGES_ * STAGE_BYTES + 1024 + 256;
    static_assert(SMEM <= 232448, "stages must fit the 227 KB of shared memory");
    // instruction descriptor: D = S32, A = B = signed int8, both K-major, N = BN, M = 128
    static constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN_ >> 3) << 17) | ((uint32_t)(I_BM >> 4) << 24);
    static_assert(NS_ * BN_ <= I_TMEM_COLS && BN_ % 32 == 0 && NS_ <= I_MAX_NS, "accumulators must fit TMEM");
};

// COLL: use of the tensor core's A-operand collector buffer.  Consecutive MMAs that share their A slice (one slice of
// the cross-covariance against every Linv slice j <= NS - 1 - i) keep it in the collector instead of re-reading 4 KB of
// shared memory each time -- SS-mode operand reads were the limiter of this kernel (6 KB per 32-clk MMA at N = 64).
//   0 = discard (default), 1 = fill (first of a run), 2 = use, 3 = lastuse          (SASS: .A_KEEP / .A_REUSE)
template <int COLL>
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (COLL == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::i8.collector::a::fill [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else if (COLL == 2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::i8.collector::a::use [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else if (COLL == 3)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::i8.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major operand plane [rows][64 B], SWIZZLE_64B: 8-row groups 512 B apart (SBO), LBO unused (1), version 1,
// layout type 4 = SWIZZLE_64B   (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc64(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}

struct I8Args {
    const int8_t* a_tiles;             // [Mtiles][KB][NS][128][64 B]   cross-covariance digits (swizzled stage images)
    const int8_t* b_tiles;             // [T][KB][NS][BN][64 B]         Linv slices
    int KB;                            // k-blocks of 64 in a row: ceil(N / 64)
    double* partial; int64_t ldp;      // [T][ldp] per-(n-tile, query) sums of squares in units of (sa sb 2^-13)^2
    int Mc, N, T, Mtiles, GM;
};

// byte offset of element (row r of its tile, k) inside a [rows][64 B] SWIZZLE_64B plane of a k-block
__host__ __device__ __forceinline__ int swz64(int r, int kk) { return r * 64 + ((((kk >> 4) ^ (r >> 1)) & 3) << 4) + (kk & 15); }

template <int BN>
__device__ __forceinline__ bool i8_item(const I8Args& p, int it, int& m0, int& I, int& kblocks) {
    const int item = (int)blockIdx.x + it * (int)gridDim.x;
    if (item >= p.T * p.Mtiles) return false;
    const int grp = item / (p.T * p.GM);
    const int r = item - grp * p.T * p.GM;
    const int gm = (p.Mtiles - grp * p.GM) < p.GM ? (p.Mtiles - grp * p.GM) : p.GM;
    I = p.T - 1 - r / gm;
    m0 = (grp * p.GM + r % gm) * I_BM;
    const int kend = ((I + 1) * BN < p.N) ? (I + 1) * BN : p.N;      // rows of this n-tile are zero beyond their diagonal
    kblocks = (kend + I_BK - 1) / I_BK;
    return true;
}

template <class C>
__global__ void __launch_bounds__(I_THREADS, 1)
predict_var_i8_kernel(const I8Args p) {
    constexpr int I_NS = C::NS, I_BN = C::BN, I_A_PLANE = C::A_PLANE, I_B_PLANE = C::B_PLANE, I_A_BYTES = C::A_BYTES;
    constexpr int I_STAGE_BYTES = C::STAGE_BYTES, I_LEVELS = C::NS, I_STAGES = C::STAGES;
    extern __shared__ unsigned char i_smem_raw[];
    const uint32_t raw = smem_addr(i_smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;
    const uint32_t bars = tiles + I_STAGES * I_STAGE_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 8 * I_STAGES;
    const uint32_t tfull = bars + 16 * I_STAGES, tempty = tfull + 8;
    const uint32_t tmem_slot = tempty + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < I_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tfull, 1);
        mbar_init(tempty, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(I_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

    if (warp == 0) {                                                     // ===== TMA producer (converged warp, one elected lane issues)
        uint32_t stage = 0, phase = 0;
        int m0, I, kblocks;
        for (int it = 0; i8_item<I_BN>(p, it, m0, I, kblocks); ++it) {
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                if (elect_one()) {
                    const uint32_t full = full0 + 8 * stage, dst = tiles + stage * I_STAGE_BYTES;
                    mbar_expect_tx(full, I_STAGE_BYTES);
                    bulk_load(dst, p.a_tiles + ((int64_t)(m0 / I_BM) * p.KB + kb) * I_A_BYTES, I_A_BYTES, full);
                    bulk_load(dst + I_A_BYTES, p.b_tiles + ((int64_t)I * p.KB + kb) * C::B_BYTES, C::B_BYTES, full);
                }
                __syncwarp();
                if (++stage == I_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {                                              // ===== MMA issuer (converged warp, one elected lane issues)
        uint32_t stage = 0, phase = 0;
        int m0, I, kblocks;
        for (int it = 0; i8_item<I_BN>(p, it, m0, I, kblocks); ++it) {
            mbar_wait(tempty, (it & 1) ^ 1);                             // epilogue has drained the accumulators
            tc_fence_after();
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(full0 + 8 * stage, phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t base = tiles + stage * I_STAGE_BYTES;
                    const uint32_t fresh = (kb == 0) ? 0u : 1u;          // first k-block of the tile overwrites the levels
                    uint64_t ad[I_NS], bd[I_NS];
#pragma unroll
                    for (int i = 0; i < I_NS; ++i) {
                        ad[i] = make_desc64(base + i * I_A_PLANE);
                        bd[i] = make_desc64(base + I_A_BYTES + i * I_B_PLANE);
                    }
#pragma unroll
                    for (int ks = 0; ks < I_BK / 32; ++ks) {             // UMMA_K = 32 int8 = 32 B: +2 in 16-B units
#pragma unroll
                        for (int i = 0; i < I_NS; ++i) {
#pragma unroll
                            for (int j = 0; j < I_NS - i; ++j) {
                                // level w = i + j: its first product of the tile is (i = 0, j = w) of ks = 0
                                const uint32_t accumulate = (ks == 0 && i == 0) ? fresh : 1u;
                                const uint32_t dcol = tmem_base + (i + j) * I_BN;
                                const uint64_t a = ad[i] + (uint64_t)(2 * ks), b = bd[j] + (uint64_t)(2 * ks);
                                constexpr bool kReuse = IPP_I8_COLLECTOR != 0;
                                if (!kReuse || I_NS - i == 1) tc_mma_i8<0>(dcol, a, b, C::IDESC, accumulate);
                                else if (j == 0) tc_mma_i8<1>(dcol, a, b, C::IDESC, accumulate);
                                else if (j == I_NS - i - 1) tc_mma_i8<3>(dcol, a, b, C::IDESC, accumulate);
                                else tc_mma_i8<2>(dcol, a, b, C::IDESC, accumulate);
                            }
                        }
                    }
                    tc_commit(empty0 + 8 * stage);
                    if (kb == kblocks - 1) tc_commit(tfull);
                }
                __syncwarp();
                if (++stage == I_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {                                                             // ===== epilogue warps 2..5: thread = query row
        const int quarter = warp & 3;
        int m0, I, kblocks;
        for (int it = 0; i8_item<I_BN>(p, it, m0, I, kblocks); ++it) {
            mbar_wait(tfull, it & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
            double q = 0.0;
#pragma unroll 1
            for (int half = 0; half < I_BN / 32; ++half) {
                double v[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] = 0.0;
#pragma unroll
                for (int w = 0; w < I_LEVELS; ++w) {
                    uint32_t r[32];
                    tmem_ld32(taddr + w * I_BN + half * 32, r);
                    tmem_ld_wait();
                    const double wt = 1.0 / (double)(1ull << (7 * w));   // level w carries weight 2^(-7 w)
#pragma unroll
                    for (int c = 0; c < 32; ++c) v[c] = fma((double)(int)r[c], wt, v[c]);   // exact: |D| < 2^31
                }
#pragma unroll
                for (int c = 0; c < 32; ++c) q = fma(v[c], v[c], q);
            }
            tc_fence_before();
            mbar_arrive(tempty);
            const int row = m0 + quarter * 32 + lane;
            if (row < p.Mc) p.partial[(int64_t)I * p.ldp + row] = q;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(I_TMEM_COLS) : "memory");
    }
}

__global__ void predict_finalize_i8_kernel(const double* __restrict__ partial, int64_t ldp, int T, int Mc,
                                           double unit2, double variance, double noise_add, double* __restrict__ var) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= Mc) return;
    double q = 0.0;
    for (int I = 0; I < T; ++I) q += partial[(int64_t)I * ldp + m];
    var[m] = (variance - q * unit2) + noise_add;
}

// x / scale in (-1, 1)  ->  ns signed 7-bit slices, round to nearest at every level (|slice| <= 64)
__device__ __forceinline__ void slice_signed(double x, double inv_scale, int ns, int (&s)[I_MAX_NS]) {
    double r = x * inv_scale, w = 64.0;
#pragma unroll
    for (int k = 0; k < I_MAX_NS; ++k) {
        if (k >= ns) { s[k] = 0; continue; }
        const double t = rint(r * w);
        s[k] = (int)t;
        r = fma(-t, 1.0 / w, r);           // exact: w is a power of two, t has <= 8 significant bits
        w *= 128.0;
    }
}

// tiles[row tile][k-block][slice][bn][64 B] of A[rows][lda] (swizzled stage images); rows >= `rows` and columns >= `cols`
// of the padded extent are zeroed
__global__ void i8_slice_kernel(const double* __restrict__ A, int64_t rows, int64_t cols, int64_t lda, double inv_scale,
                                int ns, int bn, int64_t KB, int64_t rows_pad, int8_t* __restrict__ tiles) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows_pad * KB * 64) return;
    const int64_t r = idx / (KB * 64), c = idx % (KB * 64);
    int s[I_MAX_NS] = {0, 0, 0, 0, 0, 0, 0};
    if (r < rows && c < cols) slice_signed(A[r * lda + c], inv_scale, ns, s);
    const int64_t tile = r / bn, kb = c / 64;
    const int64_t base = ((tile * KB + kb) * ns) * ((int64_t)bn * 64) + swz64((int)(r % bn), (int)(c % 64));
#pragma unroll
    for (int k = 0; k < I_MAX_NS; ++k)
        if (k < ns) tiles[base + (int64_t)k * bn * 64] = (int8_t)s[k];
}

// Cross-covariance chunk as NS int8 planes [NS][Mc][ldp] + the FP64 posterior mean.  k >= 0, so its 7 NS-bit fixed-point
// value is cut into plain base-128 digits in [0, 127] with integer shifts (the signed round-to-nearest slicing of
// slice_signed costs as many FP64-pipe instructions as the exp itself).  A warp owns QPW queries; lane l handles observations
// 4 l .. 4 l + 3 of every 128-block so that each plane store is a packed 32-bit word (128 B per warp and plane).
// queries per warp: every observation (x, alpha: three loads) is reused by I8_QPW queries.  With 2 the kernel sat on the loads
// (ncu r2a: long-scoreboard 5.0 stalls per issue, FP64 pipe 37 %, 1.4 TB/s of plane stores)
constexpr int I8_QPW = 4;

template <int D, int I_NS>
__global__ void __launch_bounds__(256) rbf_queries_i8_kernel(RbfDev k, const double* __restrict__ X, int64_t N,
                                                             const double* __restrict__ alpha,
                                                             const double* __restrict__ Xq, int64_t M, double inv_scale,
                                                             int8_t* __restrict__ tiles, int64_t KB,
                                                             double* __restrict__ mean) {
    constexpr int QPW = I8_QPW;
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
    for (int64_t n0 = 4 * lane; n0 < KB * 64; n0 += 128) {
        uint32_t packed[QPW][I_NS];
#pragma unroll
        for (int i = 0; i < QPW; ++i)
#pragma unroll
            for (int s = 0; s < I_NS; ++s) packed[i][s] = 0u;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int64_t n = n0 + e;
            if (n < N) {
                double x[D];
#pragma unroll
                for (int d = 0; d < D; ++d) x[d] = X[n * D + d];
                const double a = alpha[n];
#pragma unroll
                for (int i = 0; i < QPW; ++i) {
                    const double kv = rbf_eval<D>(k, q[i], x);              // exact exp: the mean of this mode is the FP64 mode's
                    acc[i] = fma(kv, a, acc[i]);
                    // kv / scale in [0, 1)  ->  X = rint(. 2^(7 NS));  digit s = bits [7 (NS - 1 - s), 7 (NS - s))
                    constexpr unsigned long long TOP = 1ull << (7 * I_NS);
                    unsigned long long X = (unsigned long long)__double2ll_rn(kv * inv_scale * (double)TOP);
                    if (X > TOP - 1) X = TOP - 1;
#pragma unroll
                    for (int s = 0; s < I_NS; ++s)
                        packed[i][s] |= (uint32_t)((X >> (7 * (I_NS - 1 - s))) & 127ull) << (8 * e);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < QPW; ++i)
            if (m0 + i < M) {
#pragma unroll
                for (int s = 0; s < I_NS; ++s)
                    *reinterpret_cast<uint32_t*>(tiles + ((((m0 + i) / I_BM) * KB + n0 / 64) * I_NS + s) * (int64_t)(I_BM * 64) +
                                                 swz64((int)((m0 + i) % I_BM), (int)(n0 % 64))) = packed[i][s];
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

constexpr int64_t I8_CHUNK = 32768;
static int64_t i8_chunk_rows(int64_t M) { return M < I8_CHUNK ? round_up(M, I_BM) : I8_CHUNK; }
static double pow2_at_least(double x) {
    double s = 1.0;
    while (s < x) s *= 2.0;
    while (s * 0.5 >= x && s > 1e-300) s *= 0.5;
    return s;
}

}  // namespace ipp

using namespace ipp;

static int bn_of(int n_slices) { return n_slices == 5 ? 96 : 64; }
static bool ns_ok(int n_slices) { return n_slices >= 5 && n_slices <= 7; }

extern "C" int64_t ipp_i8_slice_bytes(int64_t rows, int64_t cols, int n_slices) {
    if (rows <= 0 || cols <= 0 || !ns_ok(n_slices)) return 0;
    const int64_t bn = bn_of(n_slices);
    return round_up(rows, bn) * ((cols + 63) / 64) * 64 * n_slices;
}

extern "C" int ipp_i8_slice(const double* A, int64_t rows, int64_t cols, int64_t lda, double scale, int n_slices,
                            int8_t* tiles, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    IPP_CHECK_ARG(A && tiles && rows >= 1 && cols >= 1 && lda >= cols, "bad arguments");
    IPP_CHECK_ARG(reinterpret_cast<uintptr_t>(tiles) % 16 == 0, "tiles must be 16-byte aligned");
    IPP_CHECK_ARG(scale > 0, "scale must be a positive power of two >= max |A|");
    IPP_CHECK_ARG(ns_ok(n_slices), "n_slices must be 5, 6 or 7");
    const int bn = bn_of(n_slices);
    const int64_t KB = (cols + 63) / 64, rows_pad = round_up(rows, bn), total = rows_pad * KB * 64;
    i8_slice_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(A, rows, cols, lda, 1.0 / scale, n_slices, bn, KB, rows_pad, tiles);
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

extern "C" int64_t ipp_gp_predict_i8_workspace(int64_t N, int64_t M) {
    if (N <= 0 || M <= 0) return 16;
    const int64_t KB = (N + 63) / 64, mc = i8_chunk_rows(M), T = (N + 63) / 64;           // sized for the 6-slice layout
    return (I_MAX_NS * mc * KB * 64 + 7) / 8 + T * mc + 64;
}

namespace ipp {

template <class C>
static int predict_i8_impl(const ipp_rbf* kern, const double* X, int64_t N, const double* alpha, const int8_t* linv_tiles,
                           double linv_scale, double noise_add, const double* Xq, int64_t M, double* mean,
                           double* var, double* work, cudaStream_t stream) {
    constexpr int NS = C::NS, BN = C::BN;
    IPP_OPT_IN_SMEM(predict_var_i8_kernel<C>, C::SMEM);
    const int64_t KB = (N + 63) / 64, mc_max = i8_chunk_rows(M);
    const int T = (int)((N + BN - 1) / BN);
    int8_t* tiles = reinterpret_cast<int8_t*>(work);
    double* partial = work + (I_MAX_NS * mc_max * KB * 64 + 7) / 8;
    const double sa = pow2_at_least(kern->variance * (1.0 + 1e-12));      // k(q, x) <= variance for every pair
    // v = D * sa * sb * 2^-13 with D = sum_w D_w 2^(-7 w): the epilogue sums squares of D, so one constant at the end
    // (cross-covariance digits carry 2^(-7 - 7 s), Linv slices 2^(-6 - 7 s): 2^-13 in front of level 0)
    const double unit = sa * linv_scale / 8192.0;
    const RbfDev kd = to_dev(*kern);
    const int grid = sm_count();
    for (int64_t s = 0; s < M; s += mc_max) {
        const int64_t mc = (M - s) < mc_max ? (M - s) : mc_max;
        const unsigned qblocks = (unsigned)((mc + 8 * I8_QPW - 1) / (8 * I8_QPW));
        if (kern->dim == 2)
            rbf_queries_i8_kernel<2, NS><<<qblocks, 256, 0, stream>>>(kd, X, N, alpha, Xq + s * 2, mc, 1.0 / sa, tiles, KB, mean + s);
        else
            rbf_queries_i8_kernel<3, NS><<<qblocks, 256, 0, stream>>>(kd, X, N, alpha, Xq + s * 3, mc, 1.0 / sa, tiles, KB, mean + s);
        IPP_LAUNCH_CHECK();
        I8Args p;
        p.a_tiles = tiles; p.b_tiles = linv_tiles; p.KB = (int)KB;
        p.partial = partial; p.ldp = mc_max;
        p.Mc = (int)mc; p.N = (int)N; p.T = T; p.Mtiles = (int)((mc + I_BM - 1) / I_BM);
        {
            const int64_t panel = (int64_t)I_BM * KB * 64 * NS;          // bytes of one m-tile's planes
            int64_t gm = (48LL << 20) / panel;
            if (gm < 1) gm = 1;
            if (gm > p.Mtiles) gm = p.Mtiles;
            p.GM = (int)gm;
        }
        const int items = p.T * p.Mtiles;
        profile_mark(0, stream);
        predict_var_i8_kernel<C><<<items < grid ? items : grid, I_THREADS, C::SMEM, stream>>>(p);
        profile_mark(1, stream);
        IPP_LAUNCH_CHECK();
        predict_finalize_i8_kernel<<<(unsigned)((mc + 255) / 256), 256, 0, stream>>>(partial, mc_max, T, (int)mc, unit * unit,
                                                                                    kern->variance, noise_add, var + s);
        IPP_LAUNCH_CHECK();
    }
    return IPP_OK;
}

}  // namespace ipp

extern "C" int ipp_gp_predict_i8(const ipp_rbf* kern, const double* X, int64_t N, const double* alpha,
                                 const int8_t* linv_tiles, double linv_scale, int n_slices, double noise_add,
                                 const double* Xq, int64_t M, double* mean, double* var, double* work, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    IPP_CHECK_ARG(kern && X && alpha && linv_tiles && Xq && mean && var && work, "null pointer");
    IPP_CHECK_ARG(N >= 1 && M >= 1 && N <= 40000, "bad shape (exact INT32 accumulation: 6 pairs x 127 x 64 x N < 2^31)");
    IPP_CHECK_ARG(kern->dim == 2 || kern->dim == 3, "dimension must be 2 or 3");
    IPP_CHECK_ARG(reinterpret_cast<uintptr_t>(linv_tiles) % 16 == 0 && reinterpret_cast<uintptr_t>(work) % 16 == 0,
                  "linv tiles and work must be 16-byte aligned");
    IPP_CHECK_ARG(linv_scale > 0 && kern->variance > 0, "scales must be positive");
    IPP_CHECK_ARG(ns_ok(n_slices), "n_slices must be 5, 6 or 7 (the count linv_tiles was cut with)");
    IPP_CHECK_ARG(N <= 37000 || n_slices < 7, "n_slices = 7 needs N <= 37000 (exact INT32 accumulation of 7 pairs per level)");
    if (n_slices == 7)
        return predict_i8_impl<I8Cfg<7, 64, 2>>(kern, X, N, alpha, linv_tiles, linv_scale, noise_add, Xq, M, mean, var, work, stream);
    if (n_slices == 6)
        return predict_i8_impl<I8Cfg<6, 64, 3>>(kern, X, N, alpha, linv_tiles, linv_scale, noise_add, Xq, M, mean, var, work, stream);
    return predict_i8_impl<I8Cfg<5, 96, 3>>(kern, X, N, alpha, linv_tiles, linv_scale, noise_add, Xq, M, mean, var, work, stream);
}
