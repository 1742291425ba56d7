// FP64 tensor-core (DMMA.8x8x4) tile machinery shared by the generic GEMM and the fused
// posterior-variance kernel.  sm_100a lowers every mma.sync ... f64 shape to DMMA.8x8x4,
// so m8n8k4 is issued directly.
//
// CTA tile 128 x 128 x BK (BK = 32), 8 warps as 2 (M) x 4 (N), warp tile 64 x 32 = 8 x 4 DMMA tiles
// (64 FP64 accumulators / thread).  Operands are both "K-major" (row-major [rows][K]):
//   A tile [128][BK] and B tile [128][BK] doubles, moved by 16-byte cp.async with zero-fill at the
//   edges, 3 stages x 64 KB = 192 KB of dynamic shared memory.
// Shared-memory layout: 16-byte chunk c of row r lives at chunk position c ^ ((r & 1) << 2).
//   A warp's LDS.128 fragment read touches rows (g, g+1) x chunks (4*kg + t), i.e. two rows x
//   64 B; the swizzle sends odd rows to the other 64-byte half, so each quarter-warp reads
//   128 distinct bytes modulo the 128-byte bank window: conflict-free with no padding.
// K-permutation: within each group of 8 k's, DMMA step s (0/1) contracts physical
//   k = 8*kg + 2*t + s for quad-lane t.  A and B use the same map, so the sum is unchanged
//   and one LDS.128 feeds two DMMA steps.
// Measured on B200 (r1, N=4096 x 2^20 queries): BK16/4 stages 32.0 TFLOP/s -> lean loader 32.1 ->
//   + phase-shifted prefetch 32.6 -> BK32/3 stages 33.7 -> (r2) whole-stage / all-rows fast path of the loader 34.0
//   (cuBLAS DGEMM 35.5, DMMA peak 37.2).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ipp {

#ifndef IPP_BK
#define IPP_BK 32
#endif
#ifndef IPP_STAGES
#define IPP_STAGES 3
#endif
#ifndef IPP_PHASE_SHIFT
#define IPP_PHASE_SHIFT 1
#endif
constexpr int BM = 128, BN = 128, BK = IPP_BK;
constexpr int STAGES = IPP_STAGES;
constexpr int ROW_BYTES = BK * 8;                // 128 B (BK=16) or 256 B (BK=32)
constexpr int CHUNKS_PER_ROW = ROW_BYTES / 16;
constexpr int TILE_BYTES = BM * ROW_BYTES;
constexpr int STAGE_BYTES = 2 * TILE_BYTES;      // A + B
constexpr int GEMM_SMEM = STAGES * STAGE_BYTES;

// Warp layout of the 128 x 128 CTA tile: WARPS_M x WARPS_N warps, warp tile WM x WN.
template <int WARPS_M_, int WARPS_N_>
struct TileCfg {
    static constexpr int WARPS_M = WARPS_M_, WARPS_N = WARPS_N_;
    static constexpr int THREADS = 32 * WARPS_M_ * WARPS_N_;
    static constexpr int WM = BM / WARPS_M_, WN = BN / WARPS_N_;
    static constexpr int MI = WM / 8, NI = WN / 8;
};
using Cfg8 = TileCfg<2, 4>;      // 8 warps, 64 x 32 warp tiles, 64 accumulators / thread
using Cfg16 = TileCfg<4, 4>;     // 16 warps, 32 x 32 warp tiles, 32 accumulators / thread
#ifndef IPP_TILE_CFG
#define IPP_TILE_CFG Cfg8
#endif
using Cfg = IPP_TILE_CFG;
constexpr int GEMM_THREADS = Cfg::THREADS;
constexpr int WM = Cfg::WM, WN = Cfg::WN;
constexpr int MI = Cfg::MI, NI = Cfg::NI;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ double2 lds128(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Per-thread state for streaming [128 rows][BK k] tiles of a row-major (K-major) matrix into swizzled
// shared memory with 16-byte cp.async.  Everything that does not depend on k is computed once per
// main loop: thread tid owns chunk column (tid % CHUNKS_PER_ROW) of rows r, r + RSTEP, r + 2 RSTEP, ...
// Rows >= nrows and k >= kend are zero-filled.  G must be 16-byte aligned, ld and kbeg even.
constexpr int LOADS_PER_THREAD = (BM * CHUNKS_PER_ROW) / GEMM_THREADS;
constexpr int RSTEP = GEMM_THREADS / CHUNKS_PER_ROW;     // row distance between a thread's consecutive chunks

struct TileLoader {
    const double* base;    // &G[row0 + r][kcol]
    int64_t rstep;         // RSTEP * ld elements
    uint32_t dst;          // byte offset of the thread's first chunk inside a tile
    int kcol;              // first k of the thread's chunk column
    unsigned valid;        // bit i: row r + i * RSTEP is inside the matrix
    bool fast;             // every row of every lane of this warp is inside the matrix and the row step fits 32 bits: a
                           // whole stage (no ragged k) then needs neither row tests nor byte counts -- 3 instructions per
                           // copy instead of 14 (the loader's burst is the stretch of a k-step in which a warp issues no DMMA)
    unsigned rstep32;

    __device__ __forceinline__ void init(const double* G, int64_t ld, int row0, int nrows, int tid) {
        const int r = tid / CHUNKS_PER_ROW, ch = tid % CHUNKS_PER_ROW;
        kcol = ch * 2;
        base = G + (int64_t)(row0 + r) * ld + kcol;
        rstep = (int64_t)RSTEP * ld;
        dst = r * ROW_BYTES + ((ch ^ ((r & 1) << 2)) << 4);      // RSTEP is even: (row & 1) == (r & 1) for all i
        valid = 0;
#pragma unroll
        for (int i = 0; i < LOADS_PER_THREAD; ++i)
            if (row0 + r + i * RSTEP < nrows) valid |= 1u << i;
        rstep32 = (unsigned)rstep;
        fast = __all_sync(0xffffffffu, valid == (1u << LOADS_PER_THREAD) - 1u) && rstep * LOADS_PER_THREAD < (1LL << 31);
    }
    __device__ __forceinline__ void load(uint32_t smem_tile, int k0, int kend) const {
        if (fast && k0 + BK <= kend) {
            const double* src = base + k0;
#pragma unroll
            for (int i = 0; i < LOADS_PER_THREAD; ++i)
                cp_async16(smem_tile + dst + i * (RSTEP * ROW_BYTES), src + (unsigned)i * rstep32, 16);
            return;
        }
        const int rem = kend - (k0 + kcol);
        const int bytes = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
        const double* src = base + k0;
#pragma unroll
        for (int i = 0; i < LOADS_PER_THREAD; ++i) {
            const bool ok = ((valid >> i) & 1u) && bytes > 0;
            cp_async16(smem_tile + dst + i * (RSTEP * ROW_BYTES), ok ? (const void*)(src + i * rstep) : (const void*)base,
                       ok ? bytes : 0);
        }
    }
};

// One 8-wide k group of a stage: acc[i][j][e] holds C[wm0 + 8i + g][wn0 + 8j + 2t + e]
__device__ __forceinline__ void compute_kgroup(uint32_t sA, uint32_t sB, double (&acc)[MI][NI][2],
                                               int wm0, int wn0, int g, int t, int kg) {
    double2 a[MI], b[NI];
    const uint32_t chunk = (uint32_t)(((4 * kg + t) ^ ((g & 1) << 2)) << 4);
#pragma unroll
    for (int i = 0; i < MI; ++i) a[i] = lds128(sA + (wm0 + 8 * i + g) * ROW_BYTES + chunk);
#pragma unroll
    for (int j = 0; j < NI; ++j) b[j] = lds128(sB + (wn0 + 8 * j + g) * ROW_BYTES + chunk);
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
}

// Multi-stage main loop over k in [kbeg, kend): accumulates A[m0.., k] * B[n0.., k]^T into acc.
// All threads of the CTA must call it; on return every cp.async has landed and been consumed,
// and a __syncthreads() separates the last shared-memory read from the caller's next write.
//
// Phase shift: warps w and w + 4 share an SM sub-partition (w % 4).  One warp alone can saturate that
// sub-partition's DMMA pipe (ptxas spaces DMMAs 16 cycles apart), so the pipe only idles when BOTH of
// its warps are outside the DMMA stream.  The prefetch block (address math + cp.async issue) is
// therefore issued at the top of the k-step by the first half of the warps and in the middle of
// the k-step by the second half, instead of by everyone right after the barrier.
__device__ __forceinline__ void gemm_mainloop(uint32_t smem_base, const double* __restrict__ A, int64_t lda,
                                              int m0, int M, const double* __restrict__ B, int64_t ldb, int n0,
                                              int N, int kbeg, int kend, double (&acc)[MI][NI][2], int tid,
                                              int wm0, int wn0, int g, int t) {
    const int KT = (kend - kbeg + BK - 1) / BK;
    TileLoader la, lb;
    la.init(A, lda, m0, M, tid);
    lb.init(B, ldb, n0, N, tid);
#if IPP_PHASE_SHIFT
    const bool late_loader = ((tid >> 5) & 4) != 0;
#else
    const bool late_loader = false;
#endif
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) {
            la.load(smem_base + s * STAGE_BYTES, kbeg + s * BK, kend);
            lb.load(smem_base + s * STAGE_BYTES + TILE_BYTES, kbeg + s * BK, kend);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nk = kt + STAGES - 1;
        const uint32_t nst = smem_base + (nk % STAGES) * STAGE_BYTES;
        const uint32_t st = smem_base + (kt % STAGES) * STAGE_BYTES;
        if (!late_loader) {
            if (nk < KT) {
                la.load(nst, kbeg + nk * BK, kend);
                lb.load(nst + TILE_BYTES, kbeg + nk * BK, kend);
            }
            cp_async_commit();
        }
#pragma unroll
        for (int kg = 0; kg < BK / 8; ++kg) {
            if (kg == BK / 16 && late_loader) {
                if (nk < KT) {
                    la.load(nst, kbeg + nk * BK, kend);
                    lb.load(nst + TILE_BYTES, kbeg + nk * BK, kend);
                }
                cp_async_commit();
            }
            compute_kgroup(st, st + TILE_BYTES, acc, wm0, wn0, g, t, kg);
        }
    }
    cp_async_wait<0>();
    __syncthreads();
}

}  // namespace ipp
