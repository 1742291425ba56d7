// FP64 tensor-core (DMMA.8x8x4) tile machinery shared by the generic GEMM and the fused
// posterior-variance kernel.  sm_100a lowers every mma.sync ... f64 shape to DMMA.8x8x4,
// so m8n8k4 is issued directly.
//
// CTA tile 128 x 128 x 16, 8 warps as 2 (M) x 4 (N), warp tile 64 x 32 = 8 x 4 DMMA tiles
// (64 FP64 accumulators / thread).  Operands are both "K-major" (row-major [rows][K]):
//   A tile [128][16] and B tile [128][16] doubles, 128 B per row, moved by 16-byte cp.async
//   with zero-fill at the edges, 4 stages (128 KB of dynamic shared memory).
// Shared-memory layout: 16-byte chunk c of row r lives at chunk position c ^ ((r & 1) << 2).
//   A warp's LDS.128 fragment read touches rows (g, g+1) x chunks (4*kg + t), i.e. two rows x
//   64 B; the swizzle sends odd rows to the other 64-byte half, so each quarter-warp reads
//   128 distinct contiguous bytes: conflict-free with no padding.
// K-permutation: within each group of 8 k's, DMMA step s (0/1) contracts physical
//   k = 8*kg + 2*t + s for quad-lane t.  A and B use the same map, so the sum is unchanged
//   and one LDS.128 feeds two DMMA steps.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ipp {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int STAGES = 4;
constexpr int GEMM_THREADS = 256;
constexpr int WM = 64, WN = 32;
constexpr int MI = WM / 8, NI = WN / 8;
constexpr int TILE_BYTES = BM * BK * 8;          // 16 KB
constexpr int STAGE_BYTES = 2 * TILE_BYTES;      // A + B
constexpr int GEMM_SMEM = STAGES * STAGE_BYTES;  // 128 KB

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

// One [128 rows][16 k] tile of a row-major matrix -> swizzled shared memory.
// Rows >= nrows and k >= kend are zero-filled.  G must be 16-byte aligned, ld and k0 even.
__device__ __forceinline__ void load_tile(uint32_t smem_tile, const double* __restrict__ G, int64_t ld,
                                          int row0, int nrows, int k0, int kend, int tid) {
#pragma unroll
    for (int i = 0; i < (BM * BK * 8 / 16) / GEMM_THREADS; ++i) {
        const int c = tid + i * GEMM_THREADS;
        const int row = c >> 3, ch = c & 7;
        const int grow = row0 + row;
        const int k = k0 + ch * 2;
        int bytes = 0;
        const double* src = G;
        if (grow < nrows) {
            const int rem = kend - k;
            bytes = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
            if (bytes) src = G + (int64_t)grow * ld + k;
        }
        const uint32_t dst = smem_tile + row * (BK * 8) + ((ch ^ ((row & 1) << 2)) << 4);
        cp_async16(dst, src, bytes);
    }
}

// acc[i][j][e] holds C[wm0 + 8i + g][wn0 + 8j + 2t + e]
__device__ __forceinline__ void compute_stage(uint32_t sA, uint32_t sB, double (&acc)[MI][NI][2],
                                              int wm0, int wn0, int g, int t) {
#pragma unroll
    for (int kg = 0; kg < 2; ++kg) {
        double2 a[MI], b[NI];
        const uint32_t chunk = (uint32_t)(((4 * kg + t) ^ ((g & 1) << 2)) << 4);
#pragma unroll
        for (int i = 0; i < MI; ++i) a[i] = lds128(sA + (wm0 + 8 * i + g) * (BK * 8) + chunk);
#pragma unroll
        for (int j = 0; j < NI; ++j) b[j] = lds128(sB + (wn0 + 8 * j + g) * (BK * 8) + chunk);
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
    }
}

// Multi-stage main loop over k in [kbeg, kend): accumulates A[m0.., k] * B[n0.., k]^T into acc.
// All threads of the CTA must call it; on return every cp.async has landed and been consumed,
// and a __syncthreads() separates the last shared-memory read from the caller's next write.
__device__ __forceinline__ void gemm_mainloop(uint32_t smem_base, const double* __restrict__ A, int64_t lda,
                                              int m0, int M, const double* __restrict__ B, int64_t ldb, int n0,
                                              int N, int kbeg, int kend, double (&acc)[MI][NI][2], int tid,
                                              int wm0, int wn0, int g, int t) {
    const int KT = (kend - kbeg + BK - 1) / BK;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) {
            load_tile(smem_base + s * STAGE_BYTES, A, lda, m0, M, kbeg + s * BK, kend, tid);
            load_tile(smem_base + s * STAGE_BYTES + TILE_BYTES, B, ldb, n0, N, kbeg + s * BK, kend, tid);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nk = kt + STAGES - 1;
        if (nk < KT) {
            const int st = nk % STAGES;
            load_tile(smem_base + st * STAGE_BYTES, A, lda, m0, M, kbeg + nk * BK, kend, tid);
            load_tile(smem_base + st * STAGE_BYTES + TILE_BYTES, B, ldb, n0, N, kbeg + nk * BK, kend, tid);
        }
        cp_async_commit();
        const int st = kt % STAGES;
        compute_stage(smem_base + st * STAGE_BYTES, smem_base + st * STAGE_BYTES + TILE_BYTES, acc, wm0, wn0, g, t);
    }
    cp_async_wait<0>();
    __syncthreads();
}

}  // namespace ipp
