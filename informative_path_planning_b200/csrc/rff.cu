// a11: random-Fourier-feature posterior function samples evaluated on a grid, with a per-sample
// arg-max (general_target aq_library.py:138-139; grid stage of global_maximization :475-479), and
// the feature matrix Z = scale * cos(W X^T + b) (:171).
//   f_s(x_g) = sum_f theta[s][f] cos(W_s[f] . x_g + b_s[f])
// shared_w (S >= RFF_GEMM_MIN_S): values = Phi(grid) Theta^T is a dense FP64 contraction.  Phi[g][f] =
//   cos(W_f . x_g + b_f) is materialised once per chunk of grid points (one cos per (g, f)) and the
//   contraction runs on the DMMA tensor path (dgemm_dmma.cuh) with a fused epilogue: the 128 x 128
//   tile of function values is never stored (unless `values` is asked for); each sample column is
//   reduced to (max, first arg-max) with quad/warp shuffles + a 2-warp shared-memory step.
// shared_w with few samples: one cos per (g, f) reused by up to 32 samples held in registers (SIMT).
// per-sample W (the reference's sampling scheme) has no reuse and is cos-bound.
#include "internal.cuh"
#include "dgemm_dmma.cuh"

namespace ipp {

// cos(x) in ~21 FP64 instructions, branch-free for |x| < 8e5 (phases here are W . x + b with |x| <= a few tens): Cody-Waite
// reduction by pi/2 in three pieces (the first two have 33 significant bits, so their products with k < 2^20 are exact),
// then the fdlibm kernel polynomials on |r| <= pi/4 (< 1 ulp), sine or cosine by the parity of the quadrant with the
// coefficients selected instead of branching.  libdevice's cos() costs ~2.5x as many FP64-pipe slots -- the pipe the
// DMMAs of the grid stage share -- and made the feature pass compute-bound (3 ms for 2^20 x 1000 features against 1.3 ms
// for writing them).
// (The constants sit in constant memory: an FP64 instruction reads a constant-bank operand directly, whereas the compiler
// rebuilds every 64-bit immediate in registers at each use -- two moves per constant, ~56 of the 93 instructions this
// function issued per value, which made the feature pass issue-bound: ncu r5g, issue slots 83 % busy, FP64 pipe 48 %.)
static __constant__ double COSC[18] = {
    0.63661977236758134308, 6755399441055744.0,
    -1.57079632673412561417e+00, -6.07710050630396597660e-11, -2.02226624879595063154e-21,
    1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06, -1.98412698298579493134e-04,
    8.33333333332248946124e-03, -1.66666666666666324348e-01,
    -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07, 2.48015872894767294178e-05,
    -1.38888888888741095749e-03, 4.16666666666666019037e-02, 8.0e5};
__device__ __forceinline__ double cos_core(double x) {   // |x| < 8e5
    const double kf = fma(x, COSC[0], COSC[1]);           // k = rint(x 2 / pi) in the low word
    const int k = __double2loint(kf);
    const double kd = kf - COSC[1];
    double r = fma(kd, COSC[2], x);
    r = fma(kd, COSC[3], r);
    r = fma(kd, COSC[4], r);
    // quadrant 0: cos r, 1: -sin r, 2: -cos r, 3: sin r.  BOTH kernel polynomials are evaluated and the result is selected
    // at the end (selecting every coefficient by the parity costs four integer-pipe instructions per coefficient).
    const double z = r * r;
    double ps = COSC[5], pc = COSC[11];
#pragma unroll
    for (int i = 1; i < 6; ++i) { ps = fma(ps, z, COSC[5 + i]); pc = fma(pc, z, COSC[11 + i]); }
    const double vs = fma(r * z, ps, r);                  // sin r
    const double vc = fma(z * z, pc, fma(-0.5, z, 1.0));  // cos r
    const double v = (k & 1) ? vs : vc;
    return ((k + 1) & 2) ? -v : v;
}
__device__ __forceinline__ double cos_fast(double x) { return (fabs(x) < COSC[17]) ? cos_core(x) : cos(x); }

template <int D, int SC>
__global__ void __launch_bounds__(256) rff_eval_kernel(const double* __restrict__ W, const double* __restrict__ b,
                                                       const double* __restrict__ theta, int F, int S, int shared_w,
                                                       const double* __restrict__ grid, int64_t G,
                                                       double* __restrict__ values, double* __restrict__ part_val,
                                                       int64_t* __restrict__ part_idx) {
    __shared__ double sv[8];
    __shared__ int64_t si[8];
    const int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = gidx < G;
    double x[D];
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] = live ? grid[gidx * D + d] : 0.0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int s0 = 0; s0 < S; s0 += SC) {
        double acc[SC];
#pragma unroll
        for (int j = 0; j < SC; ++j) acc[j] = 0.0;
        if (shared_w) {
            for (int f = 0; f < F; ++f) {
                double ph = b[f];
#pragma unroll
                for (int d = 0; d < D; ++d) ph = fma(W[f * D + d], x[d], ph);
                const double c = cos_fast(ph);
#pragma unroll
                for (int j = 0; j < SC; ++j)
                    if (s0 + j < S) acc[j] = fma(theta[(int64_t)(s0 + j) * F + f], c, acc[j]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < SC; ++j) {
                if (s0 + j < S) {
                    const double* Ws = W + (int64_t)(s0 + j) * F * D;
                    const double* bs = b + (int64_t)(s0 + j) * F;
                    const double* ts = theta + (int64_t)(s0 + j) * F;
                    double a = 0.0;
                    for (int f = 0; f < F; ++f) {
                        double ph = bs[f];
#pragma unroll
                        for (int d = 0; d < D; ++d) ph = fma(Ws[f * D + d], x[d], ph);
                        a = fma(ts[f], cos_fast(ph), a);
                    }
                    acc[j] = a;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < SC; ++j) {
            if (s0 + j >= S) break;
            if (values && live) values[(int64_t)(s0 + j) * G + gidx] = acc[j];
            double v = live ? acc[j] : -INFINITY;
            int64_t ix = live ? gidx : INT64_MAX;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, v, o);
                const int64_t oi = __shfl_xor_sync(0xffffffffu, ix, o);
                if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
            }
            if (lane == 0) { sv[warp] = v; si[warp] = ix; }
            __syncthreads();
            if (threadIdx.x == 0) {
                double bv = sv[0];
                int64_t bi = si[0];
                for (int w = 1; w < 8; ++w)
                    if (sv[w] > bv || (sv[w] == bv && si[w] < bi)) { bv = sv[w]; bi = si[w]; }
                part_val[(int64_t)(s0 + j) * gridDim.x + blockIdx.x] = bv;
                part_idx[(int64_t)(s0 + j) * gridDim.x + blockIdx.x] = bi;
            }
            __syncthreads();
        }
    }
}

// one block per sample: first index of the maximum, like np.argmax
__global__ void __launch_bounds__(256) rff_argmax_finalize_kernel(const double* __restrict__ part_val,
                                                                  const int64_t* __restrict__ part_idx, int nblocks,
                                                                  double* __restrict__ max_val,
                                                                  int64_t* __restrict__ arg_max) {
    __shared__ double sv[256];
    __shared__ int64_t si[256];
    const int s = blockIdx.x;
    double v = -INFINITY;
    int64_t ix = INT64_MAX;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) {
        const double ov = part_val[(int64_t)s * nblocks + i];
        const int64_t oi = part_idx[(int64_t)s * nblocks + i];
        if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
    }
    sv[threadIdx.x] = v;
    si[threadIdx.x] = ix;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double ov = sv[threadIdx.x + o];
            const int64_t oi = si[threadIdx.x + o];
            if (ov > sv[threadIdx.x] || (ov == sv[threadIdx.x] && oi < si[threadIdx.x])) {
                sv[threadIdx.x] = ov;
                si[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { max_val[s] = sv[0]; arg_max[s] = si[0]; }
}

// Z[row][n] = scale cos(W_f . x_n + b_f), row = f, or F - 1 - f with `reverse` (the weight draw factors the REVERSED system)
template <int D>
__global__ void rff_features_kernel(const double* __restrict__ W, const double* __restrict__ b, int64_t F,
                                    const double* __restrict__ X, int64_t N, double scale, double* __restrict__ Z,
                                    int64_t ldz, int reverse) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t f = blockIdx.y;
    if (n >= N) return;
    double ph = b[f];
#pragma unroll
    for (int d = 0; d < D; ++d) ph = fma(W[f * D + d], X[n * D + d], ph);
    Z[(reverse ? F - 1 - f : f) * ldz + n] = scale * cos(ph);
}

// A[i][j] = sum_p part[p][i][j] + (i == j): the split-K partial products of Z Z^T / s2, summed in a fixed order
__global__ void rff_sum_partials_kernel(const double* __restrict__ part, int nparts, int64_t F, int64_t ld, double* __restrict__ A) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= F * ld) return;
    const int64_t i = idx / ld, j = idx % ld;
    double acc = 0.0;
    for (int p = 0; p < nparts; ++p) acc += part[(int64_t)p * F * ld + idx];
    A[idx] = (j < F) ? acc + (i == j ? 1.0 : 0.0) : 0.0;
}

// theta[F - 1 - i] = (Sigma_r rhs_r)[i] / s2 + (L_r^-T eps_r)[i],  eps_r[k] = eps[F - 1 - k];  one block per i
__global__ void __launch_bounds__(128) rff_theta_combine_kernel(const double* __restrict__ Sigma, const double* __restrict__ Linv,
                                                               int64_t ld, int F, const double* __restrict__ rhs,
                                                               const double* __restrict__ eps, double inv_s2,
                                                               double* __restrict__ theta) {
    __shared__ double sm[4], sy[4];
    const int i = blockIdx.x, tid = threadIdx.x;
    double m = 0.0, y = 0.0;
    for (int k = tid; k < F; k += 128) {
        m = fma(Sigma[(int64_t)i * ld + k], rhs[k], m);
        if (k >= i) y = fma(Linv[(int64_t)k * ld + i], eps[F - 1 - k], y);       // L_r^-1 is lower triangular
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m += __shfl_xor_sync(0xffffffffu, m, o);
        y += __shfl_xor_sync(0xffffffffu, y, o);
    }
    if ((tid & 31) == 0) { sm[tid >> 5] = m; sy[tid >> 5] = y; }
    __syncthreads();
    if (tid == 0) theta[F - 1 - i] = ((sm[0] + sm[1]) + (sm[2] + sm[3])) * inv_s2 + ((sy[0] + sy[1]) + (sy[2] + sy[3]));
}

// ---- shared-feature GEMM path -------------------------------------------------------------------
constexpr int RFF_GEMM_MIN_S = 8;
constexpr int64_t RFF_CHUNK = 148 * 128 * 2;      // grid points per Phi chunk: two full waves of 128-row tiles

// Phi[g][f] = cos(W_f . x_g + b_f) for grid points [g0, g0 + Gc), row-major [Gc][ldf]; lanes run along f
template <int D>
__global__ void __launch_bounds__(256) rff_phi_kernel(const double* __restrict__ W, const double* __restrict__ b, int F,
                                                      const double* __restrict__ grid, int64_t g0, int Gc,
                                                      double* __restrict__ Phi, int64_t ldf) {
    constexpr int ROWS = 16;
    __shared__ double sx[ROWS][D];
    const int r0 = blockIdx.x * ROWS;
    if (threadIdx.x < ROWS * D) {
        const int r = threadIdx.x / D, d = threadIdx.x % D;
        sx[r][d] = (r0 + r < Gc) ? grid[(g0 + r0 + r) * D + d] : 0.0;
    }
    __syncthreads();
    // the magnitude test of cos_fast once per (thread, feature) from the block's largest coordinates, the store pointer
    // advanced instead of recomputed, whole blocks without row tests: ~45 instead of ~65 instructions per value
    __shared__ double smax[D];
    if (threadIdx.x < D) {
        double m = 0.0;
        for (int r = 0; r < ROWS; ++r) m = fmax(m, fabs(sx[r][threadIdx.x]));
        smax[threadIdx.x] = m;
    }
    __syncthreads();
    const int rows = (Gc - r0) < ROWS ? (Gc - r0) : ROWS;
    for (int f = threadIdx.x; f < F; f += 256) {
        double w[D];
#pragma unroll
        for (int d = 0; d < D; ++d) w[d] = W[f * D + d];
        const double bf = b[f];
        double bound = fabs(bf);
#pragma unroll
        for (int d = 0; d < D; ++d) bound = fma(fabs(w[d]), smax[d], bound);
        double* out = Phi + (int64_t)r0 * ldf + f;
        if (bound < 7.9e5 && rows == ROWS) {
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                double ph = bf;
#pragma unroll
                for (int d = 0; d < D; ++d) ph = fma(w[d], sx[r][d], ph);
                *out = cos_core(ph);
                out += ldf;
            }
        } else {
            for (int r = 0; r < rows; ++r) {
                double ph = bf;
#pragma unroll
                for (int d = 0; d < D; ++d) ph = fma(w[d], sx[r][d], ph);
                out[(int64_t)r * ldf] = cos_fast(ph);
            }
        }
    }
}

// ---- meshgrid form: grid point g = j nx + i is (xs[i], ys[j]) -- np.meshgrid(xs, ys, indexing='xy') ravelled, the grid
// of the reference's global_maximization (aq_library.py:455-466).  cos(a_i + c_j) = cos a_i cos c_j - sin a_i sin c_j with
// a_i = W_f0 xs_i + b_f and c_j = W_f1 ys_j (+ W_f2 t): (nx + ny) F sincos instead of nx ny F cos.  The general-grid
// feature pass (14 FP64 instructions per cos) takes 3.2 ms for a 2^20 x 1000 grid, while a pure-write kernel reaches
// 7.2 TB/s (profiles/r3g_write_bw.log); two multiplies per feature value leave this pass bound by writing Phi.
template <int D>
__global__ void rff_mesh_tables_kernel(const double* __restrict__ W, const double* __restrict__ b, int F,
                                       const double* __restrict__ xs, int nx, const double* __restrict__ ys, int ny,
                                       double tval, double* __restrict__ TC, double* __restrict__ TS, int64_t ldf) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)(nx + ny) * ldf) return;
    const int r = (int)(idx / ldf), f = (int)(idx % ldf);
    if (f >= F) { TC[idx] = 0.0; TS[idx] = 0.0; return; }
    double ph;
    if (r < nx) ph = fma(W[f * D], xs[r], b[f]);
    else {
        ph = W[f * D + 1] * ys[r - nx];
        if (D == 3) ph = fma(W[f * D + 2], tval, ph);
    }
    double sn, cs;
    sincos(ph, &sn, &cs);
    TC[(int64_t)r * ldf + f] = cs;
    TS[(int64_t)r * ldf + f] = sn;
}

// Phi rows g0 .. g0 + Gc of the mesh.  A block owns a patch of 8 x-coordinates x 8 mesh rows: a thread keeps the eight
// (cos a_i, sin a_i) of its two features in registers and walks the eight rows, so a feature value costs half a table
// load (a block of 16 consecutive grid points re-read both x tables for every value: 2 loads per value, and the pass
// ran at the rate L2 delivers them, 122 us per 37 888-row chunk -- no faster than the cosine version; patches: 88 us;
// 16-byte loads and stores: 64 us = 4.8 TB/s written).
__global__ void __launch_bounds__(256) rff_phi_mesh_kernel(const double* __restrict__ TC, const double* __restrict__ TS, int F,
                                                           int nx, int64_t g0, int Gc, double* __restrict__ Phi, int64_t ldf) {
    const int i0 = blockIdx.x * 8;
    const int64_t j0 = g0 / nx + (int64_t)blockIdx.y * 8;
    for (int f = 2 * threadIdx.x; f < F; f += 512) {                  // two features per thread: 16-byte loads and stores
        double2 ca[8], sa[8];                                         // (ldf is even and the tables are zero beyond F)
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const bool in = i0 + u < nx;
            ca[u] = in ? *reinterpret_cast<const double2*>(TC + (int64_t)(i0 + u) * ldf + f) : make_double2(0.0, 0.0);
            sa[u] = in ? *reinterpret_cast<const double2*>(TS + (int64_t)(i0 + u) * ldf + f) : make_double2(0.0, 0.0);
        }
#pragma unroll 2
        for (int v = 0; v < 8; ++v) {
            const int64_t j = j0 + v;
            const int64_t row0 = j * nx + i0 - g0;                     // row of (i0, j) inside this chunk (may be outside it)
            if (row0 + 8 <= 0 || row0 >= Gc) continue;
            const double2 cb = *reinterpret_cast<const double2*>(TC + (nx + j) * ldf + f);
            const double2 sb = *reinterpret_cast<const double2*>(TS + (nx + j) * ldf + f);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int64_t row = row0 + u;
                if (i0 + u < nx && row >= 0 && row < Gc)
                    *reinterpret_cast<double2*>(Phi + row * ldf + f) =
                        make_double2(fma(ca[u].x, cb.x, -(sa[u].x * sb.x)), fma(ca[u].y, cb.y, -(sa[u].y * sb.y)));
            }
        }
    }
}

__global__ void rff_pad_theta_kernel(const double* __restrict__ theta, int S, int F, double* __restrict__ out, int64_t ldf) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)S * F) out[(i / F) * ldf + i % F] = theta[i];
}

struct RffGemmArgs {
    const double* TC; const double* TS; int nx;  // mesh form (MESH kernels): the two tables; Phi is never materialised
    const double* Phi; int64_t ldf; int Gc;      // A operand: [Gc][ldf]
    const double* Th; int S; int F;              // B operand: [S][ldf]
    int64_t g0, G;                               // global index of the chunk's first grid point, total grid size
    double* values;                              // optional [S][G]
    double* part_val; int64_t* part_idx;         // [S][nparts]
    int nparts, part0;                           // this chunk's m-tiles write columns part0 + blockIdx.x
};

// Phi (128 grid points) x Theta^T (up to 8 NB samples) with the per-sample arg-max in the epilogue.  The eight warps split
// the 128 rows (16 each) and every warp spans ALL sample columns, so the tile is 8 NB wide instead of a multiple of 128:
// S = 100 (config 4) runs as NB = 13 -- 104 columns, 4 % of the DMMAs on padding where the 128-wide tile spent 22 %.
// Pipeline, swizzle and k-permutation are dgemm_dmma.cuh's (3 stages x (A 32 KB + B 32 KB), B rows >= S zero-filled).
//
// MESH: the A operand is formed at fragment-load time from the two tables of the meshgrid form,
//     Phi[(i, j)][f] = C[i][f] C[nx + j][f] - S[i][f] S[nx + j][f],
// so Phi never exists in memory: per k-step a stage holds the tile's 128 rows of C[i] and of S[i] (L2-resident tables;
// tile row r is mesh column (g0 + r) % nx), the <= RFF_MESH_JT mesh rows the tile touches, and the Theta tile (two
// stages of 96 KB + 4 KB).  In the 8 x 1 warp layout no warp shares an A fragment with another, so every value is formed
// exactly once: 8 non-tensor FP64 instructions per k-group and warp beside 52 DMMAs (the pipe they share: + 8 %).
#ifndef RFF_MESH_LATE_KG
#define RFF_MESH_LATE_KG (BK / 16)
#endif
constexpr int RFF_MESH_JT = 4;                   // mesh rows a 128-point tile may touch (nx >= 64)
constexpr int RFF_MESH_TB = RFF_MESH_JT * 2 * BK * 8;                          // bytes: [jrel][cos | sin][BK]
constexpr int RFF_MESH_STAGE = 3 * TILE_BYTES + RFF_MESH_TB;
template <int NB, bool MESH>
__global__ void __launch_bounds__(GEMM_THREADS, 1) rff_gemm_argmax_kernel(RffGemmArgs p) {
    static_assert(GEMM_THREADS == 256 && NB >= 1 && NB <= 16, "eight warps, at most 128 sample columns per tile");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double s_val[8][BN];
    __shared__ int s_row[8][BN];
    const uint32_t smem_base = smem_u32(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = 16 * warp;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    double acc[2][NB][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    if constexpr (MESH) {
        const int KT = (p.F + BK - 1) / BK;
        // loader side: thread (row group r0, chunk column ch) fetches chunk ch of the C and S rows of tile rows r0 + 16 u
        const int ch = tid % CHUNKS_PER_ROW, r0 = tid / CHUNKS_PER_ROW;
        const uint32_t ldst = (uint32_t)(r0 * ROW_BYTES + ((ch ^ ((r0 & 1) << 2)) << 4));
        const int64_t gfirst = p.g0 + m0;
        const int64_t jfirst = gfirst / p.nx;
        // Everything that does not depend on k, once: 32-bit element offsets of the thread's rows in the tables and in
        // Theta (0 for a row outside the grid / beyond the samples, with its validity bit clear), so that a copy costs one
        // 64-bit multiply-add for the address; a stage away from the ragged end of F skips the byte-count logic, and a
        // tile whose 128 rows are all grid points (every tile of a grid that is a multiple of 128) skips the row tests.
        int offC[LOADS_PER_THREAD], offB[LOADS_PER_THREAD];
        unsigned vmask = 0, bmask = 0;
#pragma unroll
        for (int u = 0; u < LOADS_PER_THREAD; ++u) {
            const int r = r0 + RSTEP * u;
            const bool okc = m0 + r < p.Gc, okb = n0 + r < p.S;
            offC[u] = okc ? (int)((gfirst + r) % p.nx) * (int)p.ldf + 2 * ch : 0;
            offB[u] = okb ? (n0 + r) * (int)p.ldf + 2 * ch : 0;
            vmask |= (okc ? 1u : 0u) << u;
            bmask |= (okb ? 1u : 0u) << u;
        }
        const bool all_rows = vmask == (1u << LOADS_PER_THREAD) - 1u;
        // the mesh rows: [jrel][cos | sin][BK], 16 bytes per thread of the first RFF_MESH_JT * BK threads
        const bool jthread = tid < RFF_MESH_JT * 2 * (BK / 2);
        const int c2 = tid % (BK / 2), jcs = (tid / (BK / 2)) & 1, jrl = tid / BK;
        const bool okj = jthread && (jfirst + jrl) * p.nx < p.G;
        const double* jsrc = (jcs ? p.TS : p.TC) + (okj ? (p.nx + jfirst + jrl) * p.ldf + 2 * c2 : 0);
        const uint32_t jdst = (uint32_t)(3 * TILE_BYTES + ((jrl * 2 + jcs) * BK + 2 * c2) * 8);
        auto load_stage = [&](uint32_t st, int k0) {
            const double* tc = p.TC + k0;
            const double* ts = p.TS + k0;
            const double* th = p.Th + k0;
            if (k0 + BK <= p.F) {                         // a whole stage: 16 bytes everywhere
                if (all_rows) {
#pragma unroll
                    for (int u = 0; u < LOADS_PER_THREAD; ++u) {
                        cp_async16(st + ldst + u * (RSTEP * ROW_BYTES), tc + offC[u], 16);
                        cp_async16(st + TILE_BYTES + ldst + u * (RSTEP * ROW_BYTES), ts + offC[u], 16);
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < LOADS_PER_THREAD; ++u) {
                        const int sz = (int)((vmask >> u) & 1u) << 4;
                        cp_async16(st + ldst + u * (RSTEP * ROW_BYTES), tc + offC[u], sz);
                        cp_async16(st + TILE_BYTES + ldst + u * (RSTEP * ROW_BYTES), ts + offC[u], sz);
                    }
                }
#pragma unroll
                for (int u = 0; u < LOADS_PER_THREAD; ++u)
                    if (RSTEP * u < 8 * NB)               // (rows >= 8 NB of the Theta tile are never read)
                        cp_async16(st + 2 * TILE_BYTES + ldst + u * (RSTEP * ROW_BYTES), th + offB[u], (int)((bmask >> u) & 1u) << 4);
                if (jthread) cp_async16(st + jdst, jsrc + k0, okj ? 16 : 0);
            } else {                                      // the ragged last stage of F
                const int rem = p.F - (k0 + 2 * ch);
                const int bytes = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
#pragma unroll
                for (int u = 0; u < LOADS_PER_THREAD; ++u) {
                    const int sz = ((vmask >> u) & 1u) ? bytes : 0;
                    cp_async16(st + ldst + u * (RSTEP * ROW_BYTES), sz ? tc + offC[u] : p.TC, sz);
                    cp_async16(st + TILE_BYTES + ldst + u * (RSTEP * ROW_BYTES), sz ? ts + offC[u] : p.TS, sz);
                    const int szb = ((bmask >> u) & 1u) ? bytes : 0;
                    cp_async16(st + 2 * TILE_BYTES + ldst + u * (RSTEP * ROW_BYTES), szb ? th + offB[u] : p.Th, szb);
                }
                if (jthread) {
                    const int remj = p.F - (k0 + 2 * c2);
                    const int bj = !okj ? 0 : remj >= 2 ? 16 : (remj == 1 ? 8 : 0);
                    cp_async16(st + jdst, bj ? jsrc + k0 : p.TC, bj);
                }
            }
        };
        // consumer side: the two fragment rows of this lane and the mesh row each belongs to
        uint32_t jofs[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int64_t gg = gfirst + wm0 + 8 * i + g;
            const int64_t jr = gg / p.nx - jfirst;
            jofs[i] = (uint32_t)((jr < RFF_MESH_JT ? jr : 0) * 2 * BK * 8);
        }
        load_stage(smem_base, 0);
        cp_async_commit();
        // the two warps of a sub-partition ask for the next stage in turn (as in gemm_mainloop): the ~300 address and
        // predicate instructions of a stage's 24 copies issue no DMMA, and with every warp running them right after the
        // barrier the tensor pipe idled for that long in every k-step (ncu r3l: pipe 75 % busy against 89 % of the
        // plain GEMM form, a third of the issued instructions in the loader)
        const bool late_loader = (warp & 4) != 0;
        for (int kt = 0; kt < KT; ++kt) {
            cp_async_wait<0>();
            __syncthreads();                              // stage kt & 1 landed; everybody is done with stage (kt + 1) & 1
            if (!late_loader) {
                if (kt + 1 < KT) load_stage(smem_base + ((kt + 1) & 1) * RFF_MESH_STAGE, (kt + 1) * BK);
                cp_async_commit();
            }
            const uint32_t sC = smem_base + (kt & 1) * RFF_MESH_STAGE, sS = sC + TILE_BYTES, sB = sC + 2 * TILE_BYTES;
            const uint32_t sT = sC + 3 * TILE_BYTES;
#pragma unroll
            for (int kg = 0; kg < BK / 8; ++kg) {
                if (kg == RFF_MESH_LATE_KG && late_loader) {
                    if (kt + 1 < KT) load_stage(smem_base + ((kt + 1) & 1) * RFF_MESH_STAGE, (kt + 1) * BK);
                    cp_async_commit();
                }
                const uint32_t chunk = (uint32_t)(((4 * kg + t) ^ ((g & 1) << 2)) << 4);
                const uint32_t kofs = (uint32_t)((8 * kg + 2 * t) * 8);
                constexpr int JG = NB > 13 ? NB / 2 : NB;
                double2 a[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const double2 ac = lds128(sC + (wm0 + 8 * i + g) * ROW_BYTES + chunk);
                    const double2 as = lds128(sS + (wm0 + 8 * i + g) * ROW_BYTES + chunk);
                    const double2 cb = lds128(sT + jofs[i] + kofs), sb = lds128(sT + jofs[i] + BK * 8 + kofs);
                    a[i].x = fma(ac.x, cb.x, -(as.x * sb.x));
                    a[i].y = fma(ac.y, cb.y, -(as.y * sb.y));
                }
#pragma unroll
                for (int j0 = 0; j0 < NB; j0 += JG) {
                    double2 b[JG];
#pragma unroll
                    for (int j = 0; j < JG; ++j) b[j] = lds128(sB + (8 * (j0 + j) + g) * ROW_BYTES + chunk);
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < JG; ++j) dmma884(acc[i][j0 + j][0], acc[i][j0 + j][1], a[i].x, b[j].x);
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < JG; ++j) dmma884(acc[i][j0 + j][0], acc[i][j0 + j][1], a[i].y, b[j].y);
                }
            }
        }
        cp_async_wait<0>();
        __syncthreads();
    } else {
        const int KT = (p.F + BK - 1) / BK;
        TileLoader la, lb;
        la.init(p.Phi, p.ldf, m0, p.Gc, tid);
        lb.init(p.Th, p.ldf, n0, p.S, tid);
        const bool late_loader = (warp & 4) != 0;         // see gemm_mainloop: the two warps of a sub-partition prefetch in turn
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            if (s < KT) {
                la.load(smem_base + s * STAGE_BYTES, s * BK, p.F);
                lb.load(smem_base + s * STAGE_BYTES + TILE_BYTES, s * BK, p.F);
            }
            cp_async_commit();
        }
        for (int kt = 0; kt < KT; ++kt) {
            cp_async_wait<STAGES - 2>();
            __syncthreads();
            const int nk = kt + STAGES - 1;
            const uint32_t nst = smem_base + (nk % STAGES) * STAGE_BYTES;
            const uint32_t sA = smem_base + (kt % STAGES) * STAGE_BYTES, sB = sA + TILE_BYTES;
            if (!late_loader) {
                if (nk < KT) { la.load(nst, nk * BK, p.F); lb.load(nst + TILE_BYTES, nk * BK, p.F); }
                cp_async_commit();
            }
#pragma unroll
            for (int kg = 0; kg < BK / 8; ++kg) {
                if (kg == BK / 16 && late_loader) {
                    if (nk < KT) { la.load(nst, nk * BK, p.F); lb.load(nst + TILE_BYTES, nk * BK, p.F); }
                    cp_async_commit();
                }
                const uint32_t chunk = (uint32_t)(((4 * kg + t) ^ ((g & 1) << 2)) << 4);
                constexpr int JG = NB > 13 ? NB / 2 : NB;       // B fragments held at a time (registers: 128 accumulators at NB = 16)
                double2 a[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) a[i] = lds128(sA + (wm0 + 8 * i + g) * ROW_BYTES + chunk);
#pragma unroll
                for (int j0 = 0; j0 < NB; j0 += JG) {
                    double2 b[JG];
#pragma unroll
                    for (int j = 0; j < JG; ++j) b[j] = lds128(sB + (8 * (j0 + j) + g) * ROW_BYTES + chunk);
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < JG; ++j) dmma884(acc[i][j0 + j][0], acc[i][j0 + j][1], a[i].x, b[j].x);
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < JG; ++j) dmma884(acc[i][j0 + j][0], acc[i][j0 + j][1], a[i].y, b[j].y);
                }
            }
        }
        cp_async_wait<0>();
        __syncthreads();
    }

    if (p.values) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int row = m0 + wm0 + 8 * i + g;
            if (row >= p.Gc) continue;
#pragma unroll
            for (int j = 0; j < NB; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = n0 + 8 * j + 2 * t + e;
                    if (col < p.S) p.values[(int64_t)col * p.G + p.g0 + row] = acc[i][j][e];
                }
        }
    }
    // per sample column: (max, first row attaining it) over the tile's 128 grid points
#pragma unroll
    for (int j = 0; j < NB; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double v = -INFINITY;
            int row = 0x7fffffff;
#pragma unroll
            for (int i = 0; i < 2; ++i) {              // rows ascend with i: strict > keeps the first maximum
                const int r = m0 + wm0 + 8 * i + g;
                if (r < p.Gc && acc[i][j][e] > v) { v = acc[i][j][e]; row = r; }
            }
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, v, o);
                const int orow = __shfl_xor_sync(0xffffffffu, row, o);
                if (ov > v || (ov == v && orow < row)) { v = ov; row = orow; }
            }
            if (g == 0) {
                s_val[warp][8 * j + 2 * t + e] = v;
                s_row[warp][8 * j + 2 * t + e] = row;
            }
        }
    __syncthreads();
    if (tid < 8 * NB && n0 + tid < p.S) {
        double v = s_val[0][tid];
        int row = s_row[0][tid];
#pragma unroll
        for (int w = 1; w < 8; ++w)
            if (s_val[w][tid] > v || (s_val[w][tid] == v && s_row[w][tid] < row)) { v = s_val[w][tid]; row = s_row[w][tid]; }
        const int64_t slot = (int64_t)(n0 + tid) * p.nparts + p.part0 + blockIdx.x;
        p.part_val[slot] = v;
        p.part_idx[slot] = (row == 0x7fffffff) ? INT64_MAX : p.g0 + row;
    }
}

template <int NB, bool MESH>
static int rff_gemm_launch(const RffGemmArgs& p, dim3 tiles, cudaStream_t s) {
    constexpr int bytes = MESH ? 2 * RFF_MESH_STAGE : GEMM_SMEM;
    IPP_OPT_IN_SMEM((rff_gemm_argmax_kernel<NB, MESH>), bytes);
    rff_gemm_argmax_kernel<NB, MESH><<<tiles, GEMM_THREADS, bytes, s>>>(p);
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}
template <bool MESH>
static int rff_gemm_dispatch(const RffGemmArgs& p, dim3 tiles, cudaStream_t s) {
    const int nb = p.S > BN ? 16 : (p.S + 7) / 8;                   // tile width: 8 nb sample columns
    return nb <= 4 ? rff_gemm_launch<4, MESH>(p, tiles, s) : nb <= 8 ? rff_gemm_launch<8, MESH>(p, tiles, s)
         : nb <= 13 ? rff_gemm_launch<13, MESH>(p, tiles, s) : rff_gemm_launch<16, MESH>(p, tiles, s);
}
// IPP_RFF_MESH_FUSED=0: the mesh form with a separate feature pass (A/B measurements)
static bool rff_mesh_fused() {
    static const int use = [] { const char* e = getenv("IPP_RFF_MESH_FUSED"); return (e && e[0] == '0') ? 0 : 1; }();
    return use != 0;
}

static int64_t rff_chunk_rows(int64_t G) { return G < RFF_CHUNK ? round_up(G, BM) : RFF_CHUNK; }
static bool rff_use_gemm(int64_t S, int shared_w) { return shared_w && S >= RFF_GEMM_MIN_S; }

}  // namespace ipp

using namespace ipp;

extern "C" int64_t ipp_rff_workspace(int64_t F, int64_t S, int64_t G, int shared_w) {
    if (F <= 0 || S <= 0 || G <= 0) return 16;
    if (!rff_use_gemm(S, shared_w)) return 2 * S * ((G + 255) / 256) + 16;
    const int64_t ldf = round_up(F, 16);
    return rff_chunk_rows(G) * ldf + S * ldf + 2 * S * ((G + BM - 1) / BM) + 16;
}

extern "C" int64_t ipp_rff_mesh_workspace(int64_t F, int64_t S, int64_t nx, int64_t ny) {
    if (F <= 0 || S <= 0 || nx <= 0 || ny <= 0) return 16;
    return ipp_rff_workspace(F, S, nx * ny, 1) + 2 * (nx + ny) * round_up(F, 16);
}

namespace ipp {

// Shared-feature grid stage: Phi chunk by chunk (general grid: one cos per value; mesh: from the two tables), contraction
// with the arg-max in its epilogue, final reduction over the row tiles.
static int rff_gemm_path(const double* W, const double* b, const double* theta, int64_t F, int64_t S, int D,
                         const double* grid, const double* TC, const double* TS, int nx, int64_t G, double* values,
                         double* max_val, int64_t* arg_max, double* work, cudaStream_t s) {
    const int64_t ldf = round_up(F, 16), gc_max = rff_chunk_rows(G);
    const int nparts = (int)((G + BM - 1) / BM);
    double* Phi = work;
    double* Th = Phi + gc_max * ldf;
    double* part_val = Th + S * ldf;
    int64_t* part_idx = reinterpret_cast<int64_t*>(part_val + (int64_t)S * nparts);
    rff_pad_theta_kernel<<<(unsigned)((S * F + 255) / 256), 256, 0, s>>>(theta, (int)S, (int)F, Th, ldf);
    IPP_LAUNCH_CHECK();
    if (TC && rff_mesh_fused() && (BM + nx - 2) / nx + 1 <= RFF_MESH_JT && G < (1LL << 31) &&
        (int64_t)nx * ldf < (1LL << 31) && S * ldf < (1LL << 31)) {     // (the loader keeps 32-bit row offsets)
        RffGemmArgs p;                              // one launch over the whole grid: Phi formed in the A fragments
        p.TC = TC; p.TS = TS; p.nx = nx;
        p.Phi = nullptr; p.ldf = ldf; p.Gc = (int)G;
        p.Th = Th; p.S = (int)S; p.F = (int)F;
        p.g0 = 0; p.G = G; p.values = values;
        p.part_val = part_val; p.part_idx = part_idx;
        p.nparts = nparts; p.part0 = 0;
        dim3 tiles((unsigned)((G + BM - 1) / BM), (unsigned)((S + BN - 1) / BN));
        IPP_TRY(rff_gemm_dispatch<true>(p, tiles, s));
        rff_argmax_finalize_kernel<<<(unsigned)S, 256, 0, s>>>(part_val, part_idx, nparts, max_val, arg_max);
        IPP_LAUNCH_CHECK();
        return IPP_OK;
    }
    // (Tried: the feature pass of chunk c + 1 on a helper stream beside the contraction of chunk c, two Phi buffers.  No
    // overlap at all -- the contraction's CTA holds 238 x 256 of the SM's 65 536 registers, so no second block fits.)
    for (int64_t g0 = 0; g0 < G; g0 += gc_max) {
        const int gc = (int)((G - g0) < gc_max ? (G - g0) : gc_max);
        if (TC) {
            const int64_t nj = (g0 + gc - 1) / nx - g0 / nx + 1;           // mesh rows this chunk touches
            rff_phi_mesh_kernel<<<dim3((unsigned)((nx + 7) / 8), (unsigned)((nj + 7) / 8)), 256, 0, s>>>(TC, TS, (int)F, nx, g0, gc, Phi, ldf);
        }
        else if (D == 2) rff_phi_kernel<2><<<(gc + 15) / 16, 256, 0, s>>>(W, b, (int)F, grid, g0, gc, Phi, ldf);
        else rff_phi_kernel<3><<<(gc + 15) / 16, 256, 0, s>>>(W, b, (int)F, grid, g0, gc, Phi, ldf);
        IPP_LAUNCH_CHECK();
        RffGemmArgs p;
        p.TC = nullptr; p.TS = nullptr; p.nx = 0;
        p.Phi = Phi; p.ldf = ldf; p.Gc = gc;
        p.Th = Th; p.S = (int)S; p.F = (int)F;
        p.g0 = g0; p.G = G; p.values = values;
        p.part_val = part_val; p.part_idx = part_idx;
        p.nparts = nparts; p.part0 = (int)(g0 / BM);
        dim3 tiles((gc + BM - 1) / BM, (unsigned)((S + BN - 1) / BN));
        IPP_TRY(rff_gemm_dispatch<false>(p, tiles, s));
    }
    rff_argmax_finalize_kernel<<<(unsigned)S, 256, 0, s>>>(part_val, part_idx, nparts, max_val, arg_max);
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

}  // namespace ipp

extern "C" int ipp_rff_eval_argmax_mesh(const double* W, const double* b, const double* theta, int64_t F, int64_t S, int D,
                                        const double* xs, int64_t nx, const double* ys, int64_t ny, double tval,
                                        double* values, double* max_val, int64_t* arg_max, double* work, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    IPP_CHECK_ARG(W && b && theta && xs && ys && max_val && arg_max && work, "null pointer");
    IPP_CHECK_ARG(F >= 1 && F < (1 << 30) && S >= RFF_GEMM_MIN_S && S < (1 << 30), "shared features, at least 8 samples");
    IPP_CHECK_ARG(nx >= 1 && ny >= 1 && nx < (1 << 30) && ny < (1 << 30), "bad mesh");
    IPP_CHECK_ARG(D == 2 || D == 3, "dimension must be 2 or 3");
    const int64_t ldf = round_up(F, 16), G = nx * ny;
    double* TC = work + ipp_rff_workspace(F, S, G, 1);
    double* TS = TC + (nx + ny) * ldf;
    const unsigned nb = (unsigned)(((nx + ny) * ldf + 255) / 256);
    if (D == 2) rff_mesh_tables_kernel<2><<<nb, 256, 0, s>>>(W, b, (int)F, xs, (int)nx, ys, (int)ny, tval, TC, TS, ldf);
    else rff_mesh_tables_kernel<3><<<nb, 256, 0, s>>>(W, b, (int)F, xs, (int)nx, ys, (int)ny, tval, TC, TS, ldf);
    IPP_LAUNCH_CHECK();
    return rff_gemm_path(W, b, theta, F, S, D, nullptr, TC, TS, (int)nx, G, values, max_val, arg_max, work, s);
}

extern "C" int ipp_rff_eval_argmax(const double* W, const double* b, const double* theta, int64_t F, int64_t S, int D,
                                   int shared_w, const double* grid, int64_t G, double* values,
                                   double* max_val, int64_t* arg_max, double* work, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    IPP_CHECK_ARG(W && b && theta && grid && max_val && arg_max && work, "null pointer");
    IPP_CHECK_ARG(F >= 1 && S >= 1 && G >= 1 && F < (1 << 30) && S < (1 << 30), "bad shape");
    IPP_CHECK_ARG(D == 2 || D == 3, "dimension must be 2 or 3");
    if (rff_use_gemm(S, shared_w))
        return rff_gemm_path(W, b, theta, F, S, D, grid, nullptr, nullptr, 0, G, values, max_val, arg_max, work, s);
    const int nblocks = (int)((G + 255) / 256);
    double* part_val = work;
    int64_t* part_idx = reinterpret_cast<int64_t*>(work + (int64_t)S * nblocks);
    if (shared_w) {
        if (D == 2) rff_eval_kernel<2, 32><<<nblocks, 256, 0, s>>>(W, b, theta, (int)F, (int)S, 1, grid, G, values, part_val, part_idx);
        else rff_eval_kernel<3, 32><<<nblocks, 256, 0, s>>>(W, b, theta, (int)F, (int)S, 1, grid, G, values, part_val, part_idx);
    } else {
        if (D == 2) rff_eval_kernel<2, 4><<<nblocks, 256, 0, s>>>(W, b, theta, (int)F, (int)S, 0, grid, G, values, part_val, part_idx);
        else rff_eval_kernel<3, 4><<<nblocks, 256, 0, s>>>(W, b, theta, (int)F, (int)S, 0, grid, G, values, part_val, part_idx);
    }
    IPP_LAUNCH_CHECK();
    rff_argmax_finalize_kernel<<<(unsigned)S, 256, 0, s>>>(part_val, part_idx, nblocks, max_val, arg_max);
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

extern "C" int ipp_rff_features(const double* W, const double* b, int64_t F, int D, const double* X, int64_t N,
                                double scale, double* Z, int64_t ldz, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    IPP_CHECK_ARG(W && b && X && Z && F >= 1 && N >= 1 && ldz >= N && F <= 65535, "bad arguments");
    IPP_CHECK_ARG(D == 2 || D == 3, "dimension must be 2 or 3");
    dim3 grid((unsigned)((N + 255) / 256), (unsigned)F);
    if (D == 2) rff_features_kernel<2><<<grid, 256, 0, s>>>(W, b, F, X, N, scale, Z, ldz, 0);
    else rff_features_kernel<3><<<grid, 256, 0, s>>>(W, b, F, X, N, scale, Z, ldz, 0);
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

// ---- posterior weight draw (N >= F branch of the reference's sampler) in one call
constexpr int64_t RFF_THETA_KSPLIT = 512;         // observations per partial product of Z Z^T: F x F x N has few output tiles
static int64_t rff_theta_nparts(int64_t N) { return (N + RFF_THETA_KSPLIT - 1) / RFF_THETA_KSPLIT; }

// eps_r[s][k] = eps[s][F - 1 - k] (zero beyond F): the draws in the order of the reversed system
__global__ void rff_reverse_rows_kernel(const double* __restrict__ eps, int64_t S, int64_t F, int64_t ld, double* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S * ld) return;
    const int64_t sidx = idx / ld, k = idx % ld;
    out[idx] = k < F ? eps[sidx * F + (F - 1 - k)] : 0.0;
}

// theta[s][F - 1 - i] = m[i] / s2 + Y[s][i]
__global__ void rff_theta_assemble_kernel(const double* __restrict__ m, const double* __restrict__ Y, int64_t S, int64_t F,
                                          int64_t ld, double inv_s2, double* __restrict__ theta) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S * F) return;
    const int64_t sidx = idx / F, i = idx % F;
    theta[sidx * F + (F - 1 - i)] = fma(m[i], inv_s2, Y[sidx * ld + i]);
}

extern "C" int64_t ipp_rff_theta_workspace(int64_t F, int64_t N, int64_t S) {
    if (F <= 0 || N <= 0 || S <= 0) return 16;
    const int64_t ld = round_up(F, 16), ldz = rff_theta_nparts(N) * RFF_THETA_KSPLIT;
    return F * ldz + rff_theta_nparts(N) * F * ld + 3 * F * ld + 2 * round_up(F, 2) + 2 * S * ld + ipp_pdinv_workspace(F) + 64;
}

extern "C" int ipp_rff_theta_draw(const double* W, const double* b, int64_t F, int D, const double* X, const double* z,
                                  int64_t N, double variance, double noise_var, const double* eps, int64_t S, double* theta,
                                  int32_t* info, double* work, void* stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    IPP_CHECK_ARG(W && b && X && z && eps && theta && info && work, "null pointer");
    IPP_CHECK_ARG(F >= 1 && F <= 65535 && N >= 1 && S >= 1 && S < (1 << 30) && (D == 2 || D == 3) && noise_var > 0.0 &&
                  variance > 0.0, "bad arguments");
    const int64_t ld = round_up(F, 16), nparts = rff_theta_nparts(N), ldz = nparts * RFF_THETA_KSPLIT;
    double* Zr = work;                          // [F][ldz] reversed features, zero beyond N
    double* part = Zr + F * ldz;                // [nparts][F][ld]
    double* A = part + nparts * F * ld;         // [F][ld]   Z_r Z_r^T / s2 + I, then its Cholesky factor
    double* Sigma = A + F * ld;                 // [F][ld]   A_r^-1
    double* Linv = Sigma + F * ld;              // [F][ld]   L_r^-1
    double* rhs = Linv + F * ld;                // [F]       Z_r z
    double* mvec = rhs + round_up(F, 2);        // [F]       A_r^-1 Z_r z        (S > 1)
    double* EpsR = mvec + round_up(F, 2);       // [S][ld]   the draws, reversed  (S > 1)
    double* Y = EpsR + S * ld;                  // [S][ld]   L_r^-T eps_r         (S > 1)
    double* pwork = Y + S * ld;
    IPP_CUDA(cudaMemsetAsync(Zr, 0, sizeof(double) * F * ldz, s));
    const double scale = sqrt(2.0 * variance / (double)F);
    dim3 grid((unsigned)((N + 255) / 256), (unsigned)F);
    if (D == 2) rff_features_kernel<2><<<grid, 256, 0, s>>>(W, b, F, X, N, scale, Zr, ldz, 1);
    else rff_features_kernel<3><<<grid, 256, 0, s>>>(W, b, F, X, N, scale, Zr, ldz, 1);
    IPP_LAUNCH_CHECK();
    GemmArgs g{};                               // part[p] = Z_r[:, p-th slice] Z_r[:, p-th slice]^T / s2
    g.A = Zr; g.lda = ldz; g.B = Zr; g.ldb = ldz; g.C = part; g.ldc = ld;
    g.M = (int)F; g.N = (int)F; g.K = (int)RFF_THETA_KSPLIT; g.alpha = 1.0 / noise_var; g.beta = 0.0;
    g.batch = (int)nparts; g.sA = RFF_THETA_KSPLIT; g.sB = RFF_THETA_KSPLIT; g.sC = F * ld;
    IPP_TRY(launch_gemm(g, s));
    rff_sum_partials_kernel<<<(unsigned)((F * ld + 255) / 256), 256, 0, s>>>(part, (int)nparts, F, ld, A);
    IPP_LAUNCH_CHECK();
    IPP_TRY(launch_gemv(Zr, ldz, F, N, z, rhs, s));
    IPP_TRY(pdinv_device(A, F, ld, Sigma, Linv, ld, nullptr, info, pwork, s));
    if (S == 1) {
        rff_theta_combine_kernel<<<(unsigned)F, 128, 0, s>>>(Sigma, Linv, ld, (int)F, rhs, eps, 1.0 / noise_var, theta);
        IPP_LAUNCH_CHECK();
        return IPP_OK;
    }
    // S draws share the factorisation: Y = Eps_r U^T with U = L_r^-T, which pdinv_device leaves at the head of its workspace
    const double* U = pwork;
    IPP_TRY(launch_gemv(Sigma, ld, F, F, rhs, mvec, s));
    rff_reverse_rows_kernel<<<(unsigned)((S * ld + 255) / 256), 256, 0, s>>>(eps, S, F, ld, EpsR);
    IPP_LAUNCH_CHECK();
    GemmArgs y{};
    y.A = EpsR; y.lda = ld; y.B = U; y.ldb = ld; y.C = Y; y.ldc = ld;
    y.M = (int)S; y.N = (int)F; y.K = (int)F; y.alpha = 1.0; y.beta = 0.0;
    IPP_TRY(launch_gemm(y, s));
    rff_theta_assemble_kernel<<<(unsigned)((S * F + 255) / 256), 256, 0, s>>>(mvec, Y, S, F, ld, 1.0 / noise_var, theta);
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}
