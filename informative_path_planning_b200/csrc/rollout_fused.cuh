// The whole rollout in ONE launch (included by rollout.cu after RolloutArgs, td_load8 and the reward closed forms).
//
// The four-kernel sequence costs ~35 us per rollout at N = 900, nearly all of it launch and hand-off latency
// (profiles/r1j_launches_mcts_native_900_final.csv), and its W^-1 pass reads the full symmetric matrix.  Here
//   * only the lower 64 x 64 tiles of W^-1 are read.  With W = Ls + Dg + Ls^T (strictly-lower tiles, diagonal tiles) and
//     K = the cross-covariance rows of the rollout's points,
//         K W K^T = A + A^T,     A = sum_{J} ( sum_{I > J} K_I W_IJ + 1/2 K_J W_JJ ) K_J^T = sum_J U_J K_J^T;
//   * the tiles of the lower triangle, in column-major order, are cut into contiguous runs, one per CTA.  While a CTA
//     stays in tile column J its 8 warps keep U_J (P x 64, 8 columns per warp) in DMMA accumulators:
//         U_J[a][j] += sum_{i in I} K[a][i] W[i][j]       A operand = K_I from a fragment-ordered shared-memory image the
//                                                          CTA computes itself (1/2 folded in on the diagonal tile),
//                                                          B operand = W^-1 straight from global memory, one tile ahead
//     and when the column ends the accumulator fragments ARE the A operand of the second contraction (k permutation:
//     thread t of a quad owns columns 2t, 2t+1 in both), so  A += U_J K_J^T  needs no shared-memory round trip;
//   * on the diagonal tile the CTA also adds K_J alpha_J to its partial mean;
//   * every CTA stores its partial (A + A^T lower triangle and the mean: M (M + 1) / 2 + M doubles), the last one to finish
//     (a ticket counter) sums the partials in CTA order -- a fixed order, so the result does not depend on timing --,
//     forms C0 = Kqq - (A + A^T) and m0, and runs the level-by-level conditioning in the same launch;
//   * the rewards are not evaluated inside the sequential conditioning: each level's predictive moments are snapshot when
//     the level is reached and all levels are scored in parallel (one warp per level) at the end.
#pragma once
#include <cooperative_groups.h>

namespace ipp {

namespace cg = cooperative_groups;

constexpr int RF_B = 64, RF_THREADS = 256, RF_WARPS = 8;
constexpr int RF_STAGES = 4;                            // streaming form: ring of tile stages per CTA
constexpr int RF_AHEAD = 2;                             // ... of which this many are requested ahead of the tile being contracted
constexpr int RF_TILE = RF_B * RF_B;                    // a staged tile of W^-1: four 16-column panels [64 rows][128 B], TMA 128-byte swizzle
constexpr int RF_CLUSTER = 4;                           // CTAs per thread-block cluster: their partials meet in shared memory
constexpr int RF_SLOTS = 256;                           // ticket counters; each launch takes the next one
__device__ unsigned int g_rf_ticket[RF_SLOTS];          // zero at load; the last CTA of a launch puts its slot back to zero
constexpr int RF_GROUP = 8, RF_MAX_GROUPS = 32;         // streaming form: partials are first summed in groups of RF_GROUP CTAs
__device__ unsigned int g_rf_ticket2[RF_SLOTS][RF_MAX_GROUPS];
__device__ long long g_rf_clock[32];                     // SM clocks of the finishing CTA per phase (ipp_rollout_debug_clocks)

// The rollout's points and observations by value: [M][D] points, then [M] observations (M <= 32 in this kernel: 1 KB of
// kernel parameters instead of the 3 KB RolloutInline of the four-kernel form).
struct RolloutInlineSmall { double v[32 * (IPP_MAX_DIM + 1)]; };

// A result the host can validate without a fence: value and tag travel in ONE 16-byte store, so a reader that sees the
// tag sees the value (the four-kernel form orders plain stores before a completion word with __threadfence_system, a
// PCIe round trip of ~3.5k SM clocks at the end of every rollout).
__device__ __forceinline__ void rf_publish(double* rec, int slot, double value, unsigned long long tag) {
    *reinterpret_cast<double2*>(rec + 2 * slot) = make_double2(value, __longlong_as_double((long long)tag));
}

struct RolloutFusedArgs {
    RbfDev kern;
    const double* X; const double* alpha; const double* W; int64_t ldw; int N;
    const double* pts; const double* zobs;               // device copies (used when use_inline == 0)
    int use_inline;
    double* part;                                        // [clusters][V], V = M (M + 1) / 2 + M
    double* part2;                                       // [RF_MAX_GROUPS][V]: group sums (streaming form)
    const double* img;                                   // streaming form: [T][4][P * 64] cross-covariance images (pre-pass)
    int slot, tiles, T;
    unsigned long long tag;                              // != 0: r.reward holds 16-byte {value, tag} records, L rewards then info
    RolloutArgs r;
};

// exp(x), x <= 0, branch-free (several calls interleave in one thread): Cody-Waite reduction by ln 2 and the degree-13
// Taylor polynomial on |f| <= ln 2 / 2 (truncation 4e-18 relative), result scaled by 2^n through the exponent bits.
__device__ __forceinline__ double rf_exp_neg(double x) {
    const double xc = x < -700.0 ? -700.0 : x;
    const double tt = fma(xc, 1.4426950408889634074, 6755399441055744.0);     // n = rint(x log2 e) in the low word
    const int n = __double2loint(tt);
    const double nf = tt - 6755399441055744.0;
    double f = fma(nf, -6.93147180369123816490e-01, xc);
    f = fma(nf, -1.90821492927058770002e-10, f);
    double p = 1.6059043836821613e-10;                                        // 1/13!
    p = fma(p, f, 2.08767569878681e-09);
    p = fma(p, f, 2.505210838544172e-08);
    p = fma(p, f, 2.755731922398589e-07);
    p = fma(p, f, 2.7557319223985893e-06);
    p = fma(p, f, 2.48015873015873e-05);
    p = fma(p, f, 1.984126984126984e-04);
    p = fma(p, f, 1.388888888888889e-03);
    p = fma(p, f, 8.333333333333333e-03);
    p = fma(p, f, 4.1666666666666664e-02);
    p = fma(p, f, 1.6666666666666666e-01);
    p = fma(p, f, 0.5);
    p = fma(p, f, 1.0);
    p = fma(p, f, 1.0);
    const double r = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));   // p in [0.70, 1.42]
    return x < -700.0 ? 0.0 : r;
}

template <int D>
__device__ __forceinline__ double rf_rbf(const RbfDev& k, const double* a, const double* b) {
    double r2 = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const double u = (a[d] - b[d]) * k.il[d];
        r2 = fma(u, u, r2);
    }
    return k.variance * rf_exp_neg(-0.5 * r2);
}

template <int NT, bool STREAM>
constexpr int rf_smem_doubles() {
    constexpr int P = 8 * NT;
    if (STREAM)     // points, K_J image, mean, obs, partial (V rounded up: 16-byte alignment), then the ring of RF_STAGES x
        return P * 3 + P * RF_B + P + P + ((P * (P + 1) / 2 + P + 1) & ~1) +      // (tile of W^-1 + A image) fed by bulk copies
               RF_STAGES * (RF_TILE + P * RF_B) + 128 + 8;                        // (+ 1 KB: the ring starts on a 1024-byte boundary)
    //      points  K_I images (2)    K_J image + plain copy   mean  obs   this CTA's partial + the (RF_CLUSTER - 1) its cluster mates push
    return P * 3 + 2 * P * RF_B + 2 * P * RF_B + P + P + RF_CLUSTER * ((P * (P + 1) / 2 + P + 1) & ~1) + 8;
}

// Thread index as the code below the tile loop sees it.  The streaming form runs a ninth warp that only feeds the ring;
// its threads get indices no strided loop or guard ever reaches (tid >= 2^20, warp >= 2^15), so outside the tile loop they
// just keep the block-wide barriers company.
__device__ __forceinline__ int rf_tid() {
    return threadIdx.x < RF_THREADS ? (int)threadIdx.x : (1 << 20) + (int)(threadIdx.x & 31);
}

// All levels scored in parallel from the snapshots: warp w takes levels w, w + 8, ...  Called by every thread after a
// barrier that made the snapshots visible.
__device__ __forceinline__ void rf_score(const RolloutArgs& p, const double* snap_m, const double* snap_v, double* Cs,
                                         int& s_bad, unsigned long long tag) {
    const int M = p.M, ldc = M + 1;
    const int tid = rf_tid(), lane = tid & 31, warp = tid >> 5;
    if (tid == 0) g_rf_clock[25] = clock64();
    for (int l = warp; l < p.L; l += RF_WARPS) {
        const int b0 = p.offsets[l], kb = p.offsets[l + 1] - b0;
        double s0 = 0.0, s1 = 0.0;
        if (p.kind == IPP_REWARD_INFO) {                  // in-place Cholesky of the kb x kb block, lane = row
            double* a = Cs + b0 * ldc + b0;
            int bad = 0;
            for (int j = 0; j < kb; ++j) {
                if (lane == 0) {
                    double d = a[j * ldc + j];
                    if (!(d > 0.0)) { bad = 1; d = 1.0; }
                    a[j * ldc + j] = sqrt(d);
                }
                __syncwarp();
                if (lane > j && lane < kb) a[lane * ldc + j] /= a[j * ldc + j];
                __syncwarp();
                if (lane > j && lane < kb)
                    for (int c = j + 1; c <= lane; ++c) a[lane * ldc + c] = fma(-a[lane * ldc + j], a[c * ldc + j], a[lane * ldc + c]);
                __syncwarp();
            }
            if (lane == 0) {
                double sacc = 0.0;
                for (int j = 0; j < kb; ++j) sacc += log(a[j * ldc + j]);
                if (tag) rf_publish(p.reward, l, p.p0 * 2.0 * sacc, tag);
                else p.reward[l] = p.p0 * 2.0 * sacc;
                if (bad) s_bad = l + 1;
            }
            continue;
        }
        if (p.kind == IPP_REWARD_MES) {
            for (int idx = lane; idx < kb * p.S; idx += 32) {
                const int qi = b0 + idx / p.S;
                s0 += ro_mes_utility(p.maxes[idx % p.S], snap_m[qi], snap_v[qi]);
            }
        } else {
            for (int qi = b0 + lane; qi < b0 + kb; qi += 32) {
                s0 += (p.kind == IPP_REWARD_EI) ? (snap_m[qi] - p.p0) : snap_m[qi];
                s1 += fabs(snap_v[qi]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        if (lane == 0) {
            double r;
            if (p.kind == IPP_REWARD_UCB) {
                r = s0 + p.p0 * s1;
            } else if (p.kind == IPP_REWARD_EI) {
                const double z = s0 / s1;
                const double big_phi = 0.5 * (1.0 + erf(z / 1.41421356237309504880));
                const double small_phi = 1.0 / 2.50662827463100050242 * exp(-(z * z) / 2.0);
                r = s0 * big_phi + s1 * small_phi;
            } else {
                r = 2.0 * s0 / (double)p.S;
            }
            if (tag) rf_publish(p.reward, l, r, tag);
            else p.reward[l] = r;
        }
    }
    __syncthreads();
    if (tid == 0) {
        g_rf_clock[26] = clock64();
        if (tag) {
            rf_publish(p.reward, p.L, (double)s_bad, tag);
        } else {
            *p.info = s_bad;
            if (p.done) {
                __threadfence_system();
                *reinterpret_cast<volatile unsigned int*>(p.done) = p.done_seq;
            }
        }
    }
}

// Conditioning with the rewards deferred: same arithmetic per level as ro_condition (rollout.cu), but a level only
// snapshots (mean, variance + noise) of its points; all levels are scored afterwards, one warp per level.
// kind == IPP_REWARD_INFO: the operand is the information-gain belief ((I + vK)^-1, kernel variance v^2, diag_add = 1,
// see ipp_b200.h); a level's reward is p0 * logdet(I + C[B, B]) of its block under the conditioned covariance
// (aq_library.py:55-70 in Schur form), so the level snapshots that block (into Cs, same positions) instead of moments.
__device__ __forceinline__ void rf_condition(const RolloutArgs& p, const double* zobs, double* C, double* m, double* H,
                                             double* snap_m, double* snap_v, double* Cs, int& s_bad,
                                             unsigned long long tag) {
    const int M = p.M, ldc = M + 1;
    const int tid = rf_tid(), lane = tid & 31, warp = tid >> 5;
    constexpr int HL = RO_MAX_BLOCK + 1;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    if (tid == 0) g_rf_clock[16] = clock64();
    for (int l = 0; l < p.L; ++l) {
        const int b0 = p.offsets[l], kb = p.offsets[l + 1] - b0;
        if (tid == 0 && l < 8) g_rf_clock[17 + l] = clock64();
        if (p.kind == IPP_REWARD_INFO) {
            for (int idx = tid; idx < kb * kb; idx += RF_THREADS) {
                const int a = b0 + idx / kb, b = b0 + idx % kb;
                Cs[a * ldc + b] = C[a * ldc + b] + (a == b ? 1.0 : 0.0);
            }
        } else if (tid < kb) {
            snap_m[b0 + tid] = m[b0 + tid];
            snap_v[b0 + tid] = C[(b0 + tid) * ldc + b0 + tid] + p.noise_pred;
        }
        if (l == p.L - 1) break;                  // the last level's observation is never looked at again
        if (p.reps[l] > 0 && kb <= 4) {
            // every thread of the FIRST warp factors the (identity-padded) 4 x 4 block redundantly in registers and
            // substitutes its own row (M <= 32: one warp has a lane for every point).  The other warps skip all of it:
            // the FP64 pipe of a sub-partition is shared by its two warps, so the seven warps that computed
            // the factorisation for nothing were not free -- warp 4 shares warp 0's pipe.
            // Only the points AFTER this block are ever looked at again: rows / columns >= b1.  (No barrier here: the
            // snapshot above only reads what the barrier that ended the previous level made final.)
            const double dadd = p.diag_add / (double)p.reps[l];
            const int b1 = b0 + kb, R = M - b1;
            const int i = b1 + tid;               // this thread's point (M <= 32 < RF_THREADS: at most one)
            double dm = 0.0;
            if (warp == 0) {
            double Lf[4][4], rd[4], u[4];
            int bad = 0;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b <= a; ++b)
                    Lf[a][b] = (a < kb) ? C[(b0 + a) * ldc + b0 + b] + (a == b ? dadd : 0.0) : (a == b ? 1.0 : 0.0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool ok = Lf[j][j] > 0.0;       // branch-free: the substitutions below overlap the pivot chain
                bad = (!ok && bad == 0) ? (l + 1) : bad;
                const double r = rsqrt_fast(ok ? Lf[j][j] : 1.0);
                rd[j] = r;
#pragma unroll
                for (int i = j + 1; i < 4; ++i) Lf[i][j] *= r;
#pragma unroll
                for (int i = j + 1; i < 4; ++i)
#pragma unroll
                    for (int c = j + 1; c <= i; ++c) Lf[i][c] = fma(-Lf[i][j], Lf[c][j], Lf[i][c]);
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                double v = (a < kb) ? zobs[b0 + a] - m[b0 + a] : 0.0;
#pragma unroll
                for (int c = 0; c < a; ++c) v = fma(-Lf[a][c], u[c], v);
                u[a] = v * rd[a];
            }
            if (bad && tid == 0) s_bad = bad;
            if (i < M) {
                double h[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    double v = (a < kb) ? C[i * ldc + b0 + a] : 0.0;
#pragma unroll
                    for (int c = 0; c < a; ++c) v = fma(-Lf[a][c], h[c], v);
                    h[a] = v * rd[a];
                    H[i * HL + a] = h[a];
                    dm = fma(h[a], u[a], dm);
                }
            }
            }
            __syncthreads();
            if (i < M) m[i] += dm;
            // C -= H H^T on the rows / columns that are still ahead.  R <= 16 (the usual case): 16 threads per row, all rows
            // in one pass; else a warp per row
            if (R <= 16) {
                const int ii = b1 + (tid >> 4), jj = b1 + (tid & 15);
                if (ii < M && jj < M) {
                    double sacc = 0.0;
#pragma unroll
                    for (int a = 0; a < 4; ++a) sacc = fma(H[ii * HL + a], H[jj * HL + a], sacc);
                    C[ii * ldc + jj] -= sacc;
                }
            } else {
                for (int ii = b1 + warp; ii < M; ii += RF_WARPS) {
                    const int jj = b1 + lane;
                    if (lane < R) {
                        double sacc = 0.0;
#pragma unroll
                        for (int a = 0; a < 4; ++a) sacc = fma(H[ii * HL + a], H[jj * HL + a], sacc);
                        C[ii * ldc + jj] -= sacc;
                    }
                }
            }
        } else if (p.reps[l] > 0) {
            const double dadd = p.diag_add / (double)p.reps[l];
            const int R = M - b0;
            __syncthreads();
            for (int idx = tid; idx < R * kb; idx += RF_THREADS) {
                const int i = idx / kb, a = idx % kb;
                H[(b0 + i) * HL + a] = C[(b0 + i) * ldc + b0 + a] + (i == a ? dadd : 0.0);
            }
            if (tid < kb) H[M * HL + tid] = zobs[b0 + tid] - m[b0 + tid];
            __syncthreads();
            for (int a = 0; a < kb; ++a) {
                const double piv = H[(b0 + a) * HL + a];
                __syncthreads();
                if (!(piv > 0.0) && tid == 0) s_bad = l + 1;
                const double d = piv > 0.0 ? rsqrt_fast(piv) : 1.0;
                for (int i = b0 + tid; i <= M; i += RF_THREADS)
                    H[i * HL + a] = (i < b0 + a) ? 0.0 : H[i * HL + a] * d;
                __syncthreads();
                const int rest = kb - a - 1;
                for (int idx = tid; idx < (R + 1) * rest; idx += RF_THREADS) {
                    const int i = b0 + idx / rest, c = a + 1 + idx % rest;
                    H[i * HL + c] = fma(-H[i * HL + a], H[(b0 + c) * HL + a], H[i * HL + c]);
                }
                __syncthreads();
            }
            for (int i = b0 + tid; i < M; i += RF_THREADS) {
                double dm = 0.0;
                for (int a = 0; a < kb; ++a) dm = fma(H[i * HL + a], H[M * HL + a], dm);
                m[i] += dm;
            }
            for (int ii = b0 + warp; ii < M; ii += RF_WARPS) {
                const int jj = b0 + lane;
                if (lane < R) {
                    double s = 0.0;
                    for (int a = 0; a < kb; ++a) s = fma(H[ii * HL + a], H[jj * HL + a], s);
                    C[ii * ldc + jj] -= s;
                }
            }
        }
        __syncthreads();
    }
    __syncthreads();
    rf_score(p, snap_m, snap_v, Cs, s_bad, tag);
}

// The cross-covariance images of the streaming form, one CTA per 64-block of observations (four slots of P x 64 doubles):
// slot 0 the image in fragment order (A operand of the first contraction, B operand of the fold), slot 1 the same scaled
// by 1/2 (diagonal tiles), and -- in the first P doubles of slot 3 -- the block's share of the prior mean,
// K[a][block] . alpha[block] (fixed summation order).
template <int D, int NT>
__global__ void __launch_bounds__(RF_THREADS) rollout_images_kernel(RbfDev kern, const double* __restrict__ X, int N,
                                                                    const double* __restrict__ alpha,
                                                                    const double* __restrict__ pts, int use_inline, int M,
                                                                    const __grid_constant__ RolloutInlineSmall pp,
                                                                    double* __restrict__ img) {
    constexpr int P = 8 * NT;
    __shared__ double s_pts[P * 3];
    __shared__ double s_red[RF_WARPS][2 * NT];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (use_inline) {
#pragma unroll
        for (int i = 0; i < P * D; ++i) {                // fully unrolled: the constant-bank loads issue back to back
            if (i < M * D) {
                const double v = pp.v[i];
                if (tid == i) s_pts[i] = v;
            }
        }
    } else {
        for (int i = tid; i < M * D; i += RF_THREADS) s_pts[i] = pts[i];
    }
    __syncthreads();
    const int jj = tid & (RF_B - 1), a0 = tid / RF_B;
    const int n = blockIdx.x * RF_B + jj;
    double x[D];
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] = n < N ? X[(int64_t)n * D + d] : 0.0;
    const double al = n < N ? alpha[n] : 0.0;
    double* out = img + (size_t)blockIdx.x * 4 * P * RF_B;
    // element (a, i) at (((a >> 3) 8 + (i >> 3)) 32 + 4 (a & 7) + ((i >> 1) & 3)) 2 + (i & 1): as the A operand of the first
    // contraction one 16-byte load per lane feeds the two k-steps that contract i = 8 h + 2 t and 8 h + 2 t + 1 (t = lane & 3);
    // the same order is the B operand of the fold (gen_k below), so slot 0 serves both
    const int ta = ((((jj >> 3) * 32) + 4 * a0 + ((jj >> 1) & 3)) * 2) + (jj & 1);
#pragma unroll
    for (int r = 0; r < 2 * NT; ++r) {
        const int a = a0 + 4 * r;
        double v = 0.0;
        if (a < M) v = n < N ? rf_rbf<D>(kern, s_pts + a * D, x) : 0.0;
        out[ta + (r >> 1) * 512 + (r & 1) * 32] = v;
        out[P * RF_B + ta + (r >> 1) * 512 + (r & 1) * 32] = 0.5 * v;
        double sacc = v * al;                             // point a lives in warps 2 a0 (jj < 32) and 2 a0 + 1
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
        if (lane == 0) s_red[warp][r] = sacc;
    }
    __syncthreads();
    if (tid < P) out[3 * P * RF_B + tid] = s_red[2 * (tid & 3)][tid >> 2] + s_red[2 * (tid & 3) + 1][tid >> 2];
}

// STREAM: large beliefs.  The images come from rollout_images_kernel (computed once instead of once per tile: at N = 4096
// a tile cost 6.1k SM clocks of which the tensor pipe needs 1.5k -- the FP64 exp of the images and the register-staged,
// 8-byte loads of W^-1 made up the rest).  Tiles of W^-1 and the A images arrive through a ring of RF_STAGES stages filled
// by the TMA engine at the request of a ninth warp and handed over with mbarriers -- full: the bytes, empty: 8 arrivals --
// so the tile loop has no block-wide barrier: the two contracting warps of an SM sub-partition drift apart and one's
// operand reads run under the other's DMMAs.
template <int D, int NT, bool STREAM>
__global__ void __launch_bounds__(STREAM ? RF_THREADS + 32 : RF_THREADS, 1) rollout_fused_kernel(const __grid_constant__ RolloutFusedArgs q,
                                                                     const __grid_constant__ RolloutInlineSmall pp,
                                                                     const __grid_constant__ CUtensorMap wmap) {
    constexpr int P = 8 * NT;
    extern __shared__ __align__(16) double sm[];
    constexpr int VP = (P * (P + 1) / 2 + P + 1) & ~1;
    double* s_pts = sm;                                  // [P][3]
    // one-shot form: [2][P * 64] fragment-ordered K[a][i] (A operand), i in tile row I; streaming form: the ring (after the
    // fixed part).  Either way the memory the hand-over and the conditioning reuse once the tile loop is over.
    double* s_ki = STREAM ? s_pts + P * 3 + P * RF_B + 2 * P + VP : s_pts + P * 3;
    double* s_kj = STREAM ? s_pts + P * 3 : s_ki + 2 * P * RF_B;      // [P * 64] fragment-ordered K[b][j] (B operand of the fold), column J
    double* s_kp = s_kj + P * RF_B;                      // [P][64]     the same, plain (one-shot form: mean on the diagonal tile)
    double* s_m0 = STREAM ? s_kj + P * RF_B : s_kp + P * RF_B;        // [P]
    double* s_z = s_m0 + P;                              // [P]         observations
    double* s_part = s_z + P;                            // [V]         this CTA's partial: A + A^T lower triangle, then the mean
    __shared__ int s_last, s_bad;
    __shared__ unsigned int s_colcnt;
    __shared__ __align__(8) unsigned long long s_bar[2 * RF_STAGES + 1];      // full[], empty[], column image
    const bool producer = STREAM && threadIdx.x >= RF_THREADS;       // the warp that feeds the ring (see rf_tid)
    const int tid = rf_tid();
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int M = q.r.M, N = q.N, T = q.T;
    const long long clk0 = clock64();
    // cluster mates store into the leader's shared memory at the end: a (split) cluster barrier first makes sure every CTA
    // of the cluster has started -- arrive here, wait just before the stores, by which time it has long completed
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned int crank = cluster.block_rank(), csize = cluster.num_blocks();
    if (!STREAM && csize > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");

    if (q.use_inline) {          // parameter space is constant memory: every thread reads the SAME word (a broadcast;
#pragma unroll                   // thread-dependent indices would be replayed address by address).  Fully unrolled: the
        for (int i = 0; i < P * D; ++i) {                 // loads have constant addresses and issue back to back (a rolled
            if (i < M * D) {                              // loop paid one load latency per point coordinate: ~1.5k clocks)
                const double v = pp.v[i];
                if (tid == i) s_pts[i] = v;
            }
        }
    } else {
        for (int i = tid; i < M * D; i += RF_THREADS) s_pts[i] = q.pts[i];
    }
    if (tid < P) s_m0[tid] = 0.0;
    __syncthreads();

    const int G = gridDim.x;
    const int k0 = (int)((blockIdx.x * (unsigned)q.tiles) / (unsigned)G);                  // tiles * SMs < 2^31
    const int k1 = (int)(((blockIdx.x + 1u) * (unsigned)q.tiles) / (unsigned)G);
    // column-major walk over the lower triangle: column J holds tiles (I = J .. T - 1, J)
    // (tiles ahead of column J: J T - J (J - 1) / 2; the square root lands within one column, the loops settle it)
    int J = (int)(0.5 * ((double)(2 * T + 1) - sqrt((double)(2 * T + 1) * (double)(2 * T + 1) - 8.0 * (double)k0)));
    J = J < 0 ? 0 : (J > T - 1 ? T - 1 : J);
    while (J > 0 && J * T - J * (J - 1) / 2 > k0) --J;
    while (J < T - 1 && (J + 1) * T - (J + 1) * J / 2 <= k0) ++J;
    int I = J + (k0 - (J * T - J * (J - 1) / 2));
    // B operand of the first contraction, one tile: thread (g, t) of warp w supplies W[I 64 + 32 c + 8 t + s][J 64 + 8 w + g]
    auto load_w = [&](double (&wv)[2][8], int I_, int J_) {
        const int col = J_ * RF_B + warp * 8 + g;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const int row = I_ * RF_B + 32 * c + 8 * t + s;
                wv[c][s] = (row < N && col < N) ? __ldg(q.W + (int64_t)row * q.ldw + col) : 0.0;
            }
    };
    // K[a][x_n] for the 64 observations n of block blk, all P rows: thread -> one observation, rows tid / 64 + 4 r
    auto gen_k = [&](int blk, double scale, double* frag_a, double* frag_b, double* plain) {
        const int jj = tid & (RF_B - 1);
        const int n = blk * RF_B + jj;
        double x[D];
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = n < N ? q.X[(int64_t)n * D + d] : 0.0;
        const int a0 = tid / RF_B;
        double v[2 * NT];
#pragma unroll
        for (int r = 0; r < 2 * NT; ++r) {
            const int a = a0 + 4 * r;                     // a0 is the same for the whole warp: the skip is a real branch
            v[r] = 0.0;
            if (a < M) v[r] = n < N ? rf_rbf<D>(q.kern, s_pts + a * D, x) : 0.0;
        }
        // element (a, jj) with a = a0 + 4 r, i.e. a >> 3 = r >> 1 and a & 7 = a0 + 4 (r & 1): everything that depends on
        // the thread is hoisted, what depends on r is a compile-time constant of the unrolled loop
        //   A operand image: ((((a>>3) 2 + c) 4 + qq) 32 + 4 (a&7) + tt) 2 + e,  jj = 32 c + 8 tt + 2 qq + e
        //   B operand image: (((a>>3) 8 + w) 32 + 4 (a&7) + tt') 2 + e,          jj = 8 w + 2 tt' + e
        const int ta = ((((jj >> 5) * 4 + ((jj >> 1) & 3)) * 32 + 4 * a0 + ((jj >> 3) & 3)) * 2) + (jj & 1);
        const int tb = (((jj >> 3) * 32 + 4 * a0 + ((jj >> 1) & 3)) * 2) + (jj & 1);
#pragma unroll
        for (int r = 0; r < 2 * NT; ++r) {
            if (frag_a) frag_a[ta + (r >> 1) * 512 + (r & 1) * 32] = scale * v[r];
            if (frag_b) {
                frag_b[tb + (r >> 1) * 512 + (r & 1) * 32] = v[r];
                plain[(a0 + 4 * r) * RF_B + jj] = v[r];
            }
        }
    };
    double U[NT][2], U2[NT][2], Z[NT][NT][2];          // (U2: the streaming form's second accumulator set, odd k-steps)
#pragma unroll
    for (int mt = 0; mt < NT; ++mt) {
        U[mt][0] = U[mt][1] = 0.0;
        U2[mt][0] = U2[mt][1] = 0.0;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) Z[mt][nt][0] = Z[mt][nt][1] = 0.0;
    }
    auto fold = [&]() {                                   // A += U_J K_J^T straight from the accumulator fragments
        const double2* kb = reinterpret_cast<const double2*>(s_kj);
        if constexpr (STREAM) {
#pragma unroll
            for (int mt = 0; mt < NT; ++mt) {
                U[mt][0] += U2[mt][0]; U[mt][1] += U2[mt][1];
                U2[mt][0] = U2[mt][1] = 0.0;
            }
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const double2 b = kb[(nt * 8 + warp) * 32 + lane];
#pragma unroll
            for (int mt = 0; mt < NT; ++mt) {
                dmma884(Z[mt][nt][0], Z[mt][nt][1], U[mt][0], b.x);
                dmma884(Z[mt][nt][0], Z[mt][nt][1], U[mt][1], b.y);
            }
        }
#pragma unroll
        for (int mt = 0; mt < NT; ++mt) U[mt][0] = U[mt][1] = 0.0;
    };

    int curJ = -1, buf = 0;
    double w_cur[2][8], w_nxt[2][8];
    long long tk[6] = {0, 0, 0, 0, 0, 0};
    tk[0] = clock64();
    if constexpr (STREAM) {
        // one stage: the tile of W^-1 as four panels of 16 observations i -- [64 rows j][128 B], 16-byte chunk c of row j at
        // c ^ (j & 7) (the TMA engine's 128-byte swizzle) --, then the A image of block I.  W^-1 is symmetric: the tile is
        // fetched from the mirrored position (rows J, columns I), so that a lane's two k-steps are ONE 16-byte load and
        // the 8 rows x 4 chunks of a warp's load spread over all banks.
        constexpr int STG = RF_TILE + P * RF_B;
        const uint32_t ring0 = (smem_addr(s_ki) + 1023u) & ~1023u;
        const uint32_t bar0 = smem_addr(s_bar);
        const uint32_t col_bar = bar0 + 8 * 2 * RF_STAGES;
        if (tid == 0) {
            for (int i = 0; i < RF_STAGES; ++i) { mbar_init(bar0 + 8 * i, 1); mbar_init(bar0 + 8 * (RF_STAGES + i), RF_WARPS); }
            mbar_init(col_bar, 1);
            s_colcnt = 0;
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        const int cnt = k1 - k0;
        auto issue_column = [&](int J_) {                 // the image of block J_ as the fold's B operand (one thread)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(col_bar, P * RF_B * 8);
            bulk_load(smem_addr(s_kj), q.img + (size_t)J_ * 4 * (P * RF_B), P * RF_B * 8, col_bar);
        };
        auto next_tile = [&](int& I_, int& J_) { if (++I_ == T) { ++J_; I_ = J_; } };
        if (producer) {
            // The ninth warp feeds the ring: local tile number n -> stage n % RF_STAGES, asked for again once all eight
            // contracting warps have let go of it.  Per tile four tensor copies (one 16-column panel each; rows / columns
            // beyond N arrive as zeros) and one bulk copy (the A image).  A copy instruction holds its thread until the TMA
            // engine takes it (~600 clocks per tile when the contracting warps issued them themselves), hence a warp of its
            // own.  The panels do not depend on the pre-pass: the first RF_STAGES tiles' go out before griddepcontrol.wait.
            auto issue_panels = [&](int n, int I_, int J_) {
                const int sg = n % RF_STAGES;
                const uint32_t full = bar0 + 8 * sg;
                mbar_expect_tx(full, (RF_TILE + P * RF_B) * 8);
#pragma unroll
                for (int pnl = 0; pnl < 4; ++pnl)
                    tma_load_2d(ring0 + (uint32_t)(sg * STG * 8 + pnl * RF_B * 16 * 8), &wmap, I_ * RF_B + 16 * pnl, J_ * RF_B, full);
            };
            auto issue_image = [&](int n, int I_, int J_) {
                const int sg = n % RF_STAGES;
                bulk_load(ring0 + (uint32_t)((sg * STG + RF_TILE) * 8), q.img + ((size_t)I_ * 4 + (I_ == J_ ? 1 : 0)) * (P * RF_B),
                          P * RF_B * 8, bar0 + 8 * sg);
            };
            const bool leader = elect_one();
            int Ia = I, Ja = J, na = 0;
            for (; na < cnt && na < RF_STAGES; ++na) { if (leader) issue_panels(na, Ia, Ja); next_tile(Ia, Ja); }
            grid_dependency_wait();                       // the images are the pre-pass kernel's output
            if (leader) {
                int Ib = I, Jb = J;
                for (int nb = 0; nb < na; ++nb) { issue_image(nb, Ib, Jb); next_tile(Ib, Jb); }
                if (cnt > 0) issue_column(J);
            }
            for (; na < cnt; ++na) {
                mbar_wait(bar0 + 8 * (RF_STAGES + na % RF_STAGES), (uint32_t)((na / RF_STAGES - 1) & 1));
                if (leader) { issue_panels(na, Ia, Ja); issue_image(na, Ia, Ja); }
                next_tile(Ia, Ja);
            }
        } else {
        grid_dependency_wait();
        tk[1] = clock64();                                // (diagnostics: tk[2..4] accumulate column / wait / contraction clocks)
        uint32_t col_phase = 0;
        for (int n = 0; n < cnt; ++n) {
            const long long c0 = clock64();
            if (J != curJ) {
                if (curJ >= 0) {
                    mbar_wait(col_bar, col_phase);
                    col_phase ^= 1u;
                    fold();
                    __syncwarp();
                    if (lane == 0) {                      // the last warp to let go of the old column's image asks for the new one
                        __threadfence_block();
                        const unsigned int old = atomicAdd(&s_colcnt, 1u);
                        if ((old & (RF_WARPS - 1)) == RF_WARPS - 1) { __threadfence_block(); issue_column(J); }
                    }
                    __syncwarp();
                }
                curJ = J;
            }
            const long long c1 = clock64();
            mbar_wait(bar0 + 8 * (n % RF_STAGES), (uint32_t)((n / RF_STAGES) & 1));
            const long long c2 = clock64();
            if (I == J && tid < P) s_m0[tid] += __ldcg(q.img + ((size_t)J * 4 + 3) * (P * RF_B) + tid);      // K_J alpha_J (pre-pass)
            // B operand of k-step (p, h, e): W[i = 16 p + 8 h + 2 t + e][j = 8 w + g] = mirrored tile row 8 w + g, chunk 4 h + t
            const uint32_t wrow = ring0 + (uint32_t)((n % RF_STAGES) * STG * 8 + (8 * warp + g) * 128);
            // (odd rows load their two chunks in the other order, so that the two rows of a quarter-warp never meet in a bank)
            const bool odd = (g & 1) != 0;
            const uint32_t wc0 = (uint32_t)((((odd ? 4 : 0) + t) ^ g) << 4), wc1 = (uint32_t)((((odd ? 0 : 4) + t) ^ g) << 4);
            const uint32_t img0 = ring0 + (uint32_t)(((n % RF_STAGES) * STG + RF_TILE) * 8 + lane * 16);
            // (the loads are asm volatile like the DMMAs: source order is issue order, so panel pnl + 1 is asked for before
            // the sixteen-clock-apart DMMAs of panel pnl are issued)
            double2 la[2], lb[2], a0[2][NT], a1[2][NT];
            auto load_panel = [&](int pnl, int bf) {
                la[bf] = lds128(wrow + pnl * (RF_B * 16 * 8) + wc0);
                lb[bf] = lds128(wrow + pnl * (RF_B * 16 * 8) + wc1);
#pragma unroll
                for (int mt = 0; mt < NT; ++mt) {
                    a0[bf][mt] = lds128(img0 + (uint32_t)(((mt * 8 + 2 * pnl) * 32) * 16));
                    a1[bf][mt] = lds128(img0 + (uint32_t)(((mt * 8 + 2 * pnl + 1) * 32) * 16));
                }
            };
            load_panel(0, 0);
#pragma unroll
            for (int pnl = 0; pnl < 4; ++pnl) {
                const int bf = pnl & 1;
                if (pnl + 1 < 4) load_panel(pnl + 1, bf ^ 1);
                const double2 w0 = odd ? lb[bf] : la[bf], w1 = odd ? la[bf] : lb[bf];
                // two accumulator sets: a DMMA waits ~110 clocks for the one before it on the same accumulator, and NT chains
                // per warp (two warps per tensor pipe) leave the pipe idle 1 clock in 8
#pragma unroll
                for (int mt = 0; mt < NT; ++mt) dmma884(U[mt][0], U[mt][1], a0[bf][mt].x, w0.x);
#pragma unroll
                for (int mt = 0; mt < NT; ++mt) dmma884(U2[mt][0], U2[mt][1], a0[bf][mt].y, w0.y);
#pragma unroll
                for (int mt = 0; mt < NT; ++mt) dmma884(U[mt][0], U[mt][1], a1[bf][mt].x, w1.x);
#pragma unroll
                for (int mt = 0; mt < NT; ++mt) dmma884(U2[mt][0], U2[mt][1], a1[bf][mt].y, w1.y);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8 * (RF_STAGES + n % RF_STAGES));      // this warp is done with the stage
            next_tile(I, J);
            const long long c3 = clock64();
            tk[2] += c1 - c0; tk[3] += c2 - c1; tk[4] += c3 - c2;
        }
        if (cnt > 0) {
            mbar_wait(col_bar, col_phase);
            fold();
        }
        tk[2] += tk[1]; tk[3] += tk[2]; tk[4] += tk[3];   // so that the differences printed below are the three sums
        }
    } else {
    if (k0 < k1) load_w(w_nxt, I, J);
    for (int k = k0; k < k1; ++k) {
        if (J != curJ) {
            if (curJ >= 0) {
                fold();
                __syncthreads();                          // everybody is done with the old column's image
            }
            curJ = J;
            gen_k(J, 1.0, nullptr, s_kj, s_kp);
        }
        if (k == k0) tk[1] = clock64();
        gen_k(I, I == J ? 0.5 : 1.0, s_ki + buf * P * RF_B, nullptr, nullptr);
        if (k == k0) tk[2] = clock64();
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int s = 0; s < 8; ++s) w_cur[c][s] = w_nxt[c][s];
        int In = I + 1, Jn = J;
        if (In == T) { ++Jn; In = Jn; }
        if (k + 1 < k1) load_w(w_nxt, In, Jn);
        __syncthreads();
        if (k == k0) tk[3] = clock64();
        if (I == J) {                                     // partial mean: K_J alpha_J, 8 lanes per point
            for (int idx = tid; idx < P * 8; idx += RF_THREADS) {
                const int a = idx >> 3, seg = idx & 7;
                double sacc = 0.0;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int n = J * RF_B + seg * 8 + u;
                    sacc = fma(s_kp[a * RF_B + seg * 8 + u], n < N ? q.alpha[n] : 0.0, sacc);
                }
                sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
                sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
                sacc += __shfl_xor_sync(0xffffffffu, sacc, 4);
                if (seg == 0) s_m0[a] += sacc;
            }
        }
        const double2* ka = reinterpret_cast<const double2*>(s_ki + buf * P * RF_B);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            double a[NT][8];
#pragma unroll
            for (int mt = 0; mt < NT; ++mt)
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    const double2 v = ka[((mt * 2 + c) * 4 + qq) * 32 + lane];
                    a[mt][2 * qq] = v.x;
                    a[mt][2 * qq + 1] = v.y;
                }
#pragma unroll
            for (int s = 0; s < 8; ++s)
#pragma unroll
                for (int mt = 0; mt < NT; ++mt) dmma884(U[mt][0], U[mt][1], a[mt][s], w_cur[c][s]);
        }
        buf ^= 1;
        I = In; J = Jn;
        if (k == k0) tk[4] = clock64();
    }
    }
    if (!STREAM && k0 < k1) fold();                       // (a CTA without tiles only pads its cluster: its partial is zero)
    __syncthreads();
    tk[5] = clock64();

    // ---- A summed over the 8 warps (fixed order), symmetrised, and handed over together with the partial mean
    double* s_z8 = s_ki;                                  // [8][P * P], over the dead images (4 P 64 >= 8 P P for P <= 32)
    if (!producer)
#pragma unroll
    for (int mt = 0; mt < NT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            s_z8[warp * P * P + (mt * 8 + g) * P + nt * 8 + 2 * t] = Z[mt][nt][0];
            s_z8[warp * P * P + (mt * 8 + g) * P + nt * 8 + 2 * t + 1] = Z[mt][nt][1];
        }
    __syncthreads();
    const int V = M * (M + 1) / 2 + M;
    for (int a = warp; a < M; a += RF_WARPS)              // pairs b <= a: lane = b (M <= 32)
        if (lane <= a) {
            double s_ab = 0.0, s_ba = 0.0;
#pragma unroll
            for (int w = 0; w < RF_WARPS; ++w) {
                s_ab += s_z8[w * P * P + a * P + lane];
                s_ba += s_z8[w * P * P + lane * P + a];
            }
            s_part[a * (a + 1) / 2 + lane] = s_ab + s_ba;
        }
    if (tid < M) s_part[M * (M + 1) / 2 + tid] = s_m0[tid];
    // ---- the RF_CLUSTER partials of a thread-block cluster meet in its leader through distributed shared memory: the
    // finishing CTA then reads a quarter of the partials.  That read is one SM pulling V x (number of partials) doubles
    // through its L2 port and was the longest single step of a rollout at N = 900 (7.5k SM clocks).  The mates PUSH: each
    // stores its partial into its own slot of the leader's shared memory, arrives at the cluster barrier (release) and is
    // done; only the leader waits (acquire) and adds the slots in rank order.  (First version: the leader pulled the mates'
    // partials -- a second cluster barrier had to keep their shared memory alive, and the remote loads were a round trip.)
    // (Large beliefs are launched with clusters of ONE CTA: whole clusters need RF_CLUSTER free SMs in one GPC, fewer of
    // them are co-resident than single CTAs, and there the partial sum is a small part of the rollout.)
    __syncthreads();                                      // s_part complete
    if (csize > 1) {
        if constexpr (!STREAM) {
            double* s_gather = s_part + VP;               // [RF_CLUSTER - 1][VP], the leader's copy is the one written
            asm volatile("barrier.cluster.wait.aligned;" ::: "memory");      // (the start-up barrier)
            if (crank != 0) {
                double* dst = cluster.map_shared_rank(s_gather, 0) + (size_t)(crank - 1) * VP;
                for (int v = tid; v < V; v += RF_THREADS) dst[v] = s_part[v];
            }
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
            if (crank != 0) return;
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
            double* mine = q.part + (size_t)(blockIdx.x / csize) * V;
            for (int v = tid; v < V; v += RF_THREADS) {
                double tot = s_part[v];
                for (unsigned int r = 1; r < csize; ++r) tot += s_gather[(size_t)(r - 1) * VP + v];
                mine[v] = tot;
            }
        } else {
            __trap();                                     // the streaming form is launched with clusters of one CTA
        }
    } else {
        double* mine = q.part + (size_t)blockIdx.x * V;
        for (int v = tid; v < V; v += RF_THREADS) mine[v] = s_part[v];
    }
    int GC = G / (int)csize;                              // partials = clusters
    const double* fin = q.part;                           // what the last CTA adds up: GC partials of V numbers
    const long long clk1 = clock64();
    if constexpr (STREAM) {
        // Large beliefs run one partial per SM: the finishing CTA pulling 148 x V doubles through its L2 port was 9k of
        // the kernel's 62k clocks at N = 4096.  Two levels, both in a fixed order: the last CTA of every group of RF_GROUP
        // (a ticket per group) adds its group's partials, and the last group to finish adds the group sums.
        if (GC > 2 * RF_GROUP && GC <= RF_GROUP * RF_MAX_GROUPS) {
            const int grp = blockIdx.x / RF_GROUP, ngroups = (GC + RF_GROUP - 1) / RF_GROUP;
            const int gsize = (GC - grp * RF_GROUP) < RF_GROUP ? (GC - grp * RF_GROUP) : RF_GROUP;
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                const unsigned int old = atomicAdd(&g_rf_ticket2[q.slot][grp], 1u);
                s_last = (old == (unsigned int)(gsize - 1));
                if (s_last) g_rf_ticket2[q.slot][grp] = 0u;
            }
            __syncthreads();
            if (!s_last) return;
            __threadfence();
            for (int v = tid; v < V; v += RF_THREADS) {
                const double* src = q.part + (size_t)grp * RF_GROUP * V + v;
                double acc[RF_GROUP];
#pragma unroll
                for (int u = 0; u < RF_GROUP; ++u) acc[u] = (u < gsize) ? __ldcg(src + (size_t)u * V) : 0.0;
                q.part2[(size_t)grp * V + v] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
            }
            fin = q.part2;
            GC = ngroups;
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int old = atomicAdd(&g_rf_ticket[q.slot], 1u);
        s_last = (old == (unsigned int)(GC - 1));
        if (s_last) g_rf_ticket[q.slot] = 0u;             // everybody has arrived: the slot is free again
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const long long clk2 = clock64();

    // ---- the last CTA: partials summed in CTA order (16 loads in flight per thread), C0 and m0, conditioning
    const int ldc = M + 1;
    double* C = s_ki;
    double* m = C + (size_t)M * ldc;
    double* H = m + M;
    double* snap_m = H + (size_t)(M + 1) * (RO_MAX_BLOCK + 1);
    double* snap_v = snap_m + M;
    double* Cs = snap_v + M;                              // [M][M+1] (information gain: the levels' blocks)
    __syncthreads();                                      // s_z8 (same memory as C) has been consumed by everybody
    for (int v = tid; v < V; v += RF_THREADS) {
        // fixed order over the clusters -- the result does not depend on their timing --, 32 independent loads in flight.
        // (Tried: exact 128-bit fixed-point accumulators fed by atomics, so that only V numbers are read here: the
        // atomics of ~120 CTAs on the same V addresses cost more than this sum saves, profiles/r2g_*.)
        const double* src = fin + v;
        double acc[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) acc[u] = 0.0;
        int c = 0;
        for (; c + 32 <= GC; c += 32) {
#pragma unroll
            for (int u = 0; u < 32; ++u) acc[u] += __ldcg(src + (size_t)(c + u) * V);
        }
        if (c < GC) {
#pragma unroll
            for (int u = 0; u < 32; ++u) acc[u] += (c + u < GC) ? __ldcg(src + (size_t)(c + u) * V) : 0.0;
        }
#pragma unroll
        for (int w = 16; w > 0; w >>= 1)                  // pairwise: five dependent additions instead of 31
#pragma unroll
            for (int u = 0; u < w; ++u) acc[u] += acc[u + w];
        const double tot = acc[0];
        if (v < M * (M + 1) / 2) {
            int a = (int)((sqrt(8.0 * v + 1.0) - 1.0) * 0.5);
            while (a * (a + 1) / 2 > v) --a;
            while ((a + 1) * (a + 2) / 2 <= v) ++a;
            const int b = v - a * (a + 1) / 2;
            const double val = rf_rbf<D>(q.kern, s_pts + a * D, s_pts + b * D) - tot;
            C[a * ldc + b] = val;
            C[b * ldc + a] = val;
        } else {
            m[v - M * (M + 1) / 2] = tot;
        }
    }
    if (q.use_inline) {
        for (int i = 0; i < M; ++i) {
            const double v = pp.v[M * D + i];
            if (tid == i) s_z[i] = v;
        }
    } else {
        for (int i = tid; i < M; i += RF_THREADS) s_z[i] = q.zobs[i];
    }
    const long long clk3 = clock64();
    rf_condition(q.r, s_z, C, m, H, snap_m, snap_v, Cs, s_bad, q.tag);
    if (tid == 0) {
        g_rf_clock[0] = clk1 - clk0; g_rf_clock[1] = clk2 - clk1; g_rf_clock[2] = clk3 - clk2; g_rf_clock[3] = clock64() - clk3;
        g_rf_clock[4] = tk[0] - clk0;     // points + setup
        g_rf_clock[5] = tk[1] - tk[0];    // first W loads issued + column image
        g_rf_clock[6] = tk[2] - tk[1];    // row image
        g_rf_clock[7] = tk[3] - tk[2];    // barrier (W loads of the next tile issued)
        g_rf_clock[8] = tk[4] - tk[3];    // mean + first contraction
        g_rf_clock[9] = tk[5] - tk[4];    // remaining tiles + fold
        g_rf_clock[10] = clk1 - tk[5];    // 8-warp sum + partial store
    }
}

}  // namespace ipp
