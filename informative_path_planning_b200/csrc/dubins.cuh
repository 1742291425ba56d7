// Dubins closed forms + the trimming loop of Dubins_Path_Generator.buffered_paths (reference paths_library.py:147-175),
// shared by the host path-set generator / device batch kernel (paths.cu) and the native tree search (tree.cu).
// Classical six-word closed form (Shkel & Lumelsky 2001) as the `dubins` C library behind paths_library.py:151-152 states
// it, word by word and in its enum order (pinned by tests/test_dubins.py against an independent geometric construction).
#pragma once
#include <math.h>
#include "internal.cuh"

namespace ipp_dubins {

#define TWO_PI 6.283185307179586476925286766559

__host__ __device__ inline double mod2pi(double t) { return t - TWO_PI * floor(t / TWO_PI); }

struct Word { int kinds[3]; double len[3]; bool ok; };   // kind: 0 = L, 1 = S, 2 = R

__host__ __device__ inline Word make_word(int k0, int k1, int k2, double a, double b, double c) {
    Word w;
    w.kinds[0] = k0; w.kinds[1] = k1; w.kinds[2] = k2;
    w.len[0] = a; w.len[1] = b; w.len[2] = c;
    w.ok = true;
    return w;
}

// The six words in the order of the `dubins` C library's DubinsPathType enum (LSL, LSR, RSL, RSR, RLR, LRL): the library
// keeps a word only when it is STRICTLY shorter than the best so far, so this order is the tie-break.
__host__ __device__ inline void all_words(double alpha, double beta, double d, Word out[6]) {
    const double sa = sin(alpha), sb = sin(beta), ca = cos(alpha), cb = cos(beta), cab = cos(alpha - beta);
    const double d_sq = d * d;
    for (int i = 0; i < 6; ++i) out[i].ok = false;
    double p2, tmp, p, t, phi;
    p2 = 2 + d_sq - (2 * cab) + (2 * d * (sa - sb));                                // LSL
    if (p2 >= 0) { tmp = atan2(cb - ca, d + sa - sb);
        out[0] = make_word(0, 1, 0, mod2pi(tmp - alpha), sqrt(p2), mod2pi(beta - tmp)); }
    p2 = -2 + d_sq + (2 * cab) + (2 * d * (sa + sb));                               // LSR
    if (p2 >= 0) { p = sqrt(p2); tmp = atan2(-ca - cb, d + sa + sb) - atan2(-2.0, p);
        out[1] = make_word(0, 1, 2, mod2pi(tmp - alpha), p, mod2pi(tmp - mod2pi(beta))); }
    p2 = -2 + d_sq + (2 * cab) - (2 * d * (sa + sb));                               // RSL
    if (p2 >= 0) { p = sqrt(p2); tmp = atan2(ca + cb, d - sa - sb) - atan2(2.0, p);
        out[2] = make_word(2, 1, 0, mod2pi(alpha - tmp), p, mod2pi(beta - tmp)); }
    p2 = 2 + d_sq - (2 * cab) + (2 * d * (sb - sa));                                // RSR
    if (p2 >= 0) { tmp = atan2(ca - cb, d - sa + sb);
        out[3] = make_word(2, 1, 2, mod2pi(alpha - tmp), sqrt(p2), mod2pi(tmp - beta)); }
    tmp = (6.0 - d_sq + 2 * cab + 2 * d * (sa - sb)) / 8.0;                         // RLR
    if (fabs(tmp) <= 1) { phi = atan2(ca - cb, d - sa + sb); p = mod2pi(TWO_PI - acos(tmp));
        t = mod2pi(alpha - phi + mod2pi(p / 2.0));
        out[4] = make_word(2, 0, 2, t, p, mod2pi(alpha - beta - t + mod2pi(p))); }
    tmp = (6.0 - d_sq + 2 * cab + 2 * d * (sb - sa)) / 8.0;                         // LRL
    if (fabs(tmp) <= 1) { phi = atan2(ca - cb, d + sa - sb); p = mod2pi(TWO_PI - acos(tmp));
        t = mod2pi(-alpha - phi + p / 2.0);
        out[5] = make_word(0, 2, 0, t, p, mod2pi(mod2pi(beta) - alpha - t + mod2pi(p))); }
}

__host__ __device__ inline void segment(double& x, double& y, double& th, double len, int kind) {
    if (kind == 0) { x += sin(th + len) - sin(th); y += -cos(th + len) + cos(th); th += len; }
    else if (kind == 2) { x += -sin(th - len) + sin(th); y += cos(th - len) - cos(th); th -= len; }
    else { x += cos(th) * len; y += sin(th) * len; }
}


struct ObstacleSet { ipp_obstacle_box box[IPP_MAX_OBSTACLES]; int n; };

// reference obstacles.py:58-75 / :145-163 / :189-193 in the box-with-hole form of include/ipp_b200.h
__host__ __device__ inline bool in_obstacle(const ipp_obstacle_box* obs, int n, double x, double y, double buff) {
    for (int i = 0; i < n; ++i) {
        const ipp_obstacle_box& b = obs[i];
        if (x > b.outer[0] - buff && x < b.outer[1] + buff && y > b.outer[2] - buff && y < b.outer[3] + buff) {
            if (b.has_hole && x >= b.hole[0] + b.hole_buff[0] * buff && x <= b.hole[1] + b.hole_buff[1] * buff &&
                y >= b.hole[2] + b.hole_buff[2] * buff && y <= b.hole[3] + b.hole_buff[3] * buff)
                continue;
            return true;
        }
    }
    return false;
}

// One (pose, goal) pair.  dense (may be NULL on the device path): every configuration; samp: the kept ones.
// Returns the number of kept samples (0 = path dropped); *dense_n = number of dense configurations of the true path;
// end[3] = its last configuration (the pose of the next belief node).
__host__ __device__ inline int dubins_one(const double* pose, const double* q1, double rho, double dense_step, int keep_every,
                                   const double* extent, double border, const ipp_obstacle_box* obs, int n_obs,
                                   double* sg, int max_samp, double* dg, int max_dense, int* dense_n, double* end) {
    const double dx = q1[0] - pose[0], dy = q1[1] - pose[1];
    const double D = sqrt(dx * dx + dy * dy), d = D / rho;      // the library's own expression (not hypot)
    const double theta = d > 0 ? mod2pi(atan2(dy, dx)) : 0.0;
    const double alpha = mod2pi(pose[2] - theta), beta = mod2pi(q1[2] - theta);
    Word w[6];
    all_words(alpha, beta, d, w);
    int best = -1;
    double best_len = 0;
    for (int i = 0; i < 6; ++i) {
        if (!w[i].ok) continue;
        const double L = w[i].len[0] + w[i].len[1] + w[i].len[2];
        if (best < 0 || L < best_len) { best = i; best_len = L; }
    }
    *dense_n = 0;
    if (best < 0) return 0;
    const Word wb = w[best];
    const double total = best_len * rho;
    // state at the start of each of the three segments (a fully traversed segment always contributes the same
    // increment, so it is evaluated once per path instead of once per sample: 2 trig calls per sample, not 12)
    double sx[4], sy[4], sth[4];
    sx[0] = 0; sy[0] = 0; sth[0] = pose[2];
    for (int k = 0; k < 3; ++k) {
        sx[k + 1] = sx[k]; sy[k + 1] = sy[k]; sth[k + 1] = sth[k];
        segment(sx[k + 1], sy[k + 1], sth[k + 1], wb.len[k] < 0 ? 0 : wb.len[k], wb.kinds[k]);
    }
    // dense sweep: cut at the first configuration outside the open extent; every keep_every-th one is a sample
    // candidate, trimmed by the reference's loop (python: `ttemp = ttemp[0:m-1]`, including its negative-index quirk)
    int nd = 0, orig = 0, cur = 0;
    // When the dense configurations are not asked for (dg == NULL) and the whole curve provably stays inside the extent
    // (start pose farther from every wall than the path is long), only every keep_every-th configuration -- the sample
    // candidates -- needs evaluating: each one is a closed form of its own arc length, so the values are identical.
    const double margin = total + 1e-9;
    const bool interior = pose[0] - extent[0] > margin && extent[1] - pose[0] > margin &&
                          pose[1] - extent[2] > margin && extent[3] - pose[1] > margin;
    const int stride = (dg == nullptr && interior && n_obs == 0) ? keep_every : 1;
    // arc length advances by repeated addition, as the `dubins` library's sample_many does (x += stepSize while
    // x < length): whether the configuration AT the path's end is sampled is decided by the same rounding
    double s = 0.0;
    for (int si = 0; nd < max_dense; si += stride, nd += stride - 1) {
        if (si) for (int a = 0; a < stride; ++a) s += dense_step;
        if (!(s < total)) break;
        // dubins_path_sample: the piece is chosen by t' < p1, t' < p1 + p2; the offset into it is t', t' - p1, t' - p1 - p2
        const double tp = s / rho;
        const int k = tp < wb.len[0] ? 0 : (tp < (wb.len[0] + wb.len[1]) ? 1 : 2);
        const double use = k == 0 ? tp : (k == 1 ? tp - wb.len[0] : tp - wb.len[0] - wb.len[1]);
        double x = sx[k], y = sy[k], th = sth[k];
        segment(x, y, th, use, wb.kinds[k]);
        const double cx = x * rho + pose[0], cy = y * rho + pose[1];
        if (!(cx > extent[0] && cx < extent[1] && cy > extent[2] && cy < extent[3])) break;
        if (n_obs && in_obstacle(obs, n_obs, cx, cy, 0.0)) break;
        const double cth = mod2pi(th);
        if (dg) { dg[3 * nd] = cx; dg[3 * nd + 1] = cy; dg[3 * nd + 2] = cth; }
        if (si % keep_every == 0 && orig < max_samp) {
            sg[3 * orig] = cx; sg[3 * orig + 1] = cy; sg[3 * orig + 2] = cth;
            ++orig;
        }
        ++nd;
    }
    cur = orig;
    for (int m = 0; m < orig; ++m) {
        const double* c = sg + 3 * m;
        if (c[0] <= extent[0] + border || c[0] >= extent[1] - border || c[1] <= extent[2] + border || c[1] >= extent[3] - border ||
            (n_obs && in_obstacle(obs, n_obs, c[0], c[1], border))) {
            int stop = m - 1;                       // python: ttemp = ttemp[0:m-1]
            if (stop < 0) stop = cur + stop;        // negative stop counts from the end
            if (stop < 0) stop = 0;
            if (stop < cur) cur = stop;
        }
    }
    if (cur < 2) return 0;
    // true_path = ftemp[0 : ftemp.index(ttemp[-1]) + 1]: first dense configuration equal to the last kept one
    const double* last = sg + 3 * (cur - 1);
    int idx = (cur - 1) * keep_every;
    if (dg) {
        for (int i = 0; i < idx; ++i)
            if (dg[3 * i] == last[0] && dg[3 * i + 1] == last[1] && dg[3 * i + 2] == last[2]) { idx = i; break; }
    }
    *dense_n = idx + 1;
    if (end) { end[0] = last[0]; end[1] = last[1]; end[2] = last[2]; }
    return cur;
}

}  // namespace ipp_dubins
