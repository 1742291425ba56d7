// Host-side (CPU, no kernel) Dubins path-set generation: the native counterpart of the `dubins` C library
// the reference calls from paths_library.py:141-175 (dubins.shortest_path + path.sample_many), plus the
// border trimming of Dubins_Path_Generator.buffered_paths for the obstacle-free world.
// Classical six-word closed form (LSL, RSR, LSR, RSL, RLR, LRL; Shkel & Lumelsky 2001).
#include <math.h>
#include "internal.cuh"

namespace {

const double TWO_PI = 6.283185307179586476925286766559;

inline double mod2pi(double t) { return t - TWO_PI * floor(t / TWO_PI); }

struct Word { int kinds[3]; double len[3]; bool ok; };   // kind: 0 = L, 1 = S, 2 = R

void all_words(double alpha, double beta, double d, Word out[6]) {
    const double sa = sin(alpha), sb = sin(beta), ca = cos(alpha), cb = cos(beta), cab = cos(alpha - beta);
    for (int i = 0; i < 6; ++i) out[i].ok = false;
    double p2, tmp, p, t;
    p2 = 2 + d * d - 2 * cab + 2 * d * (sa - sb);                                   // LSL
    if (p2 >= 0) { tmp = atan2(cb - ca, d + sa - sb);
        out[0] = {{0, 1, 0}, {mod2pi(-alpha + tmp), sqrt(p2), mod2pi(beta - tmp)}, true}; }
    p2 = 2 + d * d - 2 * cab + 2 * d * (sb - sa);                                   // RSR
    if (p2 >= 0) { tmp = atan2(ca - cb, d - sa + sb);
        out[1] = {{2, 1, 2}, {mod2pi(alpha - tmp), sqrt(p2), mod2pi(-beta + tmp)}, true}; }
    p2 = -2 + d * d + 2 * cab + 2 * d * (sa + sb);                                  // LSR
    if (p2 >= 0) { p = sqrt(p2); tmp = atan2(-ca - cb, d + sa + sb) - atan2(-2.0, p);
        out[2] = {{0, 1, 2}, {mod2pi(-alpha + tmp), p, mod2pi(-mod2pi(beta) + tmp)}, true}; }
    p2 = d * d - 2 + 2 * cab - 2 * d * (sa + sb);                                   // RSL
    if (p2 >= 0) { p = sqrt(p2); tmp = atan2(ca + cb, d - sa - sb) - atan2(2.0, p);
        out[3] = {{2, 1, 0}, {mod2pi(alpha - tmp), p, mod2pi(beta - tmp)}, true}; }
    tmp = (6.0 - d * d + 2 * cab + 2 * d * (sa - sb)) / 8.0;                        // RLR
    if (fabs(tmp) <= 1) { p = mod2pi(TWO_PI - acos(tmp)); t = mod2pi(alpha - atan2(ca - cb, d - sa + sb) + p / 2.0);
        out[4] = {{2, 0, 2}, {t, p, mod2pi(alpha - beta - t + p)}, true}; }
    tmp = (6.0 - d * d + 2 * cab + 2 * d * (sb - sa)) / 8.0;                        // LRL
    if (fabs(tmp) <= 1) { p = mod2pi(TWO_PI - acos(tmp)); t = mod2pi(-alpha - atan2(ca - cb, d + sa - sb) + p / 2.0);
        out[5] = {{0, 2, 0}, {t, p, mod2pi(mod2pi(beta) - alpha - t + p)}, true}; }
}

inline void segment(double& x, double& y, double& th, double len, int kind) {
    if (kind == 0) { x += sin(th + len) - sin(th); y += -cos(th + len) + cos(th); th += len; }
    else if (kind == 2) { x += -sin(th - len) + sin(th); y += cos(th - len) - cos(th); th -= len; }
    else { x += cos(th) * len; y += sin(th) * len; }
}

}  // namespace

// For each goal g: dense[g][0..dense_count[g]) = configurations every `dense_step` along the shortest Dubins
// path from `pose`, cut at the first one outside the open extent; samp = every `keep_every`-th of them, trimmed
// where a kept point lies within `border` of the extent exactly as the reference's loop does (including its
// `ttemp[0:m-1]` slicing quirk).  samp_count[g] = 0 when fewer than 2 points survive (path dropped, :167-171);
// dense_count[g] then is meaningless.  Arrays are host memory: samp [G][max_samp][3], dense [G][max_dense][3].
extern "C" int ipp_dubins_path_set(const double* pose, const double* goals, int G, double rho, double dense_step,
                                   int keep_every, const double* extent, double border,
                                   double* samp, int* samp_count, int max_samp,
                                   double* dense, int* dense_count, int max_dense) {
    IPP_CHECK_ARG(pose && goals && extent && samp && samp_count && dense && dense_count, "null pointer");
    IPP_CHECK_ARG(G >= 0 && rho > 0 && dense_step > 0 && keep_every >= 1 && max_samp >= 2 && max_dense >= 2, "bad argument");
    for (int gi = 0; gi < G; ++gi) {
        const double* q1 = goals + 3 * gi;
        const double dx = q1[0] - pose[0], dy = q1[1] - pose[1];
        const double D = hypot(dx, dy), d = D / rho;
        const double theta = D > 0 ? mod2pi(atan2(dy, dx)) : 0.0;
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
        samp_count[gi] = 0;
        dense_count[gi] = 0;
        if (best < 0) continue;
        const double total = best_len * rho;
        double* dg = dense + (size_t)gi * max_dense * 3;
        int nd = 0;
        // state at the start of each of the three segments (a fully traversed segment always contributes the same
        // increment, so it is evaluated once per path instead of once per sample: 2 trig calls per sample, not 12)
        double sx[4], sy[4], sth[4];
        sx[0] = 0; sy[0] = 0; sth[0] = pose[2];
        for (int k = 0; k < 3; ++k) {
            sx[k + 1] = sx[k]; sy[k + 1] = sy[k]; sth[k + 1] = sth[k];
            segment(sx[k + 1], sy[k + 1], sth[k + 1], w[best].len[k] < 0 ? 0 : w[best].len[k], w[best].kinds[k]);
        }
        for (int si = 0; nd < max_dense; ++si) {
            const double s = si * dense_step;
            if (!(s < total)) break;
            double rem = s / rho;
            int k = 0;
            while (k < 2 && !(rem < w[best].len[k])) { rem -= (w[best].len[k] < 0 ? 0 : w[best].len[k]); ++k; }
            double x = sx[k], y = sy[k], th = sth[k];
            double use = rem < w[best].len[k] ? rem : w[best].len[k];
            if (use < 0) use = 0;
            segment(x, y, th, use, w[best].kinds[k]);
            const double cx = x * rho + pose[0], cy = y * rho + pose[1];
            if (!(cx > extent[0] && cx < extent[1] && cy > extent[2] && cy < extent[3])) break;
            dg[3 * nd] = cx; dg[3 * nd + 1] = cy; dg[3 * nd + 2] = mod2pi(th);
            ++nd;
        }
        // ttemp = ftemp[0::keep_every], trimmed by the reference's loop
        int orig = (nd + keep_every - 1) / keep_every;
        if (orig > max_samp) orig = max_samp;
        int cur = orig;
        for (int m = 0; m < orig; ++m) {
            const double* c = dg + 3 * (m * keep_every);
            if (c[0] <= extent[0] + border || c[0] >= extent[1] - border || c[1] <= extent[2] + border || c[1] >= extent[3] - border) {
                int stop = m - 1;                       // python: ttemp = ttemp[0:m-1]
                if (stop < 0) stop = cur + stop;        // negative stop counts from the end
                if (stop < 0) stop = 0;
                if (stop < cur) cur = stop;
            }
        }
        if (cur < 2) continue;
        double* sg = samp + (size_t)gi * max_samp * 3;
        for (int m = 0; m < cur; ++m) {
            const double* c = dg + 3 * (m * keep_every);
            sg[3 * m] = c[0]; sg[3 * m + 1] = c[1]; sg[3 * m + 2] = c[2];
        }
        samp_count[gi] = cur;
        // true_path = ftemp[0 : ftemp.index(ttemp[-1]) + 1]: first dense configuration equal to the last kept one
        const double* last = sg + 3 * (cur - 1);
        int idx = (cur - 1) * keep_every;
        for (int i = 0; i < idx; ++i)
            if (dg[3 * i] == last[0] && dg[3 * i + 1] == last[1] && dg[3 * i + 2] == last[2]) { idx = i; break; }
        dense_count[gi] = idx + 1;
    }
    return IPP_OK;
}
