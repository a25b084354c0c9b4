// screen_device.cuh -- fine-grid closest-approach search on one coarse segment.
//
// Replaces util.upsample + model.find_conjunction as Conjunction.forward() uses
// them (/root/reference/kessler/util.py:541-555, model.py:23-52, 611-613): the
// reference interpolates both trajectories linearly from n_coarse to n_fine points
// (torch F.interpolate, align_corners=True) and scans all n_fine squared distances
// three times (argmin, threshold mask, nonzero).  Here the fine grid is never
// materialised: every coarse segment [k, k+1] owns the fine points whose torch
// source index floors to k, and on a segment the squared distance is a quadratic
// in the interpolation weight, so the arg-min and the first threshold crossing are
// located in closed form and only their neighbouring fine points are evaluated --
// with the reference's own arithmetic (lerp in km, *1e3, difference, sum of
// squares, no contraction), so the indices and distances are the reference's.
//
// __host__ __device__ for the same reason as sgp4_device.cuh (tests/hostcheck).
#pragma once
#include "sgp4_device.cuh"

namespace ksl {

constexpr long long kNoIndex = 0x7fffffffffffffffLL;

// arg-min / first-below bookkeeping.  torch semantics (model.py:23-52): first
// index wins ties, NaN counts as the minimum (torch.argmin), NaN < thr is False.
struct Best {
    double sq;         // smallest non-NaN squared distance seen
    long long j;       // its index
    long long nan_j;   // smallest index whose squared distance is NaN (kNoIndex: none)
    long long conj_j;  // smallest index below the threshold (kNoIndex: none)
    double conj_sq;
};

KSL_HD void best_init(Best &b) {
    b.sq = INFINITY;
    b.j = kNoIndex;
    b.nan_j = kNoIndex;
    b.conj_j = kNoIndex;
    b.conj_sq = 0.0;
}

KSL_HD void best_visit(Best &b, double sq, long long j, double thr2) {
    if (sq != sq) {
        if (j < b.nan_j) b.nan_j = j;
        return;
    }
    if (sq < b.sq || (sq == b.sq && j < b.j)) {
        b.sq = sq;
        b.j = j;
    }
    if (sq < thr2 && j < b.conj_j) {
        b.conj_j = j;
        b.conj_sq = sq;
    }
}

KSL_HD void best_merge(Best &b, const Best &o) {
    if (o.sq < b.sq || (o.sq == b.sq && o.j < b.j)) {
        b.sq = o.sq;
        b.j = o.j;
    }
    if (o.nan_j < b.nan_j) b.nan_j = o.nan_j;
    if (o.conj_j < b.conj_j) {
        b.conj_j = o.conj_j;
        b.conj_sq = o.conj_sq;
    }
}

// (i_min, d_min, i_conj | -1, d_conj | NaN) as find_conjunction returns them
KSL_HD void best_result(const Best &b, long long *i_min, double *d_min, long long *i_conj, double *d_conj) {
    if (b.nan_j != kNoIndex) {
        *i_min = b.nan_j;
        *d_min = NAN;
    } else {
        *i_min = (b.j == kNoIndex) ? 0 : b.j;
        *d_min = sqrt(b.sq);
    }
    if (b.conj_j != kNoIndex) {
        *i_conj = b.conj_j;
        *d_conj = sqrt(b.conj_sq);
    } else {
        *i_conj = -1;
        *d_conj = NAN;
    }
}

// torch's source index for fine point j: float32 floor of the FP64 coordinate
KSL_HD long long up_i0(long long j, double scale, long long n_in) {
    long long a = (long long)floorf((float)((double)j * scale));
    return a > n_in - 1 ? n_in - 1 : a;
}

// first fine index whose source index is >= k (the source index is monotone in j)
KSL_HD long long first_fine_of(long long k, double scale, long long n_in, long long n_out) {
    if (k <= 0) return 0;
    if (!(scale > 0.0)) return n_out;
    const double g = ceil((double)k / scale);
    long long j = g >= (double)n_out ? n_out : (long long)g;
    if (j < 0) j = 0;
    while (j > 0 && up_i0(j - 1, scale, n_in) >= k) --j;
    while (j < n_out && up_i0(j, scale, n_in) < k) ++j;
    return j;
}

// squared distance at fine point j of segment k, in the reference's arithmetic.
// t0/t1 (c0/c1): target (chaser) position in km at coarse epochs k and min(k+1, last).
KSL_HD double seg_sq(long long j, long long k, double scale, const double *t0, const double *t1, const double *c0,
                     const double *c1) {
    double l = add_rn(mul_rn((double)j, scale), -(double)k);   // torch rounds the product first: no contraction
    l = l < 0.0 ? 0.0 : l;
    l = l > 1.0 ? 1.0 : l;
    const double w0 = 1.0 - l;
    const double dx = add_rn(lerp_m(w0, t0[0], l, t1[0]), -lerp_m(w0, c0[0], l, c1[0]));
    const double dy = add_rn(lerp_m(w0, t0[1], l, t1[1]), -lerp_m(w0, c0[1], l, c1[1]));
    const double dz = add_rn(lerp_m(w0, t0[2], l, t1[2]), -lerp_m(w0, c0[2], l, c1[2]));
    return sqdist3(dx, dy, dz);
}

// Cheap, division-free test (most segments end here): with d0, d1 the relative positions (target -
// chaser, km) at the two ends, q(l) = |d0 + l (d1 - d0)|^2 on l in [0, 1] bounds every fine point of the
// segment from below.  True = that bound can neither beat the running minimum nor (while no crossing has
// been found) the threshold, so no fine point of the segment matters.  NaN / inf never skip (an infinite
// coordinate can still produce NaN fine points through 0 * inf).
// `cur`: the running minimum (m^2) to beat -- the caller's own, or any smaller one known to belong to the same trace
// (the warp kernel shares one across its lanes); `conj_open`: no threshold crossing found yet by this caller.
KSL_HD bool segment_skippable(const double d0[3], const double d1[3], double cur, bool conj_open, double thr2) {
    const double bx = d1[0] - d0[0], by = d1[1] - d0[1], bz = d1[2] - d0[2];
    const double a0 = fma_rn(d0[0], d0[0], fma_rn(d0[1], d0[1], d0[2] * d0[2]));   // km^2
    const double a1 = fma_rn(d0[0], bx, fma_rn(d0[1], by, d0[2] * bz));
    const double a2 = fma_rn(bx, bx, fma_rn(by, by, bz * bz));
    // cur is +inf until the first visit.  Rounding of the exact evaluation vs. the closed form: << 1 m^2 + 1e-9 relative
    const double want = (fmax(cur, conj_open ? thr2 : 0.0) * (1.0 + 1e-9) + 1.0) * 1e-6 + 1e-9 * a0;
    const double q1 = a0 + (a1 + a1) + a2;
    const bool rising = a1 >= 0.0, falling = (a1 + a2) <= 0.0;
    const bool out = rising ? (a0 > want) : (falling ? (q1 > want) : ((a0 - want) * a2 > a1 * a1));
    return out && (fabs(a0) < INFINITY) && (fabs(a2) < INFINITY);
}

// The same test in FP32, for the warp kernel's inner loop (the FP64 pipe is what that kernel is bound by; the FP32
// pipe idles).  d0, d1: the relative positions rounded to float (km); want_km2 = segment_want(...) rounded to float.
// Conservative: FP32 evaluation perturbs q(l) on [0, 1] by at most ~1e-6 (a0 + a2), and the margin added to the
// target is 1e-4 (a0 + a2) -- a segment whose bound lies within that margin of the running minimum is simply visited.
// Overflow / NaN give false.  (hostcheck: never true where the FP64 test says false with margin.)
KSL_HD bool segment_skippable_f32(const float d0[3], const float d1[3], float want_km2) {
    const float bx = d1[0] - d0[0], by = d1[1] - d0[1], bz = d1[2] - d0[2];
    const float a0 = fmaf(d0[0], d0[0], fmaf(d0[1], d0[1], d0[2] * d0[2]));
    const float a1 = fmaf(d0[0], bx, fmaf(d0[1], by, d0[2] * bz));
    const float a2 = fmaf(bx, bx, fmaf(by, by, bz * bz));
    const float want = fmaf(1e-4f, a0 + a2, want_km2);
    const float q1 = a0 + (a1 + a1) + a2;
    const bool rising = a1 >= 0.0f, falling = (a1 + a2) <= 0.0f;
    const bool out = rising ? (a0 > want) : (falling ? (q1 > want) : ((a0 - want) * a2 > a1 * a1));
    return out && (a0 < INFINITY) && (a2 < INFINITY);
}
// Bounds of the squared distance over the fine points of one segment, in FP32 (km^2), for the lazy search of the warp
// kernel: lb <= every fine point of the segment (the quadratic's minimum over [0, 1] minus the FP32 margin), and
// ub >= AT LEAST ONE fine point of the segment: the quadratic at most `dl` away from its minimiser on [0, 1], where dl
// (in units of the interpolation weight) is a distance within which the segment is guaranteed to hold a fine point
// (the callers pass 2 fine steps: torch's float32 floor can move a boundary point into the neighbouring segment), plus
// the margin.  q(l* + d) - q(l*) = q'(l*) d + a2 d^2 with q'(l*) = 0 at an interior minimiser, 2 a1 at l* = 0, 2 (a1 + a2)
// at l* = 1.  Returns false when the segment is not finite (then nothing is bounded).
KSL_HD bool segment_bounds_f32(const float d0[3], const float d1[3], float dl, float *lb, float *ub) {
    const float bx = d1[0] - d0[0], by = d1[1] - d0[1], bz = d1[2] - d0[2];
    const float a0 = fmaf(d0[0], d0[0], fmaf(d0[1], d0[1], d0[2] * d0[2]));
    const float a1 = fmaf(d0[0], bx, fmaf(d0[1], by, d0[2] * bz));
    const float a2 = fmaf(bx, bx, fmaf(by, by, bz * bz));
    const float margin = 1e-4f * (a0 + a2);
    float qmin, slope;
    if (a1 >= 0.0f) {                      // rising: minimum at the segment's start
        qmin = a0;
        slope = a1 + a1;
    } else if (a1 + a2 <= 0.0f) {          // falling: at its end
        qmin = a0 + (a1 + a1) + a2;
        slope = -2.0f * (a1 + a2);
    } else {                               // interior vertex (a2 > -a1 > 0)
        // (a1 / a2) a1, not a1 a1 / a2: the ratio lies in (-1, 0), while the square overflows once the separation passes
        // 1e9.6 km -- the garbage of a flagged trace, but garbage the enumeration orders all the same; an overflow here used
        // to give qmin = ub = -inf, a "bound" that pruned the rest of the trace (found by benchmarks/fuzz_screen.py)
        qmin = a0 - (a1 / a2) * a1;
        slope = 0.0f;
    }
    *lb = qmin - margin;
    *ub = qmin + fmaf(slope, dl, a2 * dl * dl) + margin;
    return (a0 < INFINITY) && (a2 < INFINITY) && (fabsf(*lb) < INFINITY) && (fabsf(*ub) < INFINITY);
}
// the target of the FP32 test: what the bound has to exceed, km^2 (cur in m^2; +inf until the first visit)
KSL_HD float segment_want_f32(double cur, bool conj_open, double thr2) {
    return (float)((fmax(cur, conj_open ? thr2 : 0.0) * (1.0 + 1e-9) + 1.0) * 1e-6);
}

// Visit segment k (coarse epochs k and k+1; k == last is torch's clamped tail
// segment with both ends at `last`).  kMode = KSL_SCREEN_BRUTE enumerates every
// fine point; KSL_SCREEN_ANALYTIC evaluates only the candidates that can hold the
// segment's minimum or first threshold crossing, and only if the segment can
// improve on `best` at all.  Callers normally run segment_skippable first.
// ja_tab (optional): ja_tab[k] = first_fine_of(k) for k = 0..n_coarse-1 and ja_tab[n_coarse] = n_fine,
// precomputed once per CTA in shared memory; NULL = evaluate first_fine_of on the fly.
template <int kMode>
KSL_HD void segment_search(long long k, long long last, double scale, long long n_coarse, long long n_fine, double thr2,
                           const double *t0, const double *t1, const double *c0, const double *c1, Best &best,
                           const int *ja_tab = nullptr) {
    long long ja, jb;
    if (ja_tab) {
        ja = ja_tab[k];
        jb = (long long)ja_tab[k + 1] - 1;
    } else {
        ja = first_fine_of(k, scale, n_coarse, n_fine);
        jb = ((k == last) ? n_fine : first_fine_of(k + 1, scale, n_coarse, n_fine)) - 1;
    }
    if (ja > jb) return;
    bool brute = (kMode == KSL_SCREEN_BRUTE);
    if (!brute) {
        // D(l) = A + l*B in metres; q(l) = a0 + 2 a1 l + a2 l^2
        const double ax = (t0[0] - c0[0]) * 1e3, ay = (t0[1] - c0[1]) * 1e3, az = (t0[2] - c0[2]) * 1e3;
        const double bx = (t1[0] - c1[0]) * 1e3 - ax, by = (t1[1] - c1[1]) * 1e3 - ay, bz = (t1[2] - c1[2]) * 1e3 - az;
        const double a0 = ax * ax + ay * ay + az * az;
        const double a1 = ax * bx + ay * by + az * bz;
        const double a2 = bx * bx + by * by + bz * bz;
        if (!(fabs(a0) < INFINITY) || !(fabs(a2) < INFINITY)) {
            brute = true;  // NaN / inf states: reference semantics by enumeration
        } else {
            const double cur = best.sq;
            const double slack = 1.0 + 1e-9 * fmax(a0, cur < INFINITY ? cur : 0.0);
            const double la = fmin(fmax(add_rn(mul_rn((double)ja, scale), -(double)k), 0.0), 1.0);
            const double lb = fmin(fmax(add_rn(mul_rn((double)jb, scale), -(double)k), 0.0), 1.0);
            const double lv = (a2 > 0.0) ? fmin(fmax(-a1 / a2, la), lb) : la;
            const double qv = a0 + lv * (2.0 * a1 + a2 * lv);  // analytic minimum over the segment
            const bool may_min = !(qv > cur + slack);
            const bool may_conj = (qv < thr2 + slack) && (ja < best.conj_j);
            if (may_min || may_conj) {
                if (!(a2 * scale * scale > 1e-4)) {
                    brute = true;  // < 1 cm of relative motion per fine step: enumerate
                } else {
                    const double inv_scale = 1.0 / scale;
                    long long jc = (long long)floor((lv + (double)k) * inv_scale);
                    jc = jc < ja ? ja : (jc > jb ? jb : jc);
                    for (long long j = (jc - 1 < ja ? ja : jc - 1); j <= jc + 2 && j <= jb; ++j)
                        best_visit(best, seg_sq(j, k, scale, t0, t1, c0, c1), j, thr2);
                    if (may_conj) {
                        // first crossing of q(l) = thr2 on the way down
                        const double disc = a1 * a1 - a2 * (a0 - thr2);
                        long long jr = ja;
                        if (disc > 0.0) {
                            const double lr = (-a1 - sqrt(disc)) / a2;
                            if (lr > la) {
                                const double g = ceil((lr + (double)k) * inv_scale);
                                jr = g > (double)jb ? jb : (long long)g;
                                if (jr < ja) jr = ja;
                            }
                        }
                        for (long long j = (jr - 2 < ja ? ja : jr - 2); j <= jr + 2 && j <= jb; ++j)
                            best_visit(best, seg_sq(j, k, scale, t0, t1, c0, c1), j, thr2);
                        best_visit(best, seg_sq(ja, k, scale, t0, t1, c0, c1), ja, thr2);
                    }
                }
            }
        }
    }
    if (brute)
        for (long long j = ja; j <= jb; ++j) best_visit(best, seg_sq(j, k, scale, t0, t1, c0, c1), j, thr2);
}

}  // namespace ksl
