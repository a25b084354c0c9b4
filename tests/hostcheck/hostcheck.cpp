// hostcheck.cpp -- TEST INFRASTRUCTURE.  Compiles kessler_b200/csrc/sgp4_device.cuh
// (the very header the CUDA kernels instantiate) for the HOST with g++, so the
// restructured SGP4 algebra can be compared with the oracle on a machine without
// a GPU (tests/test_hostcheck.py).  Never linked into libkessler_b200.so.
#include <cstdint>
#include <vector>
#include "../../kessler_b200/csrc/sgp4_device.cuh"
#include "../../kessler_b200/csrc/screen_device.cuh"
#include "../../kessler_b200/csrc/kepler_device.cuh"

extern "C" {

int hc_sgp4_init(const double *elems, int64_t n, double *consts, int32_t *err) {
    for (int64_t i = 0; i < n; ++i) {
        double el[7];
        for (int k = 0; k < 7; ++k) el[k] = elems[k * n + i];
        int e = ksl::sgp4_init(el, [&](int k, double v) { consts[(int64_t)k * n + i] = v; });
        if (err) err[i] = e;
    }
    return 0;
}

int hc_sgp4_init_fast(const double *elems, int64_t n, double *consts, int32_t *err) {
    for (int64_t i = 0; i < n; ++i) {
        double el[7];
        for (int k = 0; k < 7; ++k) el[k] = elems[k * n + i];
        int e = ksl::sgp4_init_fast(el, [&](int k, double v) { consts[(int64_t)k * n + i] = v; }, ksl::kGridSinCos);
        if (err) err[i] = e;
    }
    return 0;
}

// what the Monte Carlo kernels run per sample: fast sgp4init + fast state; ok[i] = 0 where the sample is deferred to the exact route
int hc_init_state_fast(const double *elems, int64_t n, double tsince, double *rv, int32_t *err, int32_t *ok) {
    for (int64_t i = 0; i < n; ++i) {
        double el[7];
        for (int k = 0; k < 7; ++k) el[k] = elems[k * n + i];
        int e = 0;
        ok[i] = ksl::sgp4_init_state_fast(el, tsince, rv + 6 * i, rv + 6 * i + 3, ksl::kGridSinCos, &e) ? 1 : 0;
        err[i] = e;
    }
    return 0;
}

int hc_sgp4_prop(const double *consts, int64_t n, const double *tsince, int64_t m, int per_sample_t, int want_vel,
                 double *rv, int32_t *err) {
    for (int64_t i = 0; i < n; ++i) {
        int e = 0;
        auto ld = [&](int k) { return consts[(int64_t)k * n + i]; };
        for (int64_t j = 0; j < m; ++j) {
            double t = per_sample_t ? tsince[i * m + j] : tsince[j];
            double *o = rv + (i * m + j) * 6;
            if (want_vel) e |= ksl::sgp4_prop<true>(ld, t, o, o + 3);
            else { e |= ksl::sgp4_prop<false>(ld, t, o, o + 3); o[3] = o[4] = o[5] = 0.0; }
        }
        if (err) err[i] |= e;
    }
    return 0;
}

// full state through sgp4_state (fast path with velocity over a FastView of the 36 constants, exact fallback)
int hc_sgp4_state(const double *consts, int64_t n, const double *tsince, int64_t m, double *rv, int32_t *err, int32_t *fast_used) {
    for (int64_t i = 0; i < n; ++i) {
        int e = 0;
        auto ld = [&](int k) { return consts[(int64_t)k * n + i]; };
        for (int64_t j = 0; j < m; ++j) {
            double *o = rv + (i * m + j) * 6;
            double tr[3], tv[3];
            const ksl::FastView<decltype(ld)> view{ld, ld(KSL_C_ISIMP) != 0.0};
            fast_used[i * m + j] = ksl::sgp4_prop_fast_rv<true>(view, tsince[j], tr, tv, ksl::kGridSinCos) ? 1 : 0;
            e |= ksl::sgp4_state(ld, tsince[j], o, o + 3, ksl::kGridSinCos);
        }
        if (err) err[i] |= e;
    }
    return 0;
}

// position through the fast path; fast_used[i*m + j] = 1 when the branch-free path accepted the epoch
int hc_sgp4_position(const double *consts, int64_t n, const double *tsince, int64_t m, double *pos, int32_t *err,
                     int32_t *fast_used) {
    for (int64_t i = 0; i < n; ++i) {
        double cm[ksl::KSL_NFAST];
        for (int k = 0; k < KSL_NCONST; ++k) cm[k] = consts[(int64_t)k * n + i];
        ksl::prepare_fast_record(cm, 1);
        auto ld = [&](int k) { return cm[k]; };
        int e = 0;
        for (int64_t j = 0; j < m; ++j) {
            double *o = pos + (i * m + j) * 3;
            double tmp[3];
            fast_used[i * m + j] = ksl::sgp4_prop_fast(ld, tsince[j], tmp, ksl::kGridSinCos) ? 1 : 0;
            e |= ksl::sgp4_position(cm, 1, tsince[j], o, ksl::kGridSinCos);
        }
        err[i] |= e;
    }
    return 0;
}

// the same with the near-circular form where the record qualifies (as k_screen_warp decides it: once per record, from a
// bound over |tsince| <= tmax); small_used[i] = 1 where it was selected
int hc_sgp4_position_small(const double *consts, int64_t n, const double *tsince, int64_t m, double tmax, double *pos,
                           int32_t *fast_used, int32_t *small_used) {
    for (int64_t i = 0; i < n; ++i) {
        double cm[ksl::KSL_NFAST];
        for (int k = 0; k < KSL_NCONST; ++k) cm[k] = consts[(int64_t)k * n + i];
        ksl::prepare_fast_record(cm, 1);
        auto ld = [&](int k) { return cm[k]; };
        const bool small = ksl::fast_record_is_small_e(ld, tmax);
        small_used[i] = small ? 1 : 0;
        for (int64_t j = 0; j < m; ++j) {
            double *o = pos + (i * m + j) * 3;
            const bool ok = small ? ksl::sgp4_prop_fast<true>(ld, tsince[j], o, ksl::kGridSinCos) : ksl::sgp4_prop_fast<false>(ld, tsince[j], o, ksl::kGridSinCos);
            fast_used[i * m + j] = ok ? 1 : 0;
            if (!ok) { double v[3]; ksl::sgp4_prop<false>(ld, tsince[j], o, v); }
        }
    }
    return 0;
}

int hc_upsample_index(int64_t j, int64_t n_in, int64_t n_out, int64_t *i0, int64_t *i1, double *w0, double *w1) {
    double scale = n_out > 1 ? (double)(n_in - 1) / (double)(n_out - 1) : 0.0;
    ksl::upsample_index(j, scale, n_in, i0, i1, w0, w1);
    return 0;
}

double hc_wrap_twopi(double x) { return ksl::wrap_twopi(x); }

int hc_kepler_partials(const double *states, int64_t n, double mu, double *elements, double *jac) {
    for (int64_t i = 0; i < n; ++i)
        if (!ksl::kepler_elements_and_partials(states + 6 * i, mu, elements + 6 * i, jac + 36 * i)) return (int)(i + 1);
    return 0;
}

// Host emulation of k_screen's data flow (kessler_b200.cu): kThreads "threads" walk the
// coarse epochs in chunks of kThreads with a one-epoch overlap, each keeping its own
// running Best (which is what the analytic pruning sees), merged at the end.
static int s_use_table = 1;
int hc_set_use_table(int v) { s_use_table = v; return 0; }
static int s_use_f32 = 1;
int hc_set_use_f32(int v) { s_use_f32 = v; return 0; }

// the two ways of selecting the near-circular form: is_small[i * m + j] = fast_record_is_small_e(record i, tmax[j]),
// t_limit[i] = fast_record_small_e_tmax(record i)
int hc_small_e_selection(const double *consts, int64_t n, const double *tmax, int64_t m, int32_t *is_small, double *t_limit) {
    for (int64_t i = 0; i < n; ++i) {
        double cm[ksl::KSL_NFAST];
        for (int k = 0; k < KSL_NCONST; ++k) cm[k] = consts[(int64_t)k * n + i];
        ksl::prepare_fast_record(cm, 1);
        auto ld = [&](int k) { return cm[k]; };
        t_limit[i] = ksl::fast_record_small_e_tmax(ld);
        for (int64_t j = 0; j < m; ++j) is_small[i * m + j] = ksl::fast_record_is_small_e(ld, tmax[j]) ? 1 : 0;
    }
    return 0;
}

// FP32 skip test vs the FP64 one on n segments (d0, d1: [n][3] km; cur: [n] m^2).  Returns the number of segments the
// FP32 test would skip although the FP64 test -- with its own slack -- keeps them (must be 0), and the two skip counts.
int64_t hc_skip_compare(const double *d0, const double *d1, const double *cur, int64_t n, double thr2, int conj_open,
                        int64_t *skipped32, int64_t *skipped64) {
    int64_t bad = 0, s32 = 0, s64 = 0;
    for (int64_t i = 0; i < n; ++i) {
        const double *a = d0 + 3 * i, *b = d1 + 3 * i;
        const float f0[3] = {(float)a[0], (float)a[1], (float)a[2]}, f1[3] = {(float)b[0], (float)b[1], (float)b[2]};
        const bool k32 = ksl::segment_skippable_f32(f0, f1, ksl::segment_want_f32(cur[i], conj_open != 0, thr2));
        const bool k64 = ksl::segment_skippable(a, b, cur[i], conj_open != 0, thr2);
        s32 += k32; s64 += k64;
        if (k32 && !k64) ++bad;
    }
    *skipped32 = s32; *skipped64 = s64;
    return bad;
}

// segment_bounds_f32 (the lazy search's deferral bounds) on n segments against the exact quadratic in FP64: whenever it
// returns true, lb must not exceed the quadratic's minimum over [0, 1], ub must not fall below it, and both must be finite.
// Returns the number of violations; *bounded = how many segments it bounded at all.
int64_t hc_bounds_check(const double *d0, const double *d1, int64_t n, double dl, int64_t *bounded) {
    int64_t bad = 0, nb = 0;
    for (int64_t i = 0; i < n; ++i) {
        const double *a = d0 + 3 * i, *b = d1 + 3 * i;
        const float f0[3] = {(float)a[0], (float)a[1], (float)a[2]}, f1[3] = {(float)b[0], (float)b[1], (float)b[2]};
        float lb, ub;
        if (!ksl::segment_bounds_f32(f0, f1, (float)dl, &lb, &ub)) continue;
        ++nb;
        const double bx = b[0] - a[0], by = b[1] - a[1], bz = b[2] - a[2];
        const double a0 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2], a1 = a[0] * bx + a[1] * by + a[2] * bz, a2 = bx * bx + by * by + bz * bz;
        double l = a2 > 0.0 ? -a1 / a2 : 0.0;
        l = l < 0.0 ? 0.0 : (l > 1.0 ? 1.0 : l);
        const double qmin = a0 + 2.0 * a1 * l + a2 * l * l;
        if (!(std::fabs((double)lb) < INFINITY) || !(std::fabs((double)ub) < INFINITY) || (double)lb > qmin || (double)ub < qmin) ++bad;
    }
    *bounded = nb;
    return bad;
}

int hc_screen(const double *consts_t, const double *epoch_t, int64_t n_t, const double *consts_c, const double *epoch_c,
              int64_t n, const double *times_coarse, int64_t n_coarse, int64_t n_fine, double thr, int mode,
              int64_t *i_min, double *d_min, int64_t *i_conj, double *d_conj, int32_t *err) {
    const int kThreads = 256;
    const double scale = n_fine > 1 ? (double)(n_coarse - 1) / (double)(n_fine - 1) : 0.0;
    const double thr2 = thr * thr;
    const long long last = n_coarse - 1;
    std::vector<double> pt(3 * n_coarse), pc(3 * n_coarse);
    std::vector<int> ja(n_coarse + 1);
    for (int64_t k = 0; k <= n_coarse; ++k) ja[k] = (k == n_coarse) ? (int)n_fine : (int)ksl::first_fine_of(k, scale, n_coarse, n_fine);
    const int *ja_tab = (s_use_table ? ja.data() : nullptr);
    for (int64_t s = 0; s < n; ++s) {
        const int64_t it = (n_t == 1) ? 0 : s;
        int et = 0, ec = 0;
        double cmt[ksl::KSL_NFAST], cmc[ksl::KSL_NFAST];
        for (int q = 0; q < KSL_NCONST; ++q) { cmt[q] = consts_t[(int64_t)q * n_t + it]; cmc[q] = consts_c[(int64_t)q * n + s]; }
        ksl::prepare_fast_record(cmt, 1);
        ksl::prepare_fast_record(cmc, 1);
        for (int64_t k = 0; k < n_coarse; ++k) {
            et |= ksl::sgp4_position(cmt, 1, (times_coarse[k] - epoch_t[it]) * 1440.0, &pt[3 * k], ksl::kGridSinCos);
            ec |= ksl::sgp4_position(cmc, 1, (times_coarse[k] - epoch_c[s]) * 1440.0, &pc[3 * k], ksl::kGridSinCos);
        }
        std::vector<ksl::Best> best(kThreads);
        for (auto &b : best) ksl::best_init(b);
        double cutoff = INFINITY;   // as k_screen_warp: the group's running minimum, refreshed after a chunk that visited
        for (long long base = 0; base < last || base == 0; base += kThreads - 1) {
            bool visited = false;
            for (int tid = 0; tid < kThreads; ++tid) {
                const long long k = base + tid;
                const bool owner = (k < last && tid < kThreads - 1) || (k == last);
                if (!owner) continue;
                const long long nb = (k == last) ? k : k + 1;
                const double d0[3] = {pt[3 * k] - pc[3 * k], pt[3 * k + 1] - pc[3 * k + 1], pt[3 * k + 2] - pc[3 * k + 2]};
                const double d1[3] = {pt[3 * nb] - pc[3 * nb], pt[3 * nb + 1] - pc[3 * nb + 1], pt[3 * nb + 2] - pc[3 * nb + 2]};
                // the warp kernel's FP32 form of the test (the FP64 form serves the CTA kernel and the tail segment)
                const float f0[3] = {(float)d0[0], (float)d0[1], (float)d0[2]}, f1[3] = {(float)d1[0], (float)d1[1], (float)d1[2]};
                if (mode == 0 && (s_use_f32 ? ksl::segment_skippable_f32(f0, f1, ksl::segment_want_f32(cutoff, best[tid].conj_j == ksl::kNoIndex, thr2))
                                            : ksl::segment_skippable(d0, d1, cutoff, best[tid].conj_j == ksl::kNoIndex, thr2)))
                    continue;
                visited = true;
                if (mode == 0)
                    ksl::segment_search<0>(k, last, scale, n_coarse, n_fine, thr2, &pt[3 * k], &pt[3 * nb], &pc[3 * k], &pc[3 * nb], best[tid], ja_tab);
                else
                    ksl::segment_search<1>(k, last, scale, n_coarse, n_fine, thr2, &pt[3 * k], &pt[3 * nb], &pc[3 * k], &pc[3 * nb], best[tid], ja_tab);
            }
            if (visited)
                for (int tid = 0; tid < kThreads; ++tid) cutoff = fmin(cutoff, best[tid].sq);
        }
        ksl::Best b = best[0];
        for (int tid = 1; tid < kThreads; ++tid) ksl::best_merge(b, best[tid]);
        long long im, ic;
        ksl::best_result(b, &im, d_min + s, &ic, d_conj + s);
        i_min[s] = im;
        i_conj[s] = ic;
        err[s] = et | (ec << KSL_ERR_CHASER_SHIFT);
    }
    return 0;
}
}
