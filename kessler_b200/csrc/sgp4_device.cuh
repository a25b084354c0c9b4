// sgp4_device.cuh -- FP64 near-earth SGP4 for sm_100a, written for the B200 FP64 pipe.
//
// Replaces dsgp4.initialize_tle / dsgp4.propagate / dsgp4.propagate_batch as Kessler
// calls them (/root/reference/kessler/model.py:581,592,536,541; util.py:570-576).
// Algorithm: Vallado et al. AIAA 2006-6753 near-earth branch, WGS-84, floor-mod
// wrapping, as dsgp4 computes it.  This is NOT a
// transcription of that code: the propagation is restructured for the GPU so that
// one evaluation needs 4 full-range sincos instead of 16 sin/cos + atan2 + pow:
//   * sin(mm) for the drag term, sin/cos of su, xnode, xinc and the Kepler iterates
//     after the first are obtained by rotating an already-known (sin, cos) pair
//     through a SMALL angle with a short Taylor kernel (angle-addition), with a
//     full sincos fallback when the angle is not small;
//   * atan2 is never taken: (sinu, cosu) is already the unit vector of u;
//   * am^1.5 is am*sqrt(am); x/xke is x*(1/xke);
//   * the secular angles are evaluated WITHOUT fma contraction so that their
//     rounding matches the reference's mul-then-add (they carry |tsince| ~ 5e4 min
//     of phase, where half an ulp is ~1e-9 km).
// All of it is exact algebra; results agree with the oracle to ~1e-11 km.
//
// The functions are __host__ __device__ so tests/hostcheck can compile this very
// file with g++ and compare it with the oracle on the CPU (test infrastructure;
// the shipped library only ever instantiates the device side).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/kessler_b200.h"

#if defined(__CUDACC__)
#define KSL_HD __host__ __device__ __forceinline__
#else
#define KSL_HD inline
#endif

namespace ksl {

constexpr double kMu = 398600.5;
constexpr double kRe = 6378.137;
constexpr double kJ2 = 0.00108262998905;
constexpr double kJ3 = -0.00000253215306;
constexpr double kJ4 = -0.00000161098761;
constexpr double kJ3oJ2 = kJ3 / kJ2;
constexpr double kTwoPi = 6.283185307179586;
constexpr double kInvTwoPi = 0.15915494309189535;
constexpr double kX2o3 = 2.0 / 3.0;
// xke = 60/sqrt(re^3/mu), vkmpersec = re*xke/60: values of the FP64 expressions
// dsgp4 evaluates (checked against the pickled _xke by tests/test_oracle_golden.py)
constexpr double kXke = 0.07436685316871385;
constexpr double kInvXke = 1.0 / kXke;
constexpr double kVkmPerSec = kRe * kXke / 60.0;

// ---- arithmetic with pinned rounding -------------------------------------------
#if defined(__CUDA_ARCH__)
KSL_HD double mul_rn(double a, double b) { return __dmul_rn(a, b); }
KSL_HD double add_rn(double a, double b) { return __dadd_rn(a, b); }
KSL_HD double fma_rn(double a, double b, double c) { return __fma_rn(a, b, c); }
#else
KSL_HD double mul_rn(double a, double b) { volatile double p = a * b; return p; }
KSL_HD double add_rn(double a, double b) { volatile double s = a + b; return s; }
KSL_HD double fma_rn(double a, double b, double c) { return fma(a, b, c); }
#endif
// a*b + c with the product rounded first (what `a*b + c` means in NumPy/torch)
KSL_HD double mad_2r(double a, double b, double c) { return add_rn(mul_rn(a, b), c); }

// Python/torch `x % 2pi` for finite x: the exact remainder with the sign of 2pi.
// floor(x * (1/2pi)) can be off by one next to a multiple of 2pi; fma(-q, 2pi, x) is
// exact (the remainder is representable), so one conditional step repairs it.  A tiny
// negative x gives 2pi - |x| which ROUNDS to 2pi, exactly as fmod-based `%` does.
KSL_HD double wrap_twopi(double x) {
    const double q = floor(x * kInvTwoPi);
    double r = fma_rn(-q, kTwoPi, x);
    if (r < 0.0) r += kTwoPi;                                    // q one too large
    else if (r > kTwoPi || (r == kTwoPi && x >= 0.0)) r -= kTwoPi;  // q one too small
    return r;
}

KSL_HD void sincos_full(double x, double *s, double *c) {
#if defined(__CUDA_ARCH__)
    sincos(x, s, c);
#else
    *s = sin(x);
    *c = cos(x);
#endif
}

// sin/cos of a small angle, |d| <= kSmallAngle: truncation error < d^11/11! (2.5e-19 at 0.1)
constexpr double kSmallAngle = 0.1;
KSL_HD void sincos_small(double d, double *s, double *c) {
    const double d2 = d * d;
    double ps = fma_rn(d2, 2.7557319223985893e-06, -1.9841269841269841e-04);
    ps = fma_rn(ps, d2, 8.3333333333333332e-03);
    ps = fma_rn(ps, d2, -1.6666666666666666e-01);
    *s = fma_rn(ps * d2, d, d);
    double pc = fma_rn(d2, -2.7557319223985888e-07, 2.4801587301587302e-05);
    pc = fma_rn(pc, d2, -1.3888888888888889e-03);
    pc = fma_rn(pc, d2, 4.1666666666666664e-02);
    pc = fma_rn(pc, d2, -0.5);
    *c = fma_rn(pc, d2, 1.0);
}

// (s, c) of angle a  ->  (s, c) of angle a + d
KSL_HD void rotate_by(double d, double a_plus_d, double *s, double *c) {
    if (fabs(d) <= kSmallAngle) {
        double sd, cd;
        sincos_small(d, &sd, &cd);
        const double s0 = *s, c0 = *c;
        *s = fma_rn(s0, cd, c0 * sd);
        *c = fma_rn(c0, cd, -(s0 * sd));
    } else {
        sincos_full(a_plus_d, s, c);
    }
}

// ---- tables, bit helpers, sincos on a grid ---------------------------------------------
// Tables are deliberately NOT const on the device: ptxas folds const initialisers into 64-bit immediates that cost two
// UMOV issue slots per use; a mutable __constant__ array is read with uniform LDCU loads.
#if defined(__CUDACC__)
#define KSL_TABLE __constant__ double
#else
#define KSL_TABLE static const double
#endif
KSL_HD int lo_word(double t) {
#if defined(__CUDA_ARCH__)
    return __double2loint(t);
#else
    long long b;
    memcpy(&b, &t, 8);
    return (int)(b & 0xffffffffLL);
#endif
}
KSL_HD int hi_word(double t) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(t);
#else
    long long b;
    memcpy(&b, &t, 8);
    return (int)(b >> 32);
#endif
}
// |x| < bound for bound = 2^e, decided on the high word: an integer compare on the ALU pipe instead of a DSETP on
// the FP64 pipe this code is bound by.  NaN and inf give false.
KSL_HD bool abs_below_pow2(double x, int e) { return (hi_word(x) & 0x7fffffff) < ((1023 + e) << 20); }
// ---- sincos on a 512-point grid ---------------------------------------------------
// x = k (2 pi / 512) + r, |r| <= pi / 512: (sin, cos)(k 2 pi / 512) from an 8 KB table (correctly rounded; the screen
// kernels stage it in shared memory, `grid` points there), two-term kernels for r (sin through r^5, cos through r^4:
// truncation 7e-20 / 8e-17) and the angle-addition formulas in the form S + (S (cos r - 1) + C sin r), which keeps the
// result within an ulp.  14 FP64 instructions instead of 20-21 for the fdlibm kernels on [-pi/4, pi/4], and no
// quadrant bookkeeping at all.  Valid for |x| < 2^18: the step is split into 28 + 53 bits, so fn * step_hi is exact.
//   [0] 512 / (2 pi)  [1] 1.5 * 2^52  [2] step_hi  [3] step_lo  [4..5] -1/6, 1/120  [6..7] -1/2, 1/24
#include "sincos_grid_table.cuh"   // kGridPoints, kGridTab, kGridSinCos
// kSinTerms / kCosTerms (2 or 1): terms of the kernels for sin r - r and cos r - 1.  The short forms are for callers whose
// use of the result tolerates their truncation: sin through r^3 leaves r^5 / 120 < 7e-14, cos through r^2 leaves r^4 / 24 < 6e-11.
template <int kSinTerms = 2, int kCosTerms = 2>
KSL_HD void sincos_grid(double x, const double *grid, double &s, double &c) {
    const double t = fma_rn(x, kGridTab[0], kGridTab[1]);
    const int k = lo_word(t) & (kGridPoints - 1);
    const double fn = t - kGridTab[1];
    const double r = fma_rn(-fn, kGridTab[3], fma_rn(-fn, kGridTab[2], x));
    const double z = r * r;
    const double sr = fma_rn(z * (kSinTerms == 2 ? fma_rn(z, kGridTab[5], kGridTab[4]) : kGridTab[4]), r, r);   // sin r
    const double cm1 = z * (kCosTerms == 2 ? fma_rn(z, kGridTab[7], kGridTab[6]) : kGridTab[6]);               // cos r - 1
#if defined(__CUDA_ARCH__)
    const double2 sc = *reinterpret_cast<const double2 *>(grid + 2 * k);
    const double S = sc.x, C = sc.y;
#else
    const double S = grid[2 * k], C = grid[2 * k + 1];
#endif
    s = fma_rn(C, sr, fma_rn(S, cm1, S));
    c = fma_rn(-S, sr, fma_rn(C, cm1, C));
}

// the grid point nearest to x: (S, C) = (sin, cos)(x - r) straight from the table, |r| <= pi / 512
KSL_HD double grid_point(double x, const double *grid, double &S, double &C) {
    const double t = fma_rn(x, kGridTab[0], kGridTab[1]);
    const int k = lo_word(t) & (kGridPoints - 1);
    const double fn = t - kGridTab[1];
#if defined(__CUDA_ARCH__)
    const double2 sc = *reinterpret_cast<const double2 *>(grid + 2 * k);
    S = sc.x;
    C = sc.y;
#else
    S = grid[2 * k];
    C = grid[2 * k + 1];
#endif
    return fma_rn(-fn, kGridTab[3], fma_rn(-fn, kGridTab[2], x));
}


// ---- sgp4init -----------------------------------------------------------------
// el = {no_kozai, ecco, inclo, nodeo, argpo, mo, bstar}; c[KSL_NCONST] with stride
// `cs` between consecutive constants (1 for a private record, n for SoA).
template <typename Store>
KSL_HD int sgp4_init(const double el[7], Store &&put) {
    const double no_kozai = el[0], ecco = el[1], inclo = el[2], nodeo = el[3], argpo = el[4], mo = el[5], bstar = el[6];
    const double eccsq = ecco * ecco, omeosq = 1.0 - eccsq, rteosq = sqrt(omeosq);
    double sinio, cosio;
    sincos_full(inclo, &sinio, &cosio);
    const double cosio2 = cosio * cosio;
    const double ak = pow(kXke / no_kozai, kX2o3);
    const double d1 = 0.75 * kJ2 * (3.0 * cosio2 - 1.0) / (rteosq * omeosq);
    double del = d1 / (ak * ak);
    const double adel = ak * (1.0 - del * del - del * (1.0 / 3.0 + 134.0 * del * del / 81.0));
    del = d1 / (adel * adel);
    const double no = no_kozai / (1.0 + del);
    const double ao = pow(kXke / no, kX2o3);
    const double po = ao * omeosq;
    const double con42 = 1.0 - 5.0 * cosio2;
    const double con41 = -con42 - cosio2 - cosio2;
    const double posq = po * po;
    const double rp = ao * (1.0 - ecco);
    const bool isimp = rp < (220.0 / kRe + 1.0);
    double sfour = 78.0 / kRe + 1.0;
    const double qz0 = (120.0 - 78.0) / kRe;
    double qzms24 = qz0 * qz0 * qz0 * qz0;
    const double perige = (rp - 1.0) * kRe;
    if (perige < 156.0) {
        sfour = perige - 78.0;
        if (perige < 98.0) sfour = 20.0;
        const double q = (120.0 - sfour) / kRe;
        qzms24 = q * q * q * q;
        sfour = sfour / kRe + 1.0;
    }
    const double pinvsq = 1.0 / posq;
    const double tsi = 1.0 / (ao - sfour);
    const double eta = ao * ecco * tsi, etasq = eta * eta, eeta = ecco * eta;
    const double psisq = fabs(1.0 - etasq);
    const double tsi2 = tsi * tsi;
    const double coef = qzms24 * (tsi2 * tsi2);
    const double coef1 = coef / (psisq * psisq * psisq * sqrt(psisq));
    const double cc2 = coef1 * no * (ao * (1.0 + 1.5 * etasq + eeta * (4.0 + etasq)) +
                                     0.375 * kJ2 * tsi / psisq * con41 * (8.0 + 3.0 * etasq * (8.0 + etasq)));
    const double cc1 = bstar * cc2;
    double cc3 = 0.0;
    if (ecco > 1.0e-4) cc3 = -2.0 * coef * tsi * kJ3oJ2 * no * sinio / ecco;
    const double x1mth2 = 1.0 - cosio2;
    double s2w, c2w;
    sincos_full(2.0 * argpo, &s2w, &c2w);
    (void)s2w;
    const double cc4 = 2.0 * no * coef1 * ao * omeosq *
                       (eta * (2.0 + 0.5 * etasq) + ecco * (0.5 + 2.0 * etasq) -
                        kJ2 * tsi / (ao * psisq) *
                            (-3.0 * con41 * (1.0 - 2.0 * eeta + etasq * (1.5 - 0.5 * eeta)) +
                             0.75 * x1mth2 * (2.0 * etasq - eeta * (1.0 + etasq)) * c2w));
    const double cc5 = 2.0 * coef1 * ao * omeosq * (1.0 + 2.75 * (etasq + eeta) + eeta * etasq);
    const double cosio4 = cosio2 * cosio2;
    const double temp1 = 1.5 * kJ2 * pinvsq * no;
    const double temp2 = 0.5 * temp1 * kJ2 * pinvsq;
    const double temp3 = -0.46875 * kJ4 * pinvsq * pinvsq * no;
    const double mdot = no + 0.5 * temp1 * rteosq * con41 + 0.0625 * temp2 * rteosq * (13.0 - 78.0 * cosio2 + 137.0 * cosio4);
    const double argpdot = -0.5 * temp1 * con42 + 0.0625 * temp2 * (7.0 - 114.0 * cosio2 + 395.0 * cosio4) +
                           temp3 * (3.0 - 36.0 * cosio2 + 49.0 * cosio4);
    const double xhdot1 = -temp1 * cosio;
    const double nodedot = xhdot1 + (0.5 * temp2 * (4.0 - 19.0 * cosio2) + 2.0 * temp3 * (3.0 - 7.0 * cosio2)) * cosio;
    double sw, cw;
    sincos_full(argpo, &sw, &cw);
    (void)sw;
    const double omgcof = bstar * cc3 * cw;
    double xmcof = 0.0;
    if (ecco > 1.0e-4) xmcof = -kX2o3 * coef * bstar / eeta;
    const double nodecf = 3.5 * omeosq * xhdot1 * cc1;
    const double t2cof = 1.5 * cc1;
    const double xl_den = (fabs(cosio + 1.0) > 1.5e-12) ? (1.0 + cosio) : 1.5e-12;
    const double xlcof = -0.25 * kJ3oJ2 * sinio * (3.0 + 5.0 * cosio) / xl_den;
    const double aycof = -0.5 * kJ3oJ2 * sinio;
    double smo, cmo;
    sincos_full(mo, &smo, &cmo);
    const double delmotemp = 1.0 + eta * cmo;
    const double delmo = delmotemp * delmotemp * delmotemp;
    const double x7thm1 = 7.0 * cosio2 - 1.0;
    double d2 = 0, d3 = 0, d4 = 0, t3cof = 0, t4cof = 0, t5cof = 0;
    if (!isimp) {
        const double cc1sq = cc1 * cc1;
        d2 = 4.0 * ao * tsi * cc1sq;
        const double temp = d2 * tsi * cc1 / 3.0;
        d3 = (17.0 * ao + sfour) * temp;
        d4 = 0.5 * temp * ao * tsi * (221.0 * ao + 31.0 * sfour) * cc1;
        t3cof = d2 + 2.0 * cc1sq;
        t4cof = 0.25 * (3.0 * d3 + cc1 * (12.0 * d2 + 10.0 * cc1sq));
        t5cof = 0.2 * (3.0 * d4 + 12.0 * cc1 * d3 + 6.0 * d2 * d2 + 15.0 * cc1sq * (2.0 * d2 + cc1sq));
    }
    put(KSL_C_NO, no); put(KSL_C_ECCO, ecco); put(KSL_C_INCLO, inclo); put(KSL_C_NODEO, nodeo);
    put(KSL_C_ARGPO, argpo); put(KSL_C_MO, mo); put(KSL_C_BSTAR, bstar); put(KSL_C_MDOT, mdot);
    put(KSL_C_ARGPDOT, argpdot); put(KSL_C_NODEDOT, nodedot); put(KSL_C_NODECF, nodecf); put(KSL_C_CC1, cc1);
    put(KSL_C_CC4, cc4); put(KSL_C_CC5, cc5); put(KSL_C_T2COF, t2cof); put(KSL_C_OMGCOF, omgcof);
    put(KSL_C_XMCOF, xmcof); put(KSL_C_ETA, eta); put(KSL_C_DELMO, delmo); put(KSL_C_SINMAO, smo);
    put(KSL_C_D2, d2); put(KSL_C_D3, d3); put(KSL_C_D4, d4); put(KSL_C_T3COF, t3cof); put(KSL_C_T4COF, t4cof);
    put(KSL_C_T5COF, t5cof); put(KSL_C_AYCOF, aycof); put(KSL_C_XLCOF, xlcof); put(KSL_C_CON41, con41);
    put(KSL_C_X1MTH2, x1mth2); put(KSL_C_X7THM1, x7thm1); put(KSL_C_ISIMP, isimp ? 1.0 : 0.0);
    put(KSL_C_COSIO, cosio); put(KSL_C_SINIO, sinio); put(KSL_C_AOBASE, ao); put(KSL_C_PAD, 0.0);
    if (!(no > 0.0) || !(no < 1.0e300)) return KSL_ERR_DEEPSPACE;   // NaN / inf / non-positive
    return (kTwoPi / no >= 225.0) ? KSL_ERR_DEEPSPACE : 0;
}

// ---- sgp4 ---------------------------------------------------------------------
// Load: c(k) returns constant k of this record (register array, shared memory
// broadcast or an SoA global load).  r [km], v [km/s] (v skipped when !kVel).
template <bool kVel, typename Load>
KSL_HD int sgp4_prop(Load &&c, double t, double r[3], double v[3]) {
    int err = 0;
    // secular terms: product rounded, then the sum (reference rounding; see header)
    const double xmdf = mad_2r(c(KSL_C_MDOT), t, c(KSL_C_MO));
    const double argpdf = mad_2r(c(KSL_C_ARGPDOT), t, c(KSL_C_ARGPO));
    const double nodedf = mad_2r(c(KSL_C_NODEDOT), t, c(KSL_C_NODEO));
    const double t2 = t * t;
    double nodem = mad_2r(c(KSL_C_NODECF), t2, nodedf);
    const double cc1 = c(KSL_C_CC1), bstar = c(KSL_C_BSTAR);
    double tempa = 1.0 - cc1 * t;
    double tempe = (bstar * c(KSL_C_CC4)) * t;
    double templ = c(KSL_C_T2COF) * t2;
    double mm = xmdf, argpm = argpdf;
    if (c(KSL_C_ISIMP) == 0.0) {
        double sx, cx;
        sincos_full(xmdf, &sx, &cx);
        const double delomg = c(KSL_C_OMGCOF) * t;
        const double delmtemp = 1.0 + c(KSL_C_ETA) * cx;
        const double delm = c(KSL_C_XMCOF) * (delmtemp * delmtemp * delmtemp - c(KSL_C_DELMO));
        const double temp = delomg + delm;
        mm = add_rn(xmdf, temp);
        argpm = add_rn(argpdf, -temp);
        const double t3 = t2 * t, t4 = t3 * t;
        tempa = tempa - c(KSL_C_D2) * t2 - c(KSL_C_D3) * t3 - c(KSL_C_D4) * t4;
        rotate_by(temp, mm, &sx, &cx);   // sx = sin(mm)
        tempe = tempe + (bstar * c(KSL_C_CC5)) * (sx - c(KSL_C_SINMAO));
        templ = templ + c(KSL_C_T3COF) * t3 + t4 * (c(KSL_C_T4COF) + t * c(KSL_C_T5COF));
    }
    const double no = c(KSL_C_NO);
    const double am = c(KSL_C_AOBASE) * tempa * tempa;
    double em = c(KSL_C_ECCO) - tempe;
    if (em >= 1.0 || em < -0.001) err |= KSL_ERR_ECC;
    if (em < 1.0e-6) em = 1.0e-6;
    mm = mad_2r(no, templ, mm);
    const double xlm = add_rn(add_rn(mm, argpm), nodem);
    nodem = wrap_twopi(nodem);
    argpm = wrap_twopi(argpm);
    const double xlmw = wrap_twopi(xlm);
    mm = wrap_twopi(add_rn(add_rn(xlmw, -argpm), -nodem));

    double sarg, carg;
    sincos_full(argpm, &sarg, &carg);
    const double axnl = em * carg;
    double temp = 1.0 / (am * (1.0 - em * em));
    const double aynl = em * sarg + temp * c(KSL_C_AYCOF);
    const double xl = add_rn(add_rn(add_rn(mm, argpm), nodem), temp * c(KSL_C_XLCOF) * axnl);
    const double u = wrap_twopi(add_rn(xl, -nodem));

    // Kepler: u = eo1 - axnl sin(eo1) + aynl cos(eo1); Newton with the +-0.95 clamp.
    // (se, ce) always belongs to the CURRENT eo1: a vectorised dsgp4 call iterates
    // every lane until its slowest lane converged, which leaves converged lanes with
    // sin/cos consistent with eo1; an exhausted lane keeps the pair one step behind.
    double eo1 = u, se, ce;
    sincos_full(u, &se, &ce);
    double se_prev = se, ce_prev = ce;
    bool converged = false;
#pragma unroll 1
    for (int ktr = 1; ktr <= 10; ++ktr) {
        const double den = 1.0 - ce * axnl - se * aynl;
        double d = (u - aynl * ce + axnl * se - eo1) / den;
        d = (d >= 0.95) ? 0.95 : d;
        d = (d <= -0.95) ? -0.95 : d;
        eo1 += d;
        se_prev = se;
        ce_prev = ce;
        rotate_by(d, eo1, &se, &ce);
        if (fabs(d) < 1.0e-12) { converged = true; break; }
    }
    if (!converged) { se = se_prev; ce = ce_prev; }

    const double ecose = axnl * ce + aynl * se;
    const double esine = axnl * se - aynl * ce;
    const double el2 = axnl * axnl + aynl * aynl;
    const double pl = am * (1.0 - el2);
    if (pl < 0.0) err |= KSL_ERR_SEMILATUS;
    const double rl = am * (1.0 - ecose);
    const double betal = sqrt(1.0 - el2);
    temp = esine / (1.0 + betal);
    const double amorl = am / rl;
    double sinu = amorl * (se - aynl - axnl * temp);
    double cosu = amorl * (ce - axnl + aynl * temp);
    const double sin2u = (cosu + cosu) * sinu;
    const double cos2u = 1.0 - 2.0 * sinu * sinu;
    const double ipl = 1.0 / pl;
    const double temp1 = 0.5 * kJ2 * ipl;
    const double temp2 = temp1 * ipl;
    const double con41 = c(KSL_C_CON41), x1mth2 = c(KSL_C_X1MTH2);
    const double mrt = rl * (1.0 - 1.5 * temp2 * betal * con41) + 0.5 * temp1 * x1mth2 * cos2u;
    const double cosio = c(KSL_C_COSIO), sinio = c(KSL_C_SINIO);
    const double dsu = -0.25 * temp2 * c(KSL_C_X7THM1) * sin2u;   // su    = atan2(sinu,cosu) + dsu
    const double dnode = 1.5 * temp2 * cosio * sin2u;             // xnode = nodem + dnode
    const double dinc = 1.5 * temp2 * cosio * sinio * cos2u;      // xinc  = inclo + dinc

    double sinsu, cossu, snod, cnod, sini, cosi;
    if (fabs(dsu) <= kSmallAngle && fabs(dnode) <= kSmallAngle && fabs(dinc) <= kSmallAngle) {
        // (sinu, cosu) is the unit vector of u up to rounding; renormalise to first order
        const double h = fma_rn(-0.5, fma_rn(sinu, sinu, cosu * cosu) - 1.0, 1.0);
        sinsu = sinu * h;
        cossu = cosu * h;
        rotate_by(dsu, 0.0, &sinsu, &cossu);
        sincos_full(nodem, &snod, &cnod);
        rotate_by(dnode, 0.0, &snod, &cnod);
        sini = sinio;
        cosi = cosio;
        rotate_by(dinc, 0.0, &sini, &cosi);
    } else {   // far from any physical orbit (pl << 1): take the long way round
        const double su = atan2(sinu, cosu) + dsu;
        sincos_full(su, &sinsu, &cossu);
        sincos_full(nodem + dnode, &snod, &cnod);
        sincos_full(c(KSL_C_INCLO) + dinc, &sini, &cosi);
    }
    const double xmx = -snod * cosi, xmy = cnod * cosi;
    const double ux = xmx * sinsu + cnod * cossu;
    const double uy = xmy * sinsu + snod * cossu;
    const double uz = sini * sinsu;
    const double mr = mrt * kRe;
    r[0] = mr * ux;
    r[1] = mr * uy;
    r[2] = mr * uz;
    if (mrt < 1.0) err |= KSL_ERR_DECAYED;
    if (kVel) {
        const double sqrt_am = sqrt(am);
        const double nm = kXke / (am * sqrt_am);
        if (nm <= 0.0) err |= KSL_ERR_MEANMOTION;
        const double irl = 1.0 / rl;
        const double rdotl = sqrt_am * esine * irl;
        const double rvdotl = sqrt(pl) * irl;
        const double mvt = rdotl - nm * temp1 * x1mth2 * sin2u * kInvXke;
        const double rvdot = rvdotl + nm * temp1 * (x1mth2 * cos2u + 1.5 * con41) * kInvXke;
        const double vx = xmx * cossu - cnod * sinsu;
        const double vy = xmy * cossu - snod * sinsu;
        const double vz = sini * cossu;
        v[0] = (mvt * ux + rvdot * vx) * kVkmPerSec;
        v[1] = (mvt * uy + rvdot * vy) * kVkmPerSec;
        v[2] = (mvt * uz + rvdot * vz) * kVkmPerSec;
    }
    return err;
}

// ---- sgp4, fast path ------------------------------------------------------------
// Same algebra as sgp4_prop, specialised to what the Monte Carlo path overwhelmingly
// sees (eccentricity <~ 0.1, J2 short-period angles << 1 rad) and written to keep the
// warp scheduler's issue slots for the FP64 pipe (ncu, round 1: 518 FP64 instructions
// per evaluation but 830 others -- per-call slow-path guards of the library division /
// sincos, six small-angle fallbacks, 64-bit literals materialised through UMOV pairs):
//   * no data-dependent branch: every assumption is accumulated into one flag and the
//     caller re-runs the exact sgp4_prop when it is violated (`return false`);
//   * sin / cos from a 512-point table in shared memory + two-term kernels (sincos_grid), no quadrant logic;
//   * reciprocals from rcp.approx + one Newton or one cubic step, sized to the precision each use needs
//     (J2 / long-period terms carry 1e-3 factors), no IEEE-division slow path;
//   * Kepler's equation in three sweeps from the table point nearest to u (Halley, Newton, one more step with
//     the same reciprocal); same root as the reference's loop (the last step is checked);
//   * sqrt(1 - el2) from its series (el2 = e^2 <~ 1e-2);
//   * polynomial coefficients live in constant memory.
// Position only (the closest-approach search never reads the velocity).
// Tables are deliberately NOT const on the device: ptxas folds const initialisers into 64-bit
// immediates that cost two UMOV issue slots per use; a mutable __constant__ array is read with
// uniform LDCU loads (one slot per 64 or 128 bits) that the compiler can also keep resident.
// sin(d) = d + d^3 (S0 + S1 d^2 + S2 d^4 + S3 d^6);  cos(d) = 1 + d^2 (C0 + C1 d^2 + ... + C4 d^8)
KSL_TABLE kTrigTab[10] = {-1.6666666666666666e-01, 8.3333333333333332e-03, -1.9841269841269841e-04, 2.7557319223985893e-06,
                          -0.5, 4.1666666666666664e-02, -1.3888888888888889e-03, 2.4801587301587302e-05,
                          -2.7557319223985888e-07, 0.0};
// sqrt(1 - x) = 1 - x/2 - x^2/8 - x^3/16 - 5x^4/128 - 7x^5/256 - 21x^6/1024 - 33x^7/2048
KSL_TABLE kSqrtTab[8] = {-0.5, -0.125, -0.0625, -0.0390625, -0.02734375, -0.0205078125, -0.01611328125, 0.0};
KSL_HD double rcp_seed(double b) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    return r;
#else
    // the hardware seed works on the high words only: same accuracy class (~2^-20) for the host build
    long long bits;
    memcpy(&bits, &b, 8);
    bits &= ~0xffffffffLL;
    double bt;
    memcpy(&bt, &bits, 8);
    double r = 1.0 / bt;
    memcpy(&bits, &r, 8);
    bits &= ~0xffffffffLL;
    memcpy(&r, &bits, 8);
    return r;
#endif
}
KSL_HD double rsqrt_seed(double b) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    return r;
#else
    long long bits;   // high words only, as the hardware seed (see rcp_seed)
    memcpy(&bits, &b, 8);
    bits &= ~0xffffffffLL;
    double bt;
    memcpy(&bt, &bits, 8);
    double r = 1.0 / sqrt(bt);
    memcpy(&bits, &r, 8);
    bits &= ~0xffffffffLL;
    memcpy(&r, &bits, 8);
    return r;
#endif
}
// reciprocal good to ~1e-12 relative (one Newton step on the ~2^-20 seed)
KSL_HD double rcp_1(double b) {
    double r = rcp_seed(b);
    return fma_rn(r, fma_rn(-b, r, 1.0), r);
}
// |d| <= 0.1: 11th-order kernel (truncation 2.5e-19); rotate (s, c) by d
KSL_HD void rotate_small(double d, double &s, double &c) {
    const double d2 = d * d;
    double ps = fma_rn(d2, kTrigTab[3], kTrigTab[2]);
    ps = fma_rn(ps, d2, kTrigTab[1]);
    ps = fma_rn(ps, d2, kTrigTab[0]);
    const double sd = fma_rn(ps * d2, d, d);
    double pc = fma_rn(d2, kTrigTab[8], kTrigTab[7]);
    pc = fma_rn(pc, d2, kTrigTab[6]);
    pc = fma_rn(pc, d2, kTrigTab[5]);
    pc = fma_rn(pc, d2, kTrigTab[4]);
    const double cd = fma_rn(pc, d2, 1.0);
    const double s0 = s, c0 = c;
    s = fma_rn(s0, cd, c0 * sd);
    c = fma_rn(c0, cd, -(s0 * sd));
}
constexpr double kFastMaxEl2 = 0.0144;   // e <= 0.12: bound behind the series for sqrt(1 - el2) and the Kepler step sizes

// |d| <= 2e-3 (the three J2 short-period angles; bounded through 1/pl^2 below): 3rd / 4th-order kernels,
// truncation d^5/120 < 3e-16 (1e-17 at the usual 1e-3) and d^6/720 < 1e-19
constexpr double kMicroAngle = 2.0e-3;
KSL_HD void rotate_micro(double d, double &s, double &c) {
    const double d2 = d * d;
    const double sd = fma_rn(d2 * kTrigTab[0], d, d);
    const double cd = fma_rn(fma_rn(d2, kTrigTab[5], kTrigTab[4]), d2, 1.0);
    const double s0 = s, c0 = c;
    s = fma_rn(s0, cd, c0 * sd);
    c = fma_rn(c0, cd, -(s0 * sd));
}

// x - 2pi_d rint(x / 2pi_d) for |x| < 1e5, where 2pi_d = 6.283185307179586 is the DOUBLE the reference's `%`
// divides by: the remainder is exact (2pi_d split into 33 + 20 bits, both products exact), i.e. what fmod
// leaves up to a multiple of 2pi_d, in [-pi, pi], without compare or select.
//   [0] 1/(2 pi)  [1] 1.5*2^52  [2] 2pi_d, 33 leading bits  [3] 2pi_d - [2]
KSL_TABLE kWrapTab[4] = {1.59154943091895345608e-01, 6755399441055744.0, 6.28318530693650245667e+00,
                         2.43083775330887874588e-10};
KSL_HD double wrap_pm_pi(double x) {
    const double t = fma_rn(x, kWrapTab[0], kWrapTab[1]);
    const double fn = t - kWrapTab[1];
    return fma_rn(-fn, kWrapTab[3], fma_rn(-fn, kWrapTab[2], x));
}

// ---- sgp4init, fast form (the Monte Carlo kernels: one record per sample, used once) -----------------------------
// Same algebra as sgp4_init with the library calls replaced by what the fast propagation path already uses:
// sin / cos from the 512-point grid (sincos_grid, one ulp), (xke / n)^(2/3) as cbrt(.)^2 (two ulp), reciprocals as
// rcp.approx + two Newton steps (one ulp) and divisions as multiplications by them.  The constants agree with
// sgp4_init's to a few ulp of each FACTOR (tests/test_hostcheck.py: the secular rates, which carry the phase, to
// 1e-15; positions after a day to 1e-9 km); k_sgp4_init and the screening kernels keep sgp4_init.
// Requires finite elements with |angles| < 2^18; anything else produces non-finite constants, which the fast
// propagation path refuses (the caller then takes the exact route: sgp4_init + sgp4_prop).
KSL_HD double rcp_2(double b) {
    double r = rcp_seed(b);
    r = fma_rn(r, fma_rn(-b, r, 1.0), r);
    return fma_rn(r, fma_rn(-b, r, 1.0), r);
}
// a / b to one ulp: quotient from the reciprocal, one residual correction
KSL_HD double div_2(double a, double b) {
    const double r = rcp_2(b);
    const double q = a * r;
    return fma_rn(fma_rn(-b, q, a), r, q);
}
template <typename Store>
KSL_HD int sgp4_init_fast(const double el[7], Store &&put, const double *grid) {
    const double no_kozai = el[0], ecco = el[1], inclo = el[2], nodeo = el[3], argpo = el[4], mo = el[5], bstar = el[6];
    const double eccsq = ecco * ecco, omeosq = 1.0 - eccsq, rteosq = sqrt(omeosq);
    double sinio, cosio;
    sincos_grid(inclo, grid, sinio, cosio);
    const double cosio2 = cosio * cosio;
    const double cb0 = cbrt(div_2(kXke, no_kozai));
    const double ak = cb0 * cb0;
    const double d1 = 0.75 * kJ2 * (3.0 * cosio2 - 1.0) * rcp_2(rteosq * omeosq);
    double del = d1 * rcp_2(ak * ak);
    const double adel = ak * (1.0 - del * del - del * (1.0 / 3.0 + (134.0 / 81.0) * del * del));
    del = d1 * rcp_2(adel * adel);
    const double no = div_2(no_kozai, 1.0 + del);
    const double cb1 = cbrt(div_2(kXke, no));
    const double ao = cb1 * cb1;
    const double po = ao * omeosq;
    const double con42 = 1.0 - 5.0 * cosio2;
    const double con41 = -con42 - cosio2 - cosio2;
    const double posq = po * po;
    const double rp = ao * (1.0 - ecco);
    const bool isimp = rp < (220.0 / kRe + 1.0);
    double sfour = 78.0 / kRe + 1.0;
    const double qz0 = (120.0 - 78.0) / kRe;
    double qzms24 = qz0 * qz0 * qz0 * qz0;
    const double perige = (rp - 1.0) * kRe;
    if (perige < 156.0) {
        sfour = perige - 78.0;
        if (perige < 98.0) sfour = 20.0;
        const double q = (120.0 - sfour) * (1.0 / kRe);
        qzms24 = q * q * q * q;
        sfour = sfour * (1.0 / kRe) + 1.0;
    }
    const double pinvsq = rcp_2(posq);
    const double tsi = rcp_2(ao - sfour);
    const double eta = ao * ecco * tsi, etasq = eta * eta, eeta = ecco * eta;
    const double psisq = fabs(1.0 - etasq);
    const double tsi2 = tsi * tsi;
    const double coef = qzms24 * (tsi2 * tsi2);
    const double ipsisq = rcp_2(psisq);
    const double coef1 = coef * (ipsisq * ipsisq * ipsisq) * sqrt(ipsisq);
    const double cc2 = coef1 * no * (ao * (1.0 + 1.5 * etasq + eeta * (4.0 + etasq)) +
                                     0.375 * kJ2 * tsi * ipsisq * con41 * (8.0 + 3.0 * etasq * (8.0 + etasq)));
    const double cc1 = bstar * cc2;
    const bool ecc_big = ecco > 1.0e-4;
    const double cc3 = ecc_big ? -2.0 * coef * tsi * kJ3oJ2 * no * sinio * rcp_2(ecco) : 0.0;
    const double x1mth2 = 1.0 - cosio2;
    double sw, cw;
    sincos_grid(argpo, grid, sw, cw);
    const double c2w = fma_rn(-2.0 * sw, sw, 1.0);          // cos(2 argpo)
    const double cc4 = 2.0 * no * coef1 * ao * omeosq *
                       (eta * (2.0 + 0.5 * etasq) + ecco * (0.5 + 2.0 * etasq) -
                        kJ2 * tsi * rcp_2(ao * psisq) *
                            (-3.0 * con41 * (1.0 - 2.0 * eeta + etasq * (1.5 - 0.5 * eeta)) +
                             0.75 * x1mth2 * (2.0 * etasq - eeta * (1.0 + etasq)) * c2w));
    const double cc5 = 2.0 * coef1 * ao * omeosq * (1.0 + 2.75 * (etasq + eeta) + eeta * etasq);
    const double cosio4 = cosio2 * cosio2;
    const double temp1 = 1.5 * kJ2 * pinvsq * no;
    const double temp2 = 0.5 * temp1 * kJ2 * pinvsq;
    const double temp3 = -0.46875 * kJ4 * pinvsq * pinvsq * no;
    const double mdot = no + 0.5 * temp1 * rteosq * con41 + 0.0625 * temp2 * rteosq * (13.0 - 78.0 * cosio2 + 137.0 * cosio4);
    const double argpdot = -0.5 * temp1 * con42 + 0.0625 * temp2 * (7.0 - 114.0 * cosio2 + 395.0 * cosio4) +
                           temp3 * (3.0 - 36.0 * cosio2 + 49.0 * cosio4);
    const double xhdot1 = -temp1 * cosio;
    const double nodedot = xhdot1 + (0.5 * temp2 * (4.0 - 19.0 * cosio2) + 2.0 * temp3 * (3.0 - 7.0 * cosio2)) * cosio;
    const double omgcof = bstar * cc3 * cw;
    const double xmcof = ecc_big ? -kX2o3 * coef * bstar * rcp_2(eeta) : 0.0;
    const double nodecf = 3.5 * omeosq * xhdot1 * cc1;
    const double t2cof = 1.5 * cc1;
    const double xl_den = (fabs(cosio + 1.0) > 1.5e-12) ? (1.0 + cosio) : 1.5e-12;
    const double xlcof = -0.25 * kJ3oJ2 * sinio * (3.0 + 5.0 * cosio) * rcp_2(xl_den);
    const double aycof = -0.5 * kJ3oJ2 * sinio;
    double smo, cmo;
    sincos_grid(mo, grid, smo, cmo);
    const double delmotemp = 1.0 + eta * cmo;
    const double delmo = delmotemp * delmotemp * delmotemp;
    const double x7thm1 = 7.0 * cosio2 - 1.0;
    double d2 = 0, d3 = 0, d4 = 0, t3cof = 0, t4cof = 0, t5cof = 0;
    if (!isimp) {
        const double cc1sq = cc1 * cc1;
        d2 = 4.0 * ao * tsi * cc1sq;
        const double temp = d2 * tsi * cc1 * (1.0 / 3.0);
        d3 = (17.0 * ao + sfour) * temp;
        d4 = 0.5 * temp * ao * tsi * (221.0 * ao + 31.0 * sfour) * cc1;
        t3cof = d2 + 2.0 * cc1sq;
        t4cof = 0.25 * (3.0 * d3 + cc1 * (12.0 * d2 + 10.0 * cc1sq));
        t5cof = 0.2 * (3.0 * d4 + 12.0 * cc1 * d3 + 6.0 * d2 * d2 + 15.0 * cc1sq * (2.0 * d2 + cc1sq));
    }
    put(KSL_C_NO, no); put(KSL_C_ECCO, ecco); put(KSL_C_INCLO, inclo); put(KSL_C_NODEO, nodeo);
    put(KSL_C_ARGPO, argpo); put(KSL_C_MO, mo); put(KSL_C_BSTAR, bstar); put(KSL_C_MDOT, mdot);
    put(KSL_C_ARGPDOT, argpdot); put(KSL_C_NODEDOT, nodedot); put(KSL_C_NODECF, nodecf); put(KSL_C_CC1, cc1);
    put(KSL_C_CC4, cc4); put(KSL_C_CC5, cc5); put(KSL_C_T2COF, t2cof); put(KSL_C_OMGCOF, omgcof);
    put(KSL_C_XMCOF, xmcof); put(KSL_C_ETA, eta); put(KSL_C_DELMO, delmo); put(KSL_C_SINMAO, smo);
    put(KSL_C_D2, d2); put(KSL_C_D3, d3); put(KSL_C_D4, d4); put(KSL_C_T3COF, t3cof); put(KSL_C_T4COF, t4cof);
    put(KSL_C_T5COF, t5cof); put(KSL_C_AYCOF, aycof); put(KSL_C_XLCOF, xlcof); put(KSL_C_CON41, con41);
    put(KSL_C_X1MTH2, x1mth2); put(KSL_C_X7THM1, x7thm1); put(KSL_C_ISIMP, isimp ? 1.0 : 0.0);
    put(KSL_C_COSIO, cosio); put(KSL_C_SINIO, sinio); put(KSL_C_AOBASE, ao); put(KSL_C_PAD, 0.0);
    if (!(no > 0.0) || !(no < 1.0e300)) return KSL_ERR_DEEPSPACE;   // NaN / inf / non-positive
    return (225.0 * no <= kTwoPi) ? KSL_ERR_DEEPSPACE : 0;
}

// The fast path reads a FAST RECORD: the 36 public constants followed by products sgp4_prop rebuilds at
// every call (prepare_fast_record fills them once per record, in the kernel's shared-memory copy).
enum {
    KSL_F_BC4 = KSL_NCONST,   // bstar * cc4
    KSL_F_BC5,                // bstar * cc5 (0 when isimp)
    KSL_F_KMRT,               // 3/4 J2 con41
    KSL_F_KX1,                // 1/4 J2 x1mth2
    KSL_F_KSU,                // -1/8 J2 x7thm1
    KSL_F_KNODE,              // 3/4 J2 cosio
    KSL_F_KINC,               // 3/4 J2 cosio sinio
    KSL_F_PAD,
    KSL_NFAST
};

// returns true when the result is the one sgp4_prop would give (to rounding: the angle bookkeeping below is
// algebraically the reference's with fewer intermediate roundings, ~1e-13 rad); false = assumptions violated,
// nothing usable written: call sgp4_prop.
// c: fast record (prepare_fast_record); grid: kGridSinCos or a copy of it; v (km/s) is written when kVel.
// small_e (a compile-time choice: the callers branch once per chunk on a flag that is uniform over the warp, so that each
// form stays one straight-line block the scheduler can interleave with the other object's) selects the near-circular form: with
// el2 = axnl^2 + aynl^2 < kSmallMaxEl2 (tested, like everything else, before the result is accepted) Kepler's equation
// needs two sweeps instead of three and the series in el2 two terms instead of five; same accuracy class.
constexpr double kSmallMaxEl2 = 1.0e-4;   // el < 0.01
template <bool kVel, bool small_e = false, typename Load>
KSL_HD bool sgp4_prop_fast_rv(Load &&c, double t, double r[3], double v[3], const double *grid) {
    // Angles that grow with t (mean anomaly, mean longitude: ~1e3 rad over the window) keep the reference's
    // own sequence of roundings -- each costs ~1e-13 rad = 1e-9 km; everything of magnitude <= 2 pi is free
    // to use FMAs (1e-16 rad).
    const double xmdf = mad_2r(c(KSL_C_MDOT), t, c(KSL_C_MO));
    const double argpdf = fma_rn(c(KSL_C_ARGPDOT), t, c(KSL_C_ARGPO));
    const double t2 = t * t;
    const double nodem = fma_rn(c(KSL_C_NODECF), t2, fma_rn(c(KSL_C_NODEDOT), t, c(KSL_C_NODEO)));
    // Drag / long-period block, executed unconditionally: for an `isimp` record (perigee < 220 km)
    // prepare_fast_record zeroes omgcof, xmcof and bstar*cc5, and sgp4init already left d2..d4,
    // t3cof..t5cof at 0, which turns every term below into an exact +0.
    // (sin, cos)(xmdf) only feed the drag terms: sin the bstar*cc5 term of the eccentricity (< 1e-5: the short kernel's
    // 7e-14 is 1e-18), cos the perigee / anomaly exchange delm, where the short kernel's 6e-11 is multiplied by
    // 3 xmcof eta <~ |delm| -- below 2^-4 in the near-circular form (4e-12 rad), up to 1 in the general one (two terms)
    double sx, cx;
    if (small_e) sincos_grid<1, 1>(xmdf, grid, sx, cx);
    else sincos_grid<1, 2>(xmdf, grid, sx, cx);
    const double delmtemp = fma_rn(c(KSL_C_ETA), cx, 1.0);
    const double delm = c(KSL_C_XMCOF) * (delmtemp * delmtemp * delmtemp - c(KSL_C_DELMO));
    const double temp0 = fma_rn(c(KSL_C_OMGCOF), t, delm);
    const double mm0 = add_rn(xmdf, temp0);
    const double argpm = argpdf - temp0;
    // tempa = 1 - cc1 t - d2 t^2 - d3 t^3 - d4 t^4;  templ = t2cof t^2 + t3cof t^3 + t4cof t^4 + t5cof t^5
    const double tempa =
        fma_rn(-t, fma_rn(t, fma_rn(t, fma_rn(t, c(KSL_C_D4), c(KSL_C_D3)), c(KSL_C_D2)), c(KSL_C_CC1)), 1.0);
    const double templ =
        t2 * fma_rn(t, fma_rn(t, fma_rn(t, c(KSL_C_T5COF), c(KSL_C_T4COF)), c(KSL_C_T3COF)), c(KSL_C_T2COF));
    // sin(mm0) = sin(xmdf + temp0), |temp0| < 2^-4, to 1e-8 (sin through d^3, cos through d^4): it only feeds the
    // bstar*cc5 term of the eccentricity, and bstar*cc5 < 1e-5 even for bstar = 1e-2
    // (general form: |temp0| < 1 -- a high-drag record weeks from its epoch --, sin through d^9, cos through d^10: 2e-9)
    const double q2 = temp0 * temp0;
    double sd0, cd0;
    if (small_e) {
        sd0 = fma_rn(q2 * kTrigTab[0], temp0, temp0);
        cd0 = fma_rn(fma_rn(q2, kTrigTab[5], kTrigTab[4]), q2, 1.0);
    } else {
        sd0 = fma_rn(fma_rn(fma_rn(fma_rn(q2, kTrigTab[3], kTrigTab[2]), q2, kTrigTab[1]), q2, kTrigTab[0]) * q2, temp0, temp0);
        cd0 = fma_rn(fma_rn(fma_rn(fma_rn(fma_rn(q2, kTrigTab[8], kTrigTab[7]), q2, kTrigTab[6]), q2, kTrigTab[5]), q2, kTrigTab[4]), q2, 1.0);
    }
    const double smm = fma_rn(sx, cd0, cx * sd0);
    const double tempe = fma_rn(c(KSL_F_BC5), smm - c(KSL_C_SINMAO), c(KSL_F_BC4) * t);
    const double am = c(KSL_C_AOBASE) * tempa * tempa;
    const double em_raw = c(KSL_C_ECCO) - tempe;
    // the reference clamps em at 1e-6 and raises error 1 below -0.001: the general form follows it (a near-circular,
    // high-drag record weeks from its epoch), the near-circular form simply declines below 1e-6 (one compare)
    const double em = small_e ? em_raw : fmax(em_raw, 1.0e-6);
    // validity: sincos_grid needs < 2^18; |argpm| < 32 and |nodem| < 64 keep |u| small and bound what reducing these
    // UNWRAPPED angles by 2 pi instead of the reference's 2pi_d costs (2.4e-16 rad per turn: < 3e-15); |temp0| within
    // the kernels above
    bool ok = abs_below_pow2(xmdf, 16) && abs_below_pow2(argpm, 5) && abs_below_pow2(nodem, 6) &&
              abs_below_pow2(temp0, small_e ? -4 : 0) && (em_raw >= (small_e ? 1.0e-6 : -1.0e-3));
    const double mm = mad_2r(c(KSL_C_NO), templ, mm0);
    const double xlm = add_rn(add_rn(mm, argpm), nodem);
    // The reference wraps nodem, argpm, xlm and mm = xlm - argpm - nodem to [0, 2 pi) with exact remainders
    // and rebuilds u = (mm + argpm + nodem + long-period term) - nodem from them: modulo 2pi_d that is
    // (nodem itself stays unwrapped: |nodem| < 64, where reducing by 2pi_d or by 2 pi differs by < 3e-15 rad).
    // (Dropping the wrap and letting the Kepler solver's grid reduction do it would save four instructions and cost two
    // roundings at the magnitude of xlm: the median distance to the oracle goes from 2e-12 km to 5e-10 km -- not taken.)
    const double lm = wrap_pm_pi(xlm) - nodem;

    double sarg, carg;
    // (the short sine would do in the general form as well -- em < 0.12 leaves 8e-15 -- but the kernels built with it ran
    // slower: K2 1.065 ms against 1.009 per 4.9e7 evaluations, instruction scheduling, measured twice)
    if (small_e) sincos_grid<1, 2>(argpm, grid, sarg, carg);    // enters multiplied by em < 0.01: 7e-14 there is 7e-16
    else sincos_grid(argpm, grid, sarg, carg);
    const double axnl = em * carg;
    double temp = rcp_1(am * fma_rn(-em, em, 1.0));            // * aycof / xlcof ~ 1e-3: 1e-12 is plenty
    const double aynl = fma_rn(em, sarg, temp * c(KSL_C_AYCOF));
    const double u = fma_rn(temp * c(KSL_C_XLCOF), axnl, lm);  // |u| < 4 pi

    // Kepler: u = E - axnl sin E + aynl cos E, solved for g = u - E in three sweeps, starting from the grid point E0
    // nearest to u: (sin, cos)(E0) come straight from the table and g0 = u - E0 is the reduction's remainder
    // (|g0| <= pi / 512).  With f = -(g + esine), f' = 1 - ecose, f'' = esine, the first step is Halley's,
    // d = dN (1 - dN q / 2) with the Newton step dN = (g + esine) / (1 - ecose) and q = esine / (1 - ecose): it leaves
    // 2e-5 at e = 0.12 (Newton: 9e-4),
    // the Newton step that follows 5e-11 (its reciprocal is good to 7e-6), and the third step, with the same
    // reciprocal, 1e-15.  No individual bound checks: with el2 < 0.0144 (tested below) |d1| < 0.15 (11th-order
    // rotation: 2e-17) and |d2| < 3e-5 (rotate_micro); anything else (non-finite input) fails the test on the last
    // step, |d3| < 2^-30 (then the remaining error is < 1e-5 |d3|).
    double se, ce;
    double g = grid_point(u, grid, se, ce);
    double den = fma_rn(-se, aynl, fma_rn(-ce, axnl, 1.0));
    double rden = rcp_seed(den);                                               // ~2^-20: plenty for a step that leaves 2e-5
    const double qn = fma_rn(axnl, se, -(aynl * ce)) * rden;                   // esine / (1 - ecose) = f'' / f'
    const double dn = fma_rn(g, rden, qn);
    double d = fma_rn(dn * qn * -0.5, dn, dn);
    g -= d;
    if (small_e) {
        // |E - E0| <= el + pi / 512 < 0.0162: Halley leaves (el / 6) 0.0162^3 < 8e-9, the Newton step below -- reciprocal
        // refined cubically, (el |d|)^3 < 5e-12 -- leaves (el / 2) (8e-9)^2 + 8e-9 x 5e-12 < 1e-19.
        // rotation by |d| < 0.0162 (typically 4e-3): sin through d^5 (d^7 / 7! < 6e-16, typically 3e-21), cos through d^4
        // (d^6 / 6! < 2.5e-14, typically 6e-18: 2e-10 km at the very edge of the form)
        const double d2 = d * d;
        const double ps = fma_rn(d2, kTrigTab[1], kTrigTab[0]);
        const double sd = fma_rn(ps * d2, d, d);
        const double cd = fma_rn(fma_rn(d2, kTrigTab[5], kTrigTab[4]), d2, 1.0);
        const double s0 = se, c0 = ce;
        se = fma_rn(s0, cd, c0 * sd);
        ce = fma_rn(c0, cd, -(s0 * sd));
        den = fma_rn(-se, aynl, fma_rn(-ce, axnl, 1.0));
        const double er = fma_rn(-den, rden, 1.0);
        rden = fma_rn(rden, fma_rn(er, er, er), rden);
        d = fma_rn(axnl, se, fma_rn(-aynl, ce, g)) * rden;                     // |d| < 8e-9: first-order rotation, d^2 / 2 < 4e-17
    } else {
    rotate_small(d, se, ce);
    den = fma_rn(-se, aynl, fma_rn(-ce, axnl, 1.0));
    const double er = fma_rn(-den, rden, 1.0);                                 // |er| < 0.02: the step changed den by e d
    rden = fma_rn(rden, fma_rn(er, er, er), rden);                             // cubic: 1 / den to 7e-6
    d = fma_rn(axnl, se, fma_rn(-aynl, ce, g)) * rden;
    g -= d;
    rotate_micro(d, se, ce);
    d = fma_rn(axnl, se, fma_rn(-aynl, ce, g)) * rden;
    ok = ok && abs_below_pow2(d, -30);
    }
    const double s_prev = se;
    se = fma_rn(d, ce, se);                                                    // first-order rotation: d^2 / 2 < 5e-19
    ce = fma_rn(-d, s_prev, ce);

    const double ecose = fma_rn(axnl, ce, aynl * se);
    const double esine = fma_rn(axnl, se, -(aynl * ce));
    const double el2 = fma_rn(axnl, axnl, aynl * aynl);
    ok = ok && (el2 < (small_e ? kSmallMaxEl2 : kFastMaxEl2));
    const double omel2 = 1.0 - el2;
    const double pl = am * omel2;
    const double omecose = 1.0 - ecose;
    const double rl = am * omecose;
    double betal, amorl;
    if (small_e) {
        // sqrt(1 - x) through x^2 (x^3 / 16 < 7e-14) and 1 / (1 + sqrt(1 - x)) = 1/2 + x/8 + x^2/16 (5 x^3 / 128 < 4e-14):
        // both only enter multiplied by e^2 or J2
        betal = fma_rn(fma_rn(el2, kSqrtTab[1], kSqrtTab[0]), el2, 1.0);
        temp = esine * fma_rn(fma_rn(el2, 0.0625, 0.125), el2, 0.5);
        // rden = 1 / den one (tiny) step earlier: el |d| < 1e-10 off -> one Newton step
        amorl = fma_rn(rden, fma_rn(-omecose, rden, 1.0), rden);
    } else {
    double bs = fma_rn(el2, kSqrtTab[4], kSqrtTab[3]);   // through x^5: 2e-13 at el2 = 0.0144, and betal only enters
    bs = fma_rn(bs, el2, kSqrtTab[2]);                  // multiplied by e^2 or J2
    bs = fma_rn(bs, el2, kSqrtTab[1]);
    bs = fma_rn(bs, el2, kSqrtTab[0]);
    betal = fma_rn(bs, el2, 1.0);
    temp = esine * rcp_1(1.0 + betal);                                          // enters as e * temp ~ e^2
    // am / rl = 1 / (1 - ecose): rden is that reciprocal one sweep earlier, good to 7e-6 -> cubic refinement, 3e-16
    const double ea = fma_rn(-omecose, rden, 1.0);
    amorl = fma_rn(rden, fma_rn(ea, ea, ea), rden);
    }
    const double sinu = amorl * (se - aynl - axnl * temp);
    const double cosu = amorl * (ce - axnl + aynl * temp);
    const double sin2u = (cosu + cosu) * sinu;
    const double cos2u = fma_rn(-2.0 * sinu, sinu, 1.0);
    const double ipl = rcp_1(pl);                                               // only inside J2 terms (5e-4)
    const double ipl2 = ipl * ipl;
    // mrt = rl (1 - 3/2 temp2 betal con41) + 1/2 temp1 x1mth2 cos2u,  temp1 = J2/2/pl, temp2 = temp1/pl
    const double mrt = fma_rn(rl, fma_rn(-c(KSL_F_KMRT), ipl2 * betal, 1.0), c(KSL_F_KX1) * ipl * cos2u);
    const double ps2 = ipl2 * sin2u;
    const double dsu = c(KSL_F_KSU) * ps2;
    const double dnode = c(KSL_F_KNODE) * ps2;
    const double dinc = c(KSL_F_KINC) * (ipl2 * cos2u);
    // |dsu|, |dnode|, |dinc| <= 3/4 J2 / pl^2 (|x7thm1| <= 6, |sin|, |cos| <= 1): one test covers the three
    ok = ok && abs_below_pow2(ipl2, 1);   // 2 < kMicroAngle / (0.75 J2) = 2.47

    // (sinu, cosu) is a unit vector to rounding (the last, first-order rotation is by < 1e-9)
    double sinsu = sinu, cossu = cosu;
    if (small_e) {      // |dsu| <= 1/8 J2 |x7thm1| / pl^2 < 8.2e-4 (4e-4 at 40 deg): cos through d^2 leaves d^4 / 24 < 2e-14 (1e-15)
        const double q2 = dsu * dsu;
        const double sd = fma_rn(q2 * kTrigTab[0], dsu, dsu);
        const double cd = fma_rn(q2, kTrigTab[4], 1.0);
        sinsu = fma_rn(sinu, cd, cosu * sd);
        cossu = fma_rn(cosu, cd, -(sinu * sd));
    } else {
        rotate_micro(dsu, sinsu, cossu);
    }
    double snod, cnod;
    sincos_grid(nodem + dnode, grid, snod, cnod);   // xnode, as the reference forms it (full kernels: r^5 / 120 here is 5e-10 km)
    double sini = c(KSL_C_SINIO), cosi = c(KSL_C_COSIO);
    if (small_e) {      // pl > 0.9999 am >= ~1: |dinc| <= 3/8 J2 / pl^2 < 4.1e-4, cos through d^2 leaves d^4 / 24 < 1.2e-15
        const double q2 = dinc * dinc;
        const double sd = fma_rn(q2 * kTrigTab[0], dinc, dinc);
        const double cd = fma_rn(q2, kTrigTab[4], 1.0);
        const double s0 = sini, c0 = cosi;
        sini = fma_rn(s0, cd, c0 * sd);
        cosi = fma_rn(c0, cd, -(s0 * sd));
    } else {
        rotate_micro(dinc, sini, cosi);
    }
    const double xmx = -snod * cosi, xmy = cnod * cosi;     // (only the velocity needs these two)
    const double mr = mrt * kRe;
    const double rs = mr * sinsu, rcs = mr * cossu;
    const double rsc = rs * cosi;
    r[0] = fma_rn(-snod, rsc, cnod * rcs);
    r[1] = fma_rn(cnod, rsc, snod * rcs);
    r[2] = sini * rs;
    if (kVel) {
        // sqrt(am) / rl = (am / rl) am^-1/2 and sqrt(pl) / rl = sqrt(1 - el2) (am / rl) am^-1/2: one reciprocal square
        // root (seed + one cubic step: with e = 1 - am r^2, r (1 + e/2 + 3 e^2/8) is off by 5/16 e^3), no square root
        const double r0 = rsqrt_seed(am);
        const double er0 = fma_rn(-(am * r0), r0, 1.0);
        const double rsa = fma_rn(r0, fma_rn(0.375, er0, 0.5) * er0, r0);
        const double w = amorl * rsa;
        // sqrt(1 - el2), relative: through x^5 it leaves 21/1024 x^6 < 2e-13 at el2 = 0.0144 (1e-9 m/s, 1e-21 at e = 0.01);
        // the near-circular form stops at x^2 (x^3 / 16 < 7e-14)
        double bf;
        if (small_e) {
            bf = fma_rn(el2, kSqrtTab[1], kSqrtTab[0]);
        } else {
            bf = fma_rn(el2, kSqrtTab[4], kSqrtTab[3]);
            bf = fma_rn(bf, el2, kSqrtTab[2]);
            bf = fma_rn(bf, el2, kSqrtTab[1]);
            bf = fma_rn(bf, el2, kSqrtTab[0]);
        }
        const double rvdotl = w * fma_rn(bf, el2, 1.0);
        const double k1 = (0.5 * kJ2) * ipl * (rsa * rsa * rsa);           // temp1 nm / xke
        const double mvt = fma_rn(-k1 * c(KSL_C_X1MTH2), sin2u, w * esine);
        const double rvdot = fma_rn(k1, fma_rn(c(KSL_C_X1MTH2), cos2u, 1.5 * c(KSL_C_CON41)), rvdotl);
        const double ms = mvt * kVkmPerSec, rv = rvdot * kVkmPerSec;
        const double ux = fma_rn(xmx, sinsu, cnod * cossu), uy = fma_rn(xmy, sinsu, snod * cossu), uz = sini * sinsu;
        v[0] = fma_rn(ms, ux, rv * fma_rn(xmx, cossu, -(cnod * sinsu)));
        v[1] = fma_rn(ms, uy, rv * fma_rn(xmy, cossu, -(snod * sinsu)));
        v[2] = fma_rn(ms, uz, rv * (sini * cossu));
        ok = ok && abs_below_pow2(w, 10);   // am = 0 or non-finite
    }
    // pl < 0 is impossible with el2 < 0.0144; decay (mrt < 1) and non-finite values go to the exact path
    return ok && (mrt >= 1.0) && abs_below_pow2(mr, 996);
}
template <bool kSmallE = false, typename Load>
KSL_HD bool sgp4_prop_fast(Load &&c, double t, double r[3], const double *grid) {
    double v[3];
    return sgp4_prop_fast_rv<false, kSmallE>(c, t, r, v, grid);
}
// Is the near-circular form worth selecting for this record over |tsince| <= tmax?  A bound on el over the window:
// the eccentricity and what drag can add to it, plus the long-period term of aynl (|aycof| / (a (1 - e^2)) < 1.3e-3).
// Only a performance hint -- a wrong guess costs exact-path evaluations (el2 is tested per evaluation), never accuracy.
template <typename Load>
KSL_HD bool fast_record_is_small_e(Load &&c, double tmax) {
    const double drag_e = fabs(c(KSL_F_BC4)) * tmax + 2.0 * fabs(c(KSL_F_BC5));
    const double bound = c(KSL_C_ECCO) + drag_e + 1.5e-3;
    // ... and from below: that form declines em < 1e-6 (the general one clamps there, as the reference does), so a record
    // that is circular to the last digit of its TLE would take the exact path at every epoch
    const bool above_clamp = c(KSL_C_ECCO) - drag_e > 2.0e-6;
    // ... and the drag angle temp0 = omgcof t + xmcof ((1 + eta cos M)^3 - delmo) must stay inside that form's short kernels
    const double ae = 1.0 + fabs(c(KSL_C_ETA));
    const double drag_angle = fabs(c(KSL_C_OMGCOF)) * tmax + fabs(c(KSL_C_XMCOF)) * (ae * ae * ae + fabs(c(KSL_C_DELMO)));
    return bound < 0.0095 && drag_angle < 0.06 && above_clamp;
}

// The same selection solved for tmax: the near-circular form is worth selecting for |tsince| < the value returned (every
// condition of fast_record_is_small_e is monotone in tmax); <= 0: never.  For callers that decide per tile of epochs (K2).
template <typename Load>
KSL_HD double fast_record_small_e_tmax(Load &&c) {
    const double bc4 = fabs(c(KSL_F_BC4)), bc5 = 2.0 * fabs(c(KSL_F_BC5)), om = fabs(c(KSL_C_OMGCOF));
    const double ae = 1.0 + fabs(c(KSL_C_ETA));
    const double room_up = 0.0095 - 1.5e-3 - bc5 - c(KSL_C_ECCO);                 // bound < 0.0095
    const double room_dn = c(KSL_C_ECCO) - bc5 - 2.0e-6;                           // above the em >= 1e-6 clamp
    const double room_an = 0.06 - fabs(c(KSL_C_XMCOF)) * (ae * ae * ae + fabs(c(KSL_C_DELMO)));   // drag angle < 0.06
    if (!(room_up > 0.0 && room_dn > 0.0 && room_an > 0.0)) return -1.0;
    double t = INFINITY;
    if (bc4 > 0.0) t = fmin(t, fmin(room_up, room_dn) / bc4);
    if (om > 0.0) t = fmin(t, room_an / om);
    return t;
}

// In-place preparation of a private (shared-memory / local) copy of a record, KSL_NFAST entries of `stride`
// doubles, for sgp4_prop_fast.  The exact path ignores the three masked constants when isimp is set and
// never reads beyond KSL_NCONST, so the prepared copy serves both paths.
KSL_HD void prepare_fast_record(double *c, int stride) {
    auto at = [&](int k) -> double & { return c[(long long)k * stride]; };
    if (at(KSL_C_ISIMP) != 0.0) {
        at(KSL_C_OMGCOF) = 0.0;
        at(KSL_C_XMCOF) = 0.0;
        at(KSL_C_CC5) = 0.0;
    }
    at(KSL_F_BC4) = at(KSL_C_BSTAR) * at(KSL_C_CC4);
    at(KSL_F_BC5) = at(KSL_C_BSTAR) * at(KSL_C_CC5);
    at(KSL_F_KMRT) = 0.75 * kJ2 * at(KSL_C_CON41);
    at(KSL_F_KX1) = 0.25 * kJ2 * at(KSL_C_X1MTH2);
    at(KSL_F_KSU) = -0.125 * kJ2 * at(KSL_C_X7THM1);
    at(KSL_F_KNODE) = 0.75 * kJ2 * at(KSL_C_COSIO);
    at(KSL_F_KINC) = 0.75 * kJ2 * at(KSL_C_COSIO) * at(KSL_C_SINIO);
    // KSL_F_PAD is the caller's (the warp kernel keeps the record's epoch there)
}

// A record that lives only as its 36 public constants (global memory, registers) seen as a fast record: the derived
// entries are formed on the fly (one multiplication each, what sgp4_prop does anyway) and the three constants the
// exact path ignores under isimp read as zero.  k is a literal at every call site, so the switch folds away.
template <typename Base>
struct FastView {
    Base base;
    bool isimp;
    KSL_HD double operator()(int k) const {
        switch (k) {
            case KSL_C_OMGCOF: case KSL_C_XMCOF: case KSL_C_CC5: return isimp ? 0.0 : base(k);
            case KSL_F_BC4: return base(KSL_C_BSTAR) * base(KSL_C_CC4);
            case KSL_F_BC5: return isimp ? 0.0 : base(KSL_C_BSTAR) * base(KSL_C_CC5);
            case KSL_F_KMRT: return (0.75 * kJ2) * base(KSL_C_CON41);
            case KSL_F_KX1: return (0.25 * kJ2) * base(KSL_C_X1MTH2);
            case KSL_F_KSU: return (-0.125 * kJ2) * base(KSL_C_X7THM1);
            case KSL_F_KNODE: return (0.75 * kJ2) * base(KSL_C_COSIO);
            case KSL_F_KINC: return (0.75 * kJ2) * base(KSL_C_COSIO) * base(KSL_C_SINIO);
            default: return base(k);
        }
    }
};
// full state at t from the 36 constants behind `base`: fast path, exact path when it declines.  Returns the error bits.
// (the exact path is kept out of line on the device so that its registers do not count against the callers' hot path)
template <typename Base>
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
int sgp4_state_exact(const Base &base, double t, double *r, double *v) {
    return sgp4_prop<true>(base, t, r, v);
}
template <typename Base>
KSL_HD int sgp4_state(Base base, double t, double r[3], double v[3], const double *grid) {
    const FastView<Base> view{base, base(KSL_C_ISIMP) != 0.0};
    if (sgp4_prop_fast_rv<true>(view, t, r, v, grid)) return 0;
    double rx[3], vx[3];   // scratch: the out-of-line call takes addresses, r / v should stay in registers
    const int e = sgp4_state_exact(base, t, rx, vx);
    r[0] = rx[0]; r[1] = rx[1]; r[2] = rx[2];
    v[0] = vx[0]; v[1] = vx[1]; v[2] = vx[2];
    return e;
}

// sgp4init + full state for one element set (the Monte Carlo kernels: one record per sample, used once).  The record
// stays in registers: the fast path reads it through a FastView, and when it declines, an out-of-line routine redoes
// init + the exact path from the 7 elements, so nothing ever takes the record's address.
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
int sgp4_init_state_exact(const double *el, double t, double *r, double *v) {
    double c[KSL_NCONST];
    int e = sgp4_init(el, [&](int k, double val) { c[k] = val; });
    e |= sgp4_prop<true>([&](int k) { return c[k]; }, t, r, v);
    return e;
}
KSL_HD int sgp4_init_state(const double el[7], double t, double r[3], double v[3], const double *grid) {
    double c[KSL_NCONST];
    const int e = sgp4_init(el, [&](int k, double val) { c[k] = val; });
    auto ld = [&](int k) { return c[k]; };
    const FastView<decltype(ld)> view{ld, c[KSL_C_ISIMP] != 0.0};
    if (sgp4_prop_fast_rv<true>(view, t, r, v, grid)) return e;
    double elx[7], rx[3], vx[3];
    for (int k = 0; k < 7; ++k) elx[k] = el[k];
    const int e2 = sgp4_init_state_exact(elx, t, rx, vx);
    r[0] = rx[0]; r[1] = rx[1]; r[2] = rx[2];
    v[0] = vx[0]; v[1] = vx[1]; v[2] = vx[2];
    return e2;
}

// sgp4init (fast form) + full state, everything in registers and nothing out of line: `ok` false = the sample needs the
// exact route (sgp4_init + sgp4_prop), nothing usable was written.  *err: sgp4init's flag (deep space).
// small_e: evaluate through the near-circular form (the caller's guess for the whole launch; a sample outside it is declined)
template <bool small_e = false>
KSL_HD bool sgp4_init_state_fast(const double el[7], double t, double r[3], double v[3], const double *grid, int *err) {
    double c[KSL_NCONST];
    *err = sgp4_init_fast(el, [&](int k, double val) { c[k] = val; }, grid);
    auto ld = [&](int k) { return c[k]; };
    const FastView<decltype(ld)> view{ld, c[KSL_C_ISIMP] != 0.0};
    return sgp4_prop_fast_rv<true, small_e>(view, t, r, v, grid);
}

// position at t through the fast path, falling back to the exact path when needed (c: masked record).  The exact path
// is kept out of line on the device so that its registers do not count against the hot path.
#if defined(__CUDACC__)
__device__ __noinline__ int sgp4_position_exact(const double *c, int stride, double t, double *r) {
    double v[3];
    return sgp4_prop<false>([&](int k) { return c[(long long)k * stride]; }, t, r, v);
}
#endif
// c: pointer to constant 0 of the record, `stride` doubles between consecutive constants
KSL_HD int sgp4_position(const double *c, int stride, double t, double r[3], const double *grid, bool small_e = false) {
    auto ld = [&](int k) { return c[(long long)k * stride]; };
    if (small_e ? sgp4_prop_fast<true>(ld, t, r, grid) : sgp4_prop_fast<false>(ld, t, r, grid)) return 0;
#if defined(__CUDA_ARCH__)
    return sgp4_position_exact(c, stride, t, r);
#else
    double v[3];
    return sgp4_prop<false>([&](int k) { return c[(long long)k * stride]; }, t, r, v);
#endif
}

// ---- util.upsample index arithmetic (torch upsample_linear1d, align_corners) ----
// scale = (n_in-1)/(n_out-1) in FP64; the integer part goes through FLOAT32.
KSL_HD void upsample_index(int64_t j, double scale, int64_t n_in, int64_t *i0, int64_t *i1, double *w0, double *w1) {
    const double src = mul_rn((double)j, scale);   // rounded product, as torch forms it (no contraction with the subtraction below)
    int64_t a = (int64_t)floorf((float)src);
    if (a > n_in - 1) a = n_in - 1;
    const int64_t b = (a + 1 < n_in) ? a + 1 : n_in - 1;
    double l = add_rn(src, -(double)a);
    l = l < 0.0 ? 0.0 : l;
    l = l > 1.0 ? 1.0 : l;
    *i0 = a;
    *i1 = b;
    *w1 = l;
    *w0 = 1.0 - l;
}

// one interpolated coordinate, rounding exactly as torch's CPU upsample_linear1d kernel:
// fma(w0, x0, w1*x1) (established against torch 2.11, tests/test_oracle_golden.py)
KSL_HD double lerp_torch(double w0, double x0, double w1, double x1) { return fma_rn(w0, x0, mul_rn(w1, x1)); }
// ... and in metres (util.py:579 `*1e3`)
KSL_HD double lerp_m(double w0, double x0, double w1, double x1) { return mul_rn(lerp_torch(w0, x0, w1, x1), 1.0e3); }

// squared distance, direct-difference form, no contraction
KSL_HD double sqdist3(double dx, double dy, double dz) {
    return add_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)), mul_rn(dz, dz));
}

}  // namespace ksl
