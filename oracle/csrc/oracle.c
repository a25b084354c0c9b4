/*
 * oracle.c -- scalar C restatement of the Kessler Monte Carlo hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): linked by tests/ and by
 * the cpu_baseline / reference legs of bench.py, never by kessler_b200.
 *
 * Follows, function by function:
 *   orc_sgp4_init / orc_sgp4_prop  dsgp4.initialize_tle / dsgp4.propagate[_batch]
 *       (called at /root/reference/kessler/model.py:581,592,536,541 and
 *       util.py:570-576).  dsgp4 (esa/dSGP4) is an un-vendored, un-pinned
 *       dependency (setup.py:35); restated from Vallado et al., AIAA 2006-6753,
 *       near-earth branch, WGS-84, floor-mod wrapping.
 *   orc_upsample                   kessler/util.py:541-555 (F.interpolate linear,
 *       align_corners=True, incl. torch's float32 floor of the source index)
 *   orc_find_conjunction           kessler/model.py:23-52
 *   orc_screen                     kessler/model.py:574-613 (space segment of forward())
 *   orc_tca_states / orc_mc_cov    kessler/model.py:536-550
 *   orc_pc_count                   kessler/model.py:432-437 (direct-difference form)
 *
 * Build: oracle/Makefile  (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_NCONST 36
enum {
    C_NO = 0, C_ECCO, C_INCLO, C_NODEO, C_ARGPO, C_MO, C_BSTAR,
    C_MDOT, C_ARGPDOT, C_NODEDOT, C_NODECF, C_CC1, C_CC4, C_CC5, C_T2COF,
    C_OMGCOF, C_XMCOF, C_ETA, C_DELMO, C_SINMAO, C_D2, C_D3, C_D4,
    C_T3COF, C_T4COF, C_T5COF, C_AYCOF, C_XLCOF, C_CON41, C_X1MTH2,
    C_X7THM1, C_ISIMP, C_COSIO, C_SINIO, C_AOBASE, C_PAD
};

#define ERR_ECC 1
#define ERR_MEANMOTION 2
#define ERR_SEMILATUS 4
#define ERR_DECAYED 8
#define ERR_DEEPSPACE 16

static const double MU = 398600.5, RE = 6378.137;
static const double J2 = 0.00108262998905, J3 = -0.00000253215306, J4 = -0.00000161098761;
#define TWOPI 6.283185307179586

static double xke_(void) { return 60.0 / sqrt(RE * RE * RE / MU); }

static double floormod(double x, double y)
{ /* Python/torch `%`: result has the sign of y */
    double r = fmod(x, y);
    if (r != 0.0 && ((r < 0.0) != (y < 0.0))) r += y;
    return r;
}

/* ---- sgp4init -------------------------------------------------------------- */
static int init_one(const double el[7], double c[ORC_NCONST])
{
    const double xke = xke_(), j3oj2 = J3 / J2, x2o3 = 2.0 / 3.0;
    double no_kozai = el[0], ecco = el[1], inclo = el[2], nodeo = el[3], argpo = el[4], mo = el[5], bstar = el[6];
    double eccsq = ecco * ecco, omeosq = 1.0 - eccsq, rteosq = sqrt(omeosq);
    double cosio = cos(inclo), cosio2 = cosio * cosio;
    double ak = pow(xke / no_kozai, x2o3);
    double d1 = 0.75 * J2 * (3.0 * cosio2 - 1.0) / (rteosq * omeosq);
    double del = d1 / (ak * ak);
    double adel = ak * (1.0 - del * del - del * (1.0 / 3.0 + 134.0 * del * del / 81.0));
    del = d1 / (adel * adel);
    double no = no_kozai / (1.0 + del);
    double ao = pow(xke / no, x2o3);
    double sinio = sin(inclo);
    double po = ao * omeosq;
    double con42 = 1.0 - 5.0 * cosio2;
    double con41 = -con42 - cosio2 - cosio2;
    double posq = po * po;
    double rp = ao * (1.0 - ecco);
    int isimp = rp < (220.0 / RE + 1.0);
    double sfour = 78.0 / RE + 1.0;
    double qz = (120.0 - 78.0) / RE;
    double qzms24 = qz * qz * qz * qz;
    double perige = (rp - 1.0) * RE;
    if (perige < 156.0) {
        sfour = perige - 78.0;
        if (perige < 98.0) sfour = 20.0;
        double q = (120.0 - sfour) / RE;
        qzms24 = q * q * q * q;
        sfour = sfour / RE + 1.0;
    }
    double pinvsq = 1.0 / posq;
    double tsi = 1.0 / (ao - sfour);
    double eta = ao * ecco * tsi, etasq = eta * eta, eeta = ecco * eta;
    double psisq = fabs(1.0 - etasq);
    double coef = qzms24 * pow(tsi, 4.0);
    double coef1 = coef / pow(psisq, 3.5);
    double cc2 = coef1 * no * (ao * (1.0 + 1.5 * etasq + eeta * (4.0 + etasq))
                 + 0.375 * J2 * tsi / psisq * con41 * (8.0 + 3.0 * etasq * (8.0 + etasq)));
    double cc1 = bstar * cc2;
    double cc3 = 0.0;
    if (ecco > 1.0e-4) cc3 = -2.0 * coef * tsi * j3oj2 * no * sinio / ecco;
    double x1mth2 = 1.0 - cosio2;
    double cc4 = 2.0 * no * coef1 * ao * omeosq * (eta * (2.0 + 0.5 * etasq) + ecco * (0.5 + 2.0 * etasq)
                 - J2 * tsi / (ao * psisq) * (-3.0 * con41 * (1.0 - 2.0 * eeta + etasq * (1.5 - 0.5 * eeta))
                 + 0.75 * x1mth2 * (2.0 * etasq - eeta * (1.0 + etasq)) * cos(2.0 * argpo)));
    double cc5 = 2.0 * coef1 * ao * omeosq * (1.0 + 2.75 * (etasq + eeta) + eeta * etasq);
    double cosio4 = cosio2 * cosio2;
    double temp1 = 1.5 * J2 * pinvsq * no;
    double temp2 = 0.5 * temp1 * J2 * pinvsq;
    double temp3 = -0.46875 * J4 * pinvsq * pinvsq * no;
    double mdot = no + 0.5 * temp1 * rteosq * con41 + 0.0625 * temp2 * rteosq * (13.0 - 78.0 * cosio2 + 137.0 * cosio4);
    double argpdot = -0.5 * temp1 * con42 + 0.0625 * temp2 * (7.0 - 114.0 * cosio2 + 395.0 * cosio4)
                     + temp3 * (3.0 - 36.0 * cosio2 + 49.0 * cosio4);
    double xhdot1 = -temp1 * cosio;
    double nodedot = xhdot1 + (0.5 * temp2 * (4.0 - 19.0 * cosio2) + 2.0 * temp3 * (3.0 - 7.0 * cosio2)) * cosio;
    double omgcof = bstar * cc3 * cos(argpo);
    double xmcof = 0.0;
    if (ecco > 1.0e-4) xmcof = -x2o3 * coef * bstar / eeta;
    double nodecf = 3.5 * omeosq * xhdot1 * cc1;
    double t2cof = 1.5 * cc1;
    double xlcof;
    if (fabs(cosio + 1.0) > 1.5e-12) xlcof = -0.25 * j3oj2 * sinio * (3.0 + 5.0 * cosio) / (1.0 + cosio);
    else xlcof = -0.25 * j3oj2 * sinio * (3.0 + 5.0 * cosio) / 1.5e-12;
    double aycof = -0.5 * j3oj2 * sinio;
    double delmotemp = 1.0 + eta * cos(mo);
    double delmo = delmotemp * delmotemp * delmotemp;
    double sinmao = sin(mo);
    double x7thm1 = 7.0 * cosio2 - 1.0;
    double d2 = 0, d3 = 0, d4 = 0, t3cof = 0, t4cof = 0, t5cof = 0;
    if (!isimp) {
        double cc1sq = cc1 * cc1;
        d2 = 4.0 * ao * tsi * cc1sq;
        double temp = d2 * tsi * cc1 / 3.0;
        d3 = (17.0 * ao + sfour) * temp;
        d4 = 0.5 * temp * ao * tsi * (221.0 * ao + 31.0 * sfour) * cc1;
        t3cof = d2 + 2.0 * cc1sq;
        t4cof = 0.25 * (3.0 * d3 + cc1 * (12.0 * d2 + 10.0 * cc1sq));
        t5cof = 0.2 * (3.0 * d4 + 12.0 * cc1 * d3 + 6.0 * d2 * d2 + 15.0 * cc1sq * (2.0 * d2 + cc1sq));
    }
    c[C_NO] = no; c[C_ECCO] = ecco; c[C_INCLO] = inclo; c[C_NODEO] = nodeo; c[C_ARGPO] = argpo; c[C_MO] = mo;
    c[C_BSTAR] = bstar; c[C_MDOT] = mdot; c[C_ARGPDOT] = argpdot; c[C_NODEDOT] = nodedot; c[C_NODECF] = nodecf;
    c[C_CC1] = cc1; c[C_CC4] = cc4; c[C_CC5] = cc5; c[C_T2COF] = t2cof; c[C_OMGCOF] = omgcof; c[C_XMCOF] = xmcof;
    c[C_ETA] = eta; c[C_DELMO] = delmo; c[C_SINMAO] = sinmao; c[C_D2] = d2; c[C_D3] = d3; c[C_D4] = d4;
    c[C_T3COF] = t3cof; c[C_T4COF] = t4cof; c[C_T5COF] = t5cof; c[C_AYCOF] = aycof; c[C_XLCOF] = xlcof;
    c[C_CON41] = con41; c[C_X1MTH2] = x1mth2; c[C_X7THM1] = x7thm1; c[C_ISIMP] = isimp ? 1.0 : 0.0;
    c[C_COSIO] = cosio; c[C_SINIO] = sinio; c[C_AOBASE] = ao; c[C_PAD] = 0.0;
    if (!(isfinite(no) && no > 0.0)) return ERR_DEEPSPACE;
    return (TWOPI / no >= 225.0) ? ERR_DEEPSPACE : 0;
}

/* ---- sgp4 ------------------------------------------------------------------ */
static int prop_one(const double c[ORC_NCONST], double t, double rv[6])
{
    const double xke = xke_(), vkmpersec = RE * xke / 60.0;
    int err = 0;
    double xmdf = c[C_MO] + c[C_MDOT] * t;
    double argpdf = c[C_ARGPO] + c[C_ARGPDOT] * t;
    double nodedf = c[C_NODEO] + c[C_NODEDOT] * t;
    double argpm = argpdf, mm = xmdf;
    double t2 = t * t;
    double nodem = nodedf + c[C_NODECF] * t2;
    double tempa = 1.0 - c[C_CC1] * t;
    double tempe = c[C_BSTAR] * c[C_CC4] * t;
    double templ = c[C_T2COF] * t2;
    if (c[C_ISIMP] == 0.0) {
        double delomg = c[C_OMGCOF] * t;
        double delmtemp = 1.0 + c[C_ETA] * cos(xmdf);
        double delm = c[C_XMCOF] * (delmtemp * delmtemp * delmtemp - c[C_DELMO]);
        double temp = delomg + delm;
        mm = xmdf + temp;
        argpm = argpdf - temp;
        double t3 = t2 * t, t4 = t3 * t;
        tempa = tempa - c[C_D2] * t2 - c[C_D3] * t3 - c[C_D4] * t4;
        tempe = tempe + c[C_BSTAR] * c[C_CC5] * (sin(mm) - c[C_SINMAO]);
        templ = templ + c[C_T3COF] * t3 + t4 * (c[C_T4COF] + t * c[C_T5COF]);
    }
    double no = c[C_NO];
    double am = c[C_AOBASE] * tempa * tempa;   /* (xke/no)^(2/3), same value sgp4init computed */
    double nm = xke / pow(am, 1.5);
    double em = c[C_ECCO] - tempe;
    if (em >= 1.0 || em < -0.001) err |= ERR_ECC;
    if (nm <= 0.0) err |= ERR_MEANMOTION;
    if (em < 1.0e-6) em = 1.0e-6;
    mm = mm + no * templ;
    double xlm = mm + argpm + nodem;
    nodem = floormod(nodem, TWOPI);
    argpm = floormod(argpm, TWOPI);
    xlm = floormod(xlm, TWOPI);
    mm = floormod(xlm - argpm - nodem, TWOPI);

    double sinim = c[C_SINIO], cosim = c[C_COSIO];
    double axnl = em * cos(argpm);
    double temp = 1.0 / (am * (1.0 - em * em));
    double aynl = em * sin(argpm) + temp * c[C_AYCOF];
    double xl = mm + argpm + nodem + temp * c[C_XLCOF] * axnl;
    double u = floormod(xl - nodem, TWOPI);
    double eo1 = u, tem5 = 9999.9, sineo1 = 0.0, coseo1 = 1.0;
    int ktr = 1;
    while (fabs(tem5) >= 1.0e-12 && ktr <= 10) {
        sineo1 = sin(eo1);
        coseo1 = cos(eo1);
        tem5 = 1.0 - coseo1 * axnl - sineo1 * aynl;
        tem5 = (u - aynl * coseo1 + axnl * sineo1 - eo1) / tem5;
        if (tem5 >= 0.95) tem5 = 0.95;
        if (tem5 <= -0.95) tem5 = -0.95;
        eo1 = eo1 + tem5;
        ktr++;
    }
    /* Vallado leaves sineo1/coseo1 one update behind eo1.  A vectorised
     * dsgp4.propagate call keeps iterating EVERY lane until the slowest of its
     * (6000) lanes has converged, so a converged lane has seen at least one
     * more pass and its sin/cos match eo1 to rounding; model that by a refresh.
     * A lane that ran out of iterations keeps the stale pair, as in dsgp4. */
    if (fabs(tem5) < 1.0e-12) {
        sineo1 = sin(eo1);
        coseo1 = cos(eo1);
    }
    double ecose = axnl * coseo1 + aynl * sineo1;
    double esine = axnl * sineo1 - aynl * coseo1;
    double el2 = axnl * axnl + aynl * aynl;
    double pl = am * (1.0 - el2);
    if (pl < 0.0) err |= ERR_SEMILATUS;
    double rl = am * (1.0 - ecose);
    double rdotl = sqrt(am) * esine / rl;
    double rvdotl = sqrt(pl) / rl;
    double betal = sqrt(1.0 - el2);
    temp = esine / (1.0 + betal);
    double sinu = am / rl * (sineo1 - aynl - axnl * temp);
    double cosu = am / rl * (coseo1 - axnl + aynl * temp);
    double su = atan2(sinu, cosu);
    double sin2u = (cosu + cosu) * sinu;
    double cos2u = 1.0 - 2.0 * sinu * sinu;
    temp = 1.0 / pl;
    double temp1 = 0.5 * J2 * temp;
    double temp2 = temp1 * temp;
    double mrt = rl * (1.0 - 1.5 * temp2 * betal * c[C_CON41]) + 0.5 * temp1 * c[C_X1MTH2] * cos2u;
    su = su - 0.25 * temp2 * c[C_X7THM1] * sin2u;
    double xnode = nodem + 1.5 * temp2 * cosim * sin2u;
    double xinc = c[C_INCLO] + 1.5 * temp2 * cosim * sinim * cos2u;
    double mvt = rdotl - nm * temp1 * c[C_X1MTH2] * sin2u / xke;
    double rvdot = rvdotl + nm * temp1 * (c[C_X1MTH2] * cos2u + 1.5 * c[C_CON41]) / xke;
    double sinsu = sin(su), cossu = cos(su), snod = sin(xnode), cnod = cos(xnode), sini = sin(xinc), cosi = cos(xinc);
    double xmx = -snod * cosi, xmy = cnod * cosi;
    double ux = xmx * sinsu + cnod * cossu, uy = xmy * sinsu + snod * cossu, uz = sini * sinsu;
    double vx = xmx * cossu - cnod * sinsu, vy = xmy * cossu - snod * sinsu, vz = sini * cossu;
    double mr = mrt * RE;
    rv[0] = mr * ux; rv[1] = mr * uy; rv[2] = mr * uz;
    rv[3] = (mvt * ux + rvdot * vx) * vkmpersec;
    rv[4] = (mvt * uy + rvdot * vy) * vkmpersec;
    rv[5] = (mvt * uz + rvdot * vz) * vkmpersec;
    if (mrt < 1.0) err |= ERR_DECAYED;
    return err;
}

/* elems: SoA [7][n]; consts: SoA [ORC_NCONST][n] */
int orc_sgp4_init(const double *elems, int64_t n, double *consts, int32_t *err)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        double el[7], c[ORC_NCONST];
        for (int k = 0; k < 7; k++) el[k] = elems[k * n + i];
        int e = init_one(el, c);
        for (int k = 0; k < ORC_NCONST; k++) consts[k * n + i] = c[k];
        if (err) err[i] = e;
    }
    return 0;
}

/* rv: [n][m][6] km, km/s; tsince [m] shared or [n][m] (per_sample_t); err [n] OR over epochs */
int orc_sgp4_prop(const double *consts, int64_t n, const double *tsince, int64_t m, int per_sample_t,
                  double *rv, int32_t *err)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        double c[ORC_NCONST];
        for (int k = 0; k < ORC_NCONST; k++) c[k] = consts[k * n + i];
        int e = 0;
        for (int64_t j = 0; j < m; j++) {
            double t = per_sample_t ? tsince[i * m + j] : tsince[j];
            e |= prop_one(c, t, rv + (i * m + j) * 6);
        }
        if (err) err[i] |= e;
    }
    return 0;
}

/* ---- util.upsample: index/weight of torch upsample_linear1d, align_corners ---- */
static inline void up_idx(int64_t j, int64_t n_in, int64_t n_out, int64_t *i0, int64_t *i1, double *w0, double *w1)
{
    double scale = (n_out > 1) ? (double)(n_in - 1) / (double)(n_out - 1) : 0.0;
    double src = (double)j * scale;
    int64_t a = (int64_t)floorf((float)src);   /* float32 floor: torch quirk */
    if (a > n_in - 1) a = n_in - 1;
    int64_t b = a + 1 < n_in ? a + 1 : n_in - 1;
    double l = src - (double)a;
    if (l < 0.0) l = 0.0;
    if (l > 1.0) l = 1.0;
    *i0 = a; *i1 = b; *w1 = l; *w0 = 1.0 - l;
}

/* torch's CPU upsample_linear1d kernel evaluates w0*x0 + w1*x1 as fma(w0, x0, w1*x1)
 * (established empirically against torch 2.11: tests/test_oracle_golden.py checks
 * bit-equality on the production 6000 -> 600000 shape) */
static inline double lerp_torch(double w0, double x0, double w1, double x1) { return fma(w0, x0, w1 * x1); }

int orc_upsample(const double *x, int64_t n_in, int64_t ch, int64_t n_out, double *y)
{
    for (int64_t j = 0; j < n_out; j++) {
        int64_t i0, i1; double w0, w1;
        up_idx(j, n_in, n_out, &i0, &i1, &w0, &w1);
        for (int64_t c = 0; c < ch; c++) y[j * ch + c] = lerp_torch(w0, x[i0 * ch + c], w1, x[i1 * ch + c]);
    }
    return 0;
}

/* ---- find_conjunction (model.py:23-52); NaN is the minimum, first index wins ---- */
int orc_find_conjunction(const double *tr0, const double *tr1, int64_t n, double thr,
                         int64_t *i_min, double *d_min, int64_t *i_conj, double *d_conj)
{
    double best = INFINITY, thr2 = thr * thr, sqc = NAN;
    int64_t bi = 0, ci = -1;
    int have_nan = 0;
    for (int64_t j = 0; j < n; j++) {
        double dx = tr0[3 * j] - tr1[3 * j], dy = tr0[3 * j + 1] - tr1[3 * j + 1], dz = tr0[3 * j + 2] - tr1[3 * j + 2];
        double sq = (dx * dx + dy * dy) + dz * dz;
        if (!have_nan) {
            if (isnan(sq)) { have_nan = 1; bi = j; best = sq; }
            else if (j == 0 || sq < best) { best = sq; bi = j; }
        }
        if (ci < 0 && sq < thr2) { ci = j; sqc = sq; }
    }
    *i_min = bi; *d_min = sqrt(best); *i_conj = ci; *d_conj = (ci >= 0) ? sqrt(sqc) : NAN;
    return 0;
}

/* ---- one full trace, brute force (model.py:574-613) --------------------------
 * elems_t SoA [7][n_t] (n_t = 1: one target shared by all samples, else n_t = n);
 * epoch_* = mjd of the TLE epoch (util.py:570,575); times_coarse [n_coarse] =
 * times[::upsample_factor] (util.py:574); n_fine = len(times).
 * out: i_min,i_conj int64 [n]; d_min,d_conj double [n] (metres); err int32 [n]
 * (bits 0-4 target, bits 8-12 chaser). Deep-space on either -> i_min=-1. */
int orc_screen(const double *elems_t, const double *epoch_t, int64_t n_t,
               const double *elems_c, const double *epoch_c, int64_t n,
               const double *times_coarse, int64_t n_coarse, int64_t n_fine, double thr,
               int64_t *i_min, double *d_min, int64_t *i_conj, double *d_conj, int32_t *err)
{
    int rc = 0;
#pragma omp parallel
    {
        double *st = (double *)malloc(sizeof(double) * 6 * n_coarse);
        double *sc = (double *)malloc(sizeof(double) * 6 * n_coarse);
        if (!st || !sc) rc = -1;
#pragma omp for schedule(dynamic, 1)
        for (int64_t s = 0; s < n; s++) {
            if (rc) continue;
            int64_t it = (n_t == 1) ? 0 : s;
            double el[7], ct[ORC_NCONST], cc[ORC_NCONST];
            for (int k = 0; k < 7; k++) el[k] = elems_t[k * n_t + it];
            int e0 = init_one(el, ct);
            for (int k = 0; k < 7; k++) el[k] = elems_c[k * n + s];
            int e1 = init_one(el, cc);
            if ((e0 | e1) & ERR_DEEPSPACE) {
                i_min[s] = -1; d_min[s] = NAN; i_conj[s] = -1; d_conj[s] = NAN; err[s] = e0 | (e1 << 8);
                continue;
            }
            for (int64_t k = 0; k < n_coarse; k++) {
                e0 |= prop_one(ct, (times_coarse[k] - epoch_t[it]) * 1440.0, st + 6 * k);
                e1 |= prop_one(cc, (times_coarse[k] - epoch_c[s]) * 1440.0, sc + 6 * k);
            }
            double best = INFINITY, thr2 = thr * thr, sqc = NAN;
            int64_t bi = 0, ci = -1;
            int have_nan = 0;
            for (int64_t j = 0; j < n_fine; j++) {
                int64_t a, b; double w0, w1;
                up_idx(j, n_coarse, n_fine, &a, &b, &w0, &w1);
                double d[3];
                for (int x = 0; x < 3; x++) {
                    double pt = lerp_torch(w0, st[6 * a + x], w1, st[6 * b + x]) * 1e3;   /* util.py:579: *1e3 -> m */
                    double pc = lerp_torch(w0, sc[6 * a + x], w1, sc[6 * b + x]) * 1e3;
                    d[x] = pt - pc;
                }
                double sq = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2];
                if (!have_nan) {
                    if (isnan(sq)) { have_nan = 1; bi = j; best = sq; }
                    else if (j == 0 || sq < best) { best = sq; bi = j; }
                }
                if (ci < 0 && sq < thr2) { ci = j; sqc = sq; }
            }
            i_min[s] = bi; d_min[s] = sqrt(best); i_conj[s] = ci; d_conj[s] = (ci >= 0) ? sqrt(sqc) : NAN;
            err[s] = e0 | (e1 << 8);
        }
        free(st); free(sc);
    }
    return rc;
}

/* ---- model.py:536-541: batch of element sets at one tsince -> [n][6] metres ---- */
int orc_tca_states(const double *elems /*SoA [7][n]*/, int64_t n, double tsince, double *states_m, int32_t *err)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        double el[7], c[ORC_NCONST], rv[6];
        for (int k = 0; k < 7; k++) el[k] = elems[k * n + i];
        int e = init_one(el, c);
        e |= prop_one(c, tsince, rv);
        for (int k = 0; k < 6; k++) states_m[6 * i + k] = rv[k] * 1e3;
        if (err) err[i] = e;
    }
    return 0;
}

static void rot_from_state(const double s[6], double R[9])
{ /* util.py:304-320 */
    double rn = sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
    double u[3] = { s[0] / rn, s[1] / rn, s[2] / rn };
    double w[3] = { s[1] * s[5] - s[2] * s[4], s[2] * s[3] - s[0] * s[5], s[0] * s[4] - s[1] * s[3] };
    double wn = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    w[0] /= wn; w[1] /= wn; w[2] /= wn;
    double t[3] = { w[1] * u[2] - w[2] * u[1], w[2] * u[0] - w[0] * u[2], w[0] * u[1] - w[1] * u[0] };
    for (int k = 0; k < 3; k++) { R[k] = u[k]; R[3 + k] = t[k]; R[6 + k] = w[k]; }
}

/* ---- model.py:544-550: mean [6], RTN covariance [6][6] (ddof = 1), two-pass ---- */
int orc_mc_cov(const double *states_m /*[n][6]*/, int64_t n, double *mean, double *cov)
{
    double m[6] = { 0 };
    for (int64_t i = 0; i < n; i++) for (int k = 0; k < 6; k++) m[k] += states_m[6 * i + k];
    for (int k = 0; k < 6; k++) { m[k] /= (double)n; mean[k] = m[k]; }
    double R[9];
    rot_from_state(m, R);
    /* np.cov subtracts the sample mean of the rotated rows */
    double mrow[6] = { 0 };
    double *rot = (double *)malloc(sizeof(double) * 6 * n);
    if (!rot) return -1;
    for (int64_t i = 0; i < n; i++) for (int h = 0; h < 2; h++) for (int a = 0; a < 3; a++) {
        const double *s = states_m + 6 * i + 3 * h;
        double v = R[3 * a] * s[0] + R[3 * a + 1] * s[1] + R[3 * a + 2] * s[2];
        rot[6 * i + 3 * h + a] = v; mrow[3 * h + a] += v;
    }
    for (int k = 0; k < 6; k++) mrow[k] /= (double)n;
    for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) {
        double acc = 0;
        for (int64_t i = 0; i < n; i++) acc += (rot[6 * i + a] - mrow[a]) * (rot[6 * i + b] - mrow[b]);
        cov[6 * a + b] = acc / (double)(n - 1);
    }
    free(rot);
    return 0;
}

/* ---- model.py:432-437: collision count, direct-difference form, no FMA ------- */
int orc_pc_count(const double *t_xyz /*SoA [3][nt]*/, int64_t nt, const double *c_xyz /*SoA [3][nc]*/, int64_t nc,
                 double thr, int pairing /*0 all pairs, 1 elementwise*/, uint64_t *count)
{
    double thr2 = thr * thr;
    uint64_t total = 0;
    if (pairing == 1) {
        int64_t n = nt < nc ? nt : nc;
#pragma omp parallel for reduction(+ : total) schedule(static)
        for (int64_t i = 0; i < n; i++) {
            double dx = t_xyz[i] - c_xyz[i], dy = t_xyz[nt + i] - c_xyz[nc + i], dz = t_xyz[2 * nt + i] - c_xyz[2 * nc + i];
            total += ((dx * dx + dy * dy) + dz * dz) < thr2;
        }
    } else {
#pragma omp parallel for reduction(+ : total) schedule(static)
        for (int64_t i = 0; i < nt; i++) {
            double tx = t_xyz[i], ty = t_xyz[nt + i], tz = t_xyz[2 * nt + i];
            uint64_t c = 0;
            for (int64_t j = 0; j < nc; j++) {
                double dx = tx - c_xyz[j], dy = ty - c_xyz[nc + j], dz = tz - c_xyz[2 * nc + j];
                c += ((dx * dx + dy * dy) + dz * dz) < thr2;
            }
            total += c;
        }
    }
    *count = total;
    return 0;
}

int orc_nconst(void) { return ORC_NCONST; }
