/*
 * kessler_b200.h -- C ABI of libkessler_b200.so: the B200 (sm_100a) implementation
 * of Kessler's Monte Carlo conjunction hot path.
 *
 * The reference (kesslerlib/kessler v1.0.1) is pure Python and has no FFI; the
 * path sits behind Python calls into the third-party package dsgp4.  Each entry
 * point below names the reference call it replaces (file:line under
 * /root/reference/kessler/).  INTEGRATION.md shows the ctypes stubs a Kessler
 * maintainer would add to route those calls here.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - every function returns 0 on success, <0 on failure (KSL_E_*); it never
 *     throws.  ksl_last_error() gives the message of the calling thread's last
 *     failure.  Per-sample NUMERICAL conditions never fail a call: they are
 *     reported in the int32 `err` arrays as a bitmask (KSL_ERR_*).
 *   - pointers named d_* / without suffix are DEVICE pointers owned by the
 *     caller (e.g. torch tensors' data_ptr()); the *_host entry points take HOST
 *     pointers and do the transfers themselves (that is the end-to-end path).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     all device-pointer entry points are stream-ordered and asynchronous.
 *   - there is no CPU fallback: without a CUDA device every compute entry point
 *     returns KSL_E_CUDA.
 *
 * Units follow dsgp4: mean motion rad/min, angles rad, tsince minutes, SGP4
 * output km and km/s (TEME); the Kessler-level entry points return metres, as
 * util.propagate_upsample does (util.py:579).
 */
#ifndef KESSLER_B200_H
#define KESSLER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KSL_VERSION 200

/* return codes */
#define KSL_OK 0
#define KSL_E_ARG (-1)   /* bad argument (NULL pointer, negative size, unsupported combination) */
#define KSL_E_CUDA (-2)  /* CUDA runtime error, or no CUDA device */
#define KSL_E_NOMEM (-3) /* workspace too small / allocation failed */

/* per-sample error bits (dsgp4 keeps the last Vallado code in tle._error; only
 * the deep-space refusal is observable to Kessler, model.py:580-599) */
#define KSL_ERR_ECC 1        /* em >= 1 or em < -0.001 at some epoch (Vallado error 1) */
#define KSL_ERR_MEANMOTION 2 /* nm <= 0 (Vallado error 2) */
#define KSL_ERR_SEMILATUS 4  /* pl < 0 (Vallado error 4) */
#define KSL_ERR_DECAYED 8    /* mrt < 1: below the surface (Vallado error 6) */
#define KSL_ERR_DEEPSPACE 16 /* period >= 225 min: dsgp4.initialize_tle raises */
/* in ksl_screen the target's bits are in 0..7, the chaser's in 8..15 */
#define KSL_ERR_CHASER_SHIFT 8

/* layout of one initialised-TLE record, structure-of-arrays: consts[k*n + i] */
#define KSL_NCONST 36
enum ksl_const_index {
    KSL_C_NO = 0, KSL_C_ECCO, KSL_C_INCLO, KSL_C_NODEO, KSL_C_ARGPO, KSL_C_MO, KSL_C_BSTAR,
    KSL_C_MDOT, KSL_C_ARGPDOT, KSL_C_NODEDOT, KSL_C_NODECF, KSL_C_CC1, KSL_C_CC4, KSL_C_CC5, KSL_C_T2COF,
    KSL_C_OMGCOF, KSL_C_XMCOF, KSL_C_ETA, KSL_C_DELMO, KSL_C_SINMAO, KSL_C_D2, KSL_C_D3, KSL_C_D4,
    KSL_C_T3COF, KSL_C_T4COF, KSL_C_T5COF, KSL_C_AYCOF, KSL_C_XLCOF, KSL_C_CON41, KSL_C_X1MTH2,
    KSL_C_X7THM1, KSL_C_ISIMP, KSL_C_COSIO, KSL_C_SINIO, KSL_C_AOBASE, KSL_C_PAD
};
/* element order of the SoA input `elems[k*n + i]`, k = 0..6 */
enum ksl_elem_index { KSL_E_NO_KOZAI = 0, KSL_E_ECCO, KSL_E_INCLO, KSL_E_NODEO, KSL_E_ARGPO, KSL_E_MO, KSL_E_BSTAR };

/* ksl_screen `mode` */
#define KSL_SCREEN_ANALYTIC 0 /* per-coarse-segment closed form + exact re-evaluation of the candidates */
#define KSL_SCREEN_BRUTE 1    /* every fine-grid point, as the reference does */
/* OR-ed into `mode` (testing / tuning): pin the work decomposition instead of choosing by batch size */
#define KSL_SCREEN_FORCE_CTA 16  /* one CTA per trace  (small batches) */
#define KSL_SCREEN_FORCE_WARP 32 /* one warp per trace (large batches) */

/* ksl_pc_count `pairing` */
#define KSL_PAIR_ALL 0
#define KSL_PAIR_ELEMENTWISE 1

/* summary record written by ksl_screen_summary: uint64 counts + double moments */
#define KSL_SUM_NCOUNT 8
enum ksl_summary_count {
    KSL_SUM_N = 0,       /* traces seen */
    KSL_SUM_PROP_ERR,    /* deep-space on target or chaser -> `conj=False` (model.py:584,595) */
    KSL_SUM_CONJ_ANY,    /* i_conj >= 0 */
    KSL_SUM_CONJ_REF,    /* i_conj > 0: what `if i_conj:` accepts (model.py:617) */
    KSL_SUM_COLLISION,   /* d_min < collision threshold */
    KSL_SUM_NUM_ERR,     /* any numerical error bit other than deep-space */
    KSL_SUM_RESERVED0, KSL_SUM_RESERVED1
};
#define KSL_SUM_NMOMENT 4 /* n_valid, sum d_min, sum d_min^2 (about KSL_SUM_SHIFT), min d_min */

int ksl_version(void);
int ksl_nconst(void);
const char *ksl_last_error(void);
/* number of visible CUDA devices (0 when there is no driver/GPU); never fails */
int ksl_device_count(void);

/* ---- dsgp4.initialize_tle(tle | [tle...])  (model.py:581,592,536) ------------
 * elems SoA [7][n] -> consts SoA [KSL_NCONST][n]; err[n] = 0 or KSL_ERR_DEEPSPACE. */
int ksl_sgp4_init(const double *elems, int64_t n, double *consts, int32_t *err, void *stream);

/* ---- dsgp4.propagate(tle, tsinces) / propagate_batch(batch, tsinces) ----------
 * (util.py:571,576; model.py:541).  tsince is [m] shared by all n records
 * (per_sample_t = 0) or [n][m] (per_sample_t = 1).  rv_km [n][m][6] = x,y,z km,
 * vx,vy,vz km/s.  err[n] is OR-ed with the bits raised at any epoch (caller zeroes
 * or passes the array ksl_sgp4_init filled).
 * Numerical contract of every SGP4 evaluation behind this ABI: Vallado's near-earth
 * algorithm as dsgp4 computes it; with identical constants the state agrees with
 * that algorithm to 4e-12 km / 1e-14 km/s typically and 1e-8 km / 1e-11 km/s at
 * worst (tests/test_hostcheck.py; the tolerance of record is 1e-6 km). */
int ksl_sgp4_prop(const double *consts, int64_t n, const double *tsince, int64_t m, int per_sample_t,
                  double *rv_km, int32_t *err, void *stream);

/* the same, additionally ADDING to *n_exact (device, may be NULL) the number of evaluations that did not take the
 * fast path (m >= 32 only; the parity sweeps use it to report how much of the element space the fast path covers) */
int ksl_sgp4_prop_counted(const double *consts, int64_t n, const double *tsince, int64_t m, int per_sample_t,
                          double *rv_km, int32_t *err, unsigned long long *n_exact, void *stream);

/* ---- util.upsample(s, target_resolution)  (util.py:541-555) -------------------
 * x [n_in][ch] -> y [n_out][ch], torch's linear/align_corners=True arithmetic
 * including its float32 floor of the source index. */
int ksl_upsample(const double *x, int64_t n_in, int64_t ch, int64_t n_out, double *y, void *stream);

/* ---- util.propagate_upsample(tle, times_mjd, upsample_factor)  (util.py:557-580)
 * one initialised TLE (consts [KSL_NCONST], i.e. n = 1), its epoch as MJD,
 * times_sub = times_mjd[::upsample_factor] ([n_sub]), n_out = len(times_mjd);
 * out_m [n_out][6] in metres and m/s.  workspace: n_sub*6 doubles. */
int ksl_propagate_upsample(const double *consts, double epoch_mjd, const double *times_sub, int64_t n_sub,
                           int64_t n_out, double *out_m, double *workspace, int32_t *err, void *stream);

/* ---- model.find_conjunction(tr0, tr1, miss_dist_threshold)  (model.py:23-52) ---
 * tr0, tr1 [n][stride] with x,y,z in the first three columns.  out_i = {i_min,
 * i_conj (-1 = None)}, out_d = {d_min, d_conj (NaN = None)}.  torch semantics:
 * first index on ties, NaN counts as the minimum.  workspace: ksl_find_conjunction_ws(n) bytes. */
int64_t ksl_find_conjunction_ws(int64_t n);
int ksl_find_conjunction(const double *tr0, const double *tr1, int64_t n, int64_t stride, double thr,
                         int64_t *out_i, double *out_d, void *workspace, void *stream);

/* ---- the fused trace: space segment of Conjunction.forward()  (model.py:574-613)
 * For every sample s: propagate target and chaser at the n_coarse epochs
 * times_coarse[k] (MJD; tsince = (times_coarse[k] - epoch) * 1440, util.py:575),
 * up-sample linearly to n_fine points (util.py:541-555), and search the fine grid
 * for the global minimum distance and the first point below thr_m (model.py:23-52).
 *   consts_t SoA [KSL_NCONST][n_t], epoch_t [n_t], n_t = 1 (one target shared by
 *   all samples) or n;  consts_c SoA [KSL_NCONST][n], epoch_c [n].
 *   init_err_t [n_t] / init_err_c [n]: err arrays from ksl_sgp4_init (NULL = none);
 *   a deep-space record makes the trace a propagation error: i_min = -1.
 * Outputs (device, [n]): i_min, d_min [m], i_conj (-1 = None), d_conj [m] (NaN = None),
 * err (target bits 0..7, chaser bits 8..15).  The indices are those of the reference's
 * enumeration of all n_fine points (first index on ties, NaN counts as the minimum),
 * whichever search mode is used; samples are independent (any split or order of the
 * batch gives the same per-sample results).
 * workspace: ksl_screen_ws(n_t, n, n_coarse) bytes of device memory; required when n_t == 1 < n, optional
 * otherwise (with it, groups of traces are handed to warps through an atomic counter -- tail balance when
 * some traces cost more than others; without it, by static striding). */
int64_t ksl_screen_ws(int64_t n_t, int64_t n, int64_t n_coarse);
int ksl_screen(const double *consts_t, const double *epoch_t, const int32_t *init_err_t, int64_t n_t,
               const double *consts_c, const double *epoch_c, const int32_t *init_err_c, int64_t n,
               const double *times_coarse, int64_t n_coarse, int64_t n_fine, double thr_m, int mode,
               int64_t *i_min, double *d_min, int64_t *i_conj, double *d_conj, int32_t *err,
               void *workspace, void *stream);

/* The same, from the SGP4 INPUTS instead of initialised records: elems_t SoA [7][n_t], elems_c SoA [7][n]
 * (enum ksl_elem_index).  This is the fused form of model.py:574-613 -- dsgp4.initialize_tle (model.py:581,592)
 * included: from 65536 traces per call on, sgp4init runs inside the trace kernel (one lane per record, straight
 * into the warp's shared-memory copy of the record), so a trace reads its 2 x (7 + 1) doubles from HBM and nothing
 * else; smaller batches and a shared target go through ksl_sgp4_init into the workspace.  Results are bit-identical
 * to ksl_sgp4_init + ksl_screen (one body of machine code initialises records on both paths).
 * workspace: ksl_screen_elems_ws(n_t, n, n_coarse) bytes, required. */
int64_t ksl_screen_elems_ws(int64_t n_t, int64_t n, int64_t n_coarse);
int ksl_screen_elems(const double *elems_t, const double *epoch_t, int64_t n_t,
                     const double *elems_c, const double *epoch_c, int64_t n,
                     const double *times_coarse, int64_t n_coarse, int64_t n_fine, double thr_m, int mode,
                     int64_t *i_min, double *d_min, int64_t *i_conj, double *d_conj, int32_t *err,
                     void *workspace, void *stream);

/* counts (uint64 [KSL_SUM_NCOUNT]) and moments (double [KSL_SUM_NMOMENT]) over the
 * outputs of ksl_screen; deterministic (fixed reduction order); these are the
 * quantities all-reduced across GPUs.  workspace: ksl_screen_summary_ws(n) bytes. */
int64_t ksl_screen_summary_ws(int64_t n);
int ksl_screen_summary(const int64_t *i_conj, const double *d_min, const int32_t *err, int64_t n,
                       double collision_thr_m, uint64_t *counts, double *moments, void *workspace, void *stream);

/* ---- Monte Carlo covariance at TCA  (model.py:523-550) ------------------------
 * elems SoA [7][n] (already converted as model.py:500,525-533 do: mean motion
 * rad/min, e clamped at 0) -> sgp4init + sgp4 at one tsince -> states [n][6] in
 * metres (optional, NULL to skip) and the shifted moments about ref_state_m[6]:
 * moments[0] = n_ok, [1..6] = sum (x - ref), [7..27] = upper triangle (row-major)
 * of sum (x - ref)(x - ref)^T.  Deterministic.  The host finishes mean / RTN
 * rotation / np.cov(ddof=1).  err_count[0] = samples with any error bit.
 * workspace: ksl_tca_moments_ws(n) bytes. */
#define KSL_NMOMENT 28
int64_t ksl_tca_moments_ws(int64_t n);
int ksl_tca_moments(const double *elems, int64_t n, double tsince, const double *ref_state_m,
                    double *states_m, double *moments, int64_t *err_count, void *workspace, void *stream);

/* ---- Monte Carlo collision count  (model.py:432-437) --------------------------
 * t_xyz SoA [3][nt], c_xyz SoA [3][nc] (metres).  count = #pairs with
 * ((dx*dx + dy*dy) + dz*dz) < thr*thr evaluated without FMA contraction, so the
 * integer is bit-exact against the oracle for identical samples. *count is
 * ADDED to (caller zeroes it), which lets row shards accumulate. */
int ksl_pc_count(const double *t_xyz, int64_t nt, const double *c_xyz, int64_t nc, double thr_m, int pairing,
                 unsigned long long *count, void *stream);

/* ---- scale-out Monte Carlo collision probability (BASELINE configs[2]) -----------
 * The sample loop of model.py:497-541 for BOTH objects plus the miss-distance count of
 * model.py:432-437 in its elementwise form, with the draws generated on the device:
 * for global sample g = sample_offset + i, i < n:
 *   (a,e,i,raan,argp,M) = mean + L z,  z = 6 standard normals from Philox4x32-10 keyed by
 *   (seed, g, object) -- independent of how [0, N) is split over GPUs; the deviates are
 *   single-precision Box-Muller values (24 significant bits, tails to 6.8 sigma) widened to
 *   FP64, everything from the element vector on is FP64;
 *   a -> mean motion, e < 0 -> 0 (model.py:500,525-533); sgp4init + sgp4 at tsince; metres.
 * mean_*, ref_* [6] and chol_* [21] (row-major lower triangle of the Cholesky factor of the
 * element covariance, model.py:461-478) are HOST pointers (they travel as kernel arguments);
 * everything else is a device pointer.  *count / *err_count are ADDED to.
 * moments_t, moments_c [KSL_NMOMENT]: shifted moments of the two states about ref_*; moments_t
 * must have room for 2*KSL_NMOMENT doubles (moments_c may alias moments_t + KSL_NMOMENT).
 * dump (optional, [n_dump][26]): per sample the 7+7 SGP4 inputs and the 6+6 state components,
 * for parity checks against the oracle on IDENTICAL samples.  workspace: ksl_mc_pc_ws(n) bytes
 * (rows of partial moments + one mask word per 16 samples: a sample whose in-register fast path --
 * sgp4init in its fast form + the fast SGP4 state -- declines is marked there and redone on the
 * exact route by a second kernel of the same call; ksl_tca_moments works the same way).  The fast
 * path runs in its near-circular form when both clouds fit it (decided per call on the host from the
 * mean elements and six standard deviations of the eccentricity): a choice of speed, not of result. */
int64_t ksl_mc_pc_ws(int64_t n);
int ksl_mc_pc(const double *mean_t, const double *chol_t, double bstar_t, double tsince_t, const double *ref_t,
              const double *mean_c, const double *chol_c, double bstar_c, double tsince_c, const double *ref_c,
              uint64_t seed, int64_t sample_offset, int64_t n, double thr_m, double mu_m3s2, unsigned long long *count,
              double *moments_t, double *moments_c, unsigned long long *err_count, double *dump, int64_t n_dump,
              void *workspace, void *stream);

/* ---- dsgp4.newton_method(tle, time_mjd)  (model.py:399,411), batched ---------------
 * For each of n problems: the mean elements y = (no_kozai rad/min, e, i, raan, argp, M) whose SGP4 state at
 * tsince = 0 (with the given B*) equals target_km[6] (km, km/s, TEME).  resid = final ||F|| (km; km/s x 1000),
 * iters = Newton iterations used.  The caller builds the new TLE (epoch = observation time) from y. */
int ksl_newton_elements(const double *target_km, const double *bstar, int64_t n, int max_iter, double tol, double *y,
                        double *resid, int32_t *iters, void *stream);

/* ---- util.from_cartesian_to_keplerian + util.keplerian_cartesian_partials  (util.py:187-302), batched
 * states [n][6] (x,y,z,vx,vy,vz; units of mu) -> elements [n][6] = (a, e, i, Omega, omega, M) and
 * jac [n][6][6], jac[i][j] = d element_i / d state_j (exact: forward-mode dual numbers, no finite
 * differences).  NaN rows for e >= 1. */
int ksl_kepler_partials(const double *states, int64_t n, double mu, double *elements, double *jac, void *stream);

/* ---- batched ground segment helpers -------------------
 * ksl_states_at_fine: `t_states[i_cdm]` of model.py:658,670 for n_items (record, fine index) pairs without
 * materialising any trajectory: consts SoA [KSL_NCONST][n_rec]; rec[it] selects the record, epoch[it] its
 * epoch (MJD), fine_idx[it] the index on the n_fine grid; out_m [n_items][6] metres, m/s with util.upsample's
 * arithmetic; err [n_items] Vallado bits (may be NULL). */
int ksl_states_at_fine(const double *consts, int64_t n_rec, const int64_t *rec, const double *epoch, const int64_t *fine_idx,
                       int64_t n_items, const double *times_coarse, int64_t n_coarse, int64_t n_fine, double *out_m, int32_t *err,
                       void *stream);
/* ksl_pc_mvn_batched: model.py:424-437 for n_cdm CDMs at once.  mean [n_cdm][2][3] and chol [n_cdm][2][6]
 * (row-major lower triangle of the 3x3 Cholesky factor) describe the Gaussian position samples of target (0)
 * and chaser (1); n_pts samples each are drawn on the device (Philox keyed by (seed, cdm_id0 + cdm, object,
 * point)) and counts[cdm] = #{(i, j): ||t_i - c_j|| < thr_m}.  points_ws: n_cdm*2*(3*n_pts + 6) doubles; the first
 * n_cdm*2*3*n_pts are returned filled with the samples [n_cdm][2][3][n_pts] (a test can recount on identical
 * samples), the rest holds the clouds' bounding boxes, which let CDMs whose clouds are further apart than thr_m
 * contribute their exact 0 without a pair loop. */
int ksl_pc_mvn_batched(const double *mean, const double *chol, int64_t n_cdm, int64_t n_pts, uint64_t seed, int64_t cdm_id0,
                       double thr_m, unsigned long long *counts, double *points_ws, void *stream);

/* ---- prior sampling on the device  (model.py:245-319 draws + TLE(dict) quantisation) ---------
 * A prior over the eight sampled TLE fields -- order: mean_motion [rad/s], mean_anomaly, eccentricity,
 * inclination, argument_of_perigee, raan, mean_motion_first_derivative, b_star -- is packed on the HOST with
 * ksl_prior_pack (one call per field) into a buffer of ksl_prior_param_bytes() bytes, copied to the device by
 * the caller, and sampled by ksl_sample_prior.  kind: 0 uniform(lo, hi); 1 mixture of normals;
 * 2 mixture of normals truncated to [lo, hi], sampled by inverse CDF (util.py:134-160) with cdf_a/cdf_b =
 * Phi((lo-loc)/scale), Phi((hi-loc)/scale) per component.  Outputs (device): raw [8][n] the draws, elems
 * [7][n] the SGP4 inputs after the TLE-text quantisation, flag [n] = 1 where a decimal rounding was within
 * binary noise of a tie (re-quantise those on the host).  Draws depend only on (seed, sample_offset + i). */
int64_t ksl_prior_param_bytes(void);
int ksl_prior_pack(int field, int kind, int ncomp, double lo, double hi, const double *probs, const double *loc, const double *scale,
                   const double *cdf_a, const double *cdf_b, void *host_params);
int ksl_sample_prior(const void *params, uint64_t seed, int64_t sample_offset, int64_t n, double *raw, double *elems, int32_t *flag,
                     void *stream);

/* ---- end-to-end entry points: HOST buffers in, HOST buffers out ----------------
 * What a Kessler-side binding calls for a batch of traces.  Allocates (and caches
 * per device) its staging and scratch buffers; synchronous.  Batches of 2^18 samples
 * and more are streamed through in four chunks on two streams (a short first one,
 * whose upload nothing can hide, then three equal ones), so that the copies overlap
 * the kernels -- with pinned and with pageable host buffers; results do not depend
 * on the chunking.
 *   elems_t SoA [7][n_t], epoch_t [n_t], elems_c SoA [7][n], epoch_c [n],
 *   times_coarse [n_coarse]  -- all host.  Outputs host [n] (err may be NULL).
 *   summary_counts [KSL_SUM_NCOUNT] / summary_moments [KSL_SUM_NMOMENT] may be NULL. */
int ksl_screen_host(int device, const double *elems_t, const double *epoch_t, int64_t n_t,
                    const double *elems_c, const double *epoch_c, int64_t n,
                    const double *times_coarse, int64_t n_coarse, int64_t n_fine, double thr_m, int mode,
                    double collision_thr_m,
                    int64_t *i_min, double *d_min, int64_t *i_conj, double *d_conj, int32_t *err,
                    uint64_t *summary_counts, double *summary_moments);
/* frees the cached buffers of ksl_*_host (all devices) */
int ksl_host_cache_free(void);

/* ---- measurement helpers ------------------------------------------------------
 * dependent-free DFMA loop: blocks x threads x iters x 2*unroll flops; used by
 * bench.py to MEASURE the FP64 peak (MEASURED_PEAKS.json has no FP64 entry). */
int ksl_fp64_peak_kernel(int blocks, int threads, int64_t iters, double *sink, void *stream);
/* the same with a selectable shape: chains = n + 16 * mode, n in {1, 2, 4, 8} independent chains per thread, mode =
 * 0: a = fma(a, M, B) with compile-time constants (one register operand; the pipe's issue peak), 1: a = fma(a, m, b) with
 * per-thread registers (THREE 64-bit register operands -- what real code looks like), 2: fma with two register operands,
 * 3: DMUL, 4: DADD (n in {2, 8} for modes 2-4).  blocks x threads x iters x n instructions.  benchmarks/fp64_ilp.py sweeps
 * it: on B200 a DFMA with three register operands issues every third cycle, not every second (operand bandwidth). */
int ksl_fp64_chain_kernel(int blocks, int threads, int64_t iters, int chains, double *sink, void *stream);
/* number of kernel launches issued by this library since load (all threads) */
int64_t ksl_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* KESSLER_B200_H */
