// kessler_b200.cu -- sm_100a kernels + C ABI (include/kessler_b200.h) for Kessler's
// Monte Carlo conjunction hot path.  See DESIGN.md for the kernel inventory.
//
// Nothing on this path is a dense contraction, so there is no tensor-core / TMA
// code here on purpose: every kernel is FP64-vector-pipe bound (SGP4) or tiny.
// What is B200-specific is the shape: persistent grids sized from the SM count
// (one 512-thread CTA per SM, 16 warps at <= 128 registers), one warp per trace with
// the per-trace SGP4 constants, the time grid and a sincos table staged in shared
// memory (broadcast LDS operands feed the DFMA pipe), neighbour epochs and arg-min
// reductions by warp shuffle, the FP32 pipe used for the conservative pruning test.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "sgp4_device.cuh"
#include "screen_device.cuh"
#include "kepler_device.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define KSL_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) return fail(KSL_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

#define KSL_LAUNCH_CHECK()                                                                     \
    do {                                                                                       \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                    \
        cudaError_t e__ = cudaGetLastError();                                                  \
        if (e__ != cudaSuccess) return fail(KSL_E_CUDA, "kernel launch: %s", cudaGetErrorString(e__)); \
    } while (0)

struct DeviceInfo {
    int sms = 0;
    bool ok = false;
};
DeviceInfo g_dev[64];
std::mutex g_dev_mu;

int sm_count(int *out) {
    int dev = 0;
    KSL_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(KSL_E_CUDA, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (!g_dev[dev].ok) {
        KSL_CUDA(cudaDeviceGetAttribute(&g_dev[dev].sms, cudaDevAttrMultiProcessorCount, dev));
        g_dev[dev].ok = true;
    }
    *out = g_dev[dev].sms;
    return KSL_OK;
}

inline cudaStream_t S(void *s) { return reinterpret_cast<cudaStream_t>(s); }

using ksl::Best;
using ksl::best_init;
using ksl::best_merge;
using ksl::best_visit;
using ksl::kNoIndex;

// ------------------------------------------------------------------------------------
// K1  sgp4init: one thread per TLE record, SoA in, SoA out (all accesses coalesced)
// ------------------------------------------------------------------------------------
__device__ __noinline__ int sgp4_init_record(const double *el, double *rec);   // screen_kernels.cuh
__global__ void __launch_bounds__(128) k_sgp4_init(const double *__restrict__ elems, long long n,
                                                   double *__restrict__ consts, int *__restrict__ err) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double el[7], rec[KSL_NCONST];
#pragma unroll
    for (int k = 0; k < 7; ++k) el[k] = elems[k * n + i];
    const int e = sgp4_init_record(el, rec);
#pragma unroll
    for (int k = 0; k < KSL_NCONST; ++k) consts[(long long)k * n + i] = rec[k];
    if (err) err[i] = e;
}

// ------------------------------------------------------------------------------------
// K2  sgp4 over an (n records) x (m epochs) grid, one thread per (record, epoch); two shapes.
// k_sgp4_prop_tile (m >= 32, the propagate / propagate_upsample shape): a CTA owns 128 consecutive epochs of ONE
// record; the record is staged once in shared memory as a fast record (the 36 constants + the products the
// reference rebuilds per call), so every constant is a broadcast LDS operand, as in the screening kernel; the
// sincos table sits next to it.  k_sgp4_prop (any m; the propagate_batch shape: many records, few epochs each):
// epochs run along the warp, the constants are gathered through the read-only path.
// Both transpose the 6-double states of a warp through shared memory so that the stores are fully coalesced
// (48 B per evaluation is what bounds a state-writing kernel from the HBM side), choose the near-circular form of
// the fast path per evaluation, and can count the evaluations that took the exact route (n_exact, optional).
// ------------------------------------------------------------------------------------
constexpr int kPropThreads = 128;

__device__ __forceinline__ void store_states_coalesced(double (*stage)[32 * 6], int warp, int lane, long long first_eval_of_warp,
                                                       long long evals_total, double *__restrict__ rv) {
    __syncwarp();
    const long long warp_base = first_eval_of_warp * 6;
    const long long warp_valid = evals_total * 6 - warp_base;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const int o = q * 32 + lane;
        if (o < warp_valid) rv[warp_base + o] = stage[warp][o];
    }
}

constexpr int kPropTile = 2 * kPropThreads;     // epochs per tile: two per thread (two independent dependency chains in flight)
__global__ void __launch_bounds__(kPropThreads, 4) k_sgp4_prop_tile(const double *__restrict__ consts, long long n,
                                                                 const double *__restrict__ tsince, long long m, int per_sample_t,
                                                                 int tiles_per_record, long long n_tiles, double *__restrict__ rv,
                                                                 int *__restrict__ err, unsigned long long *__restrict__ n_exact) {
    __shared__ double stage[kPropThreads / 32][32 * 6];
    __shared__ double s_rec[ksl::KSL_NFAST];
    __shared__ double s_tsmall;      // |tsince| below which the staged record takes the near-circular form
    __shared__ __align__(16) double s_grid[2 * ksl::kGridPoints];
    // persistent CTAs: the sincos table is staged once per CTA, then the CTA walks its tiles (256 epochs of one record each)
    for (int q = threadIdx.x; q < 2 * ksl::kGridPoints; q += kPropThreads) s_grid[q] = ksl::kGridSinCos[q];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int nx = 0;
    long long staged = -1;
    auto ld = [&](int q) { return s_rec[q]; };
    // a CTA owns a CONTIGUOUS run of tiles: consecutive tiles belong to the same record, which is then staged once per
    // record and not once per tile (no CTA barrier in between)
    const long long per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
    const long long tile_end = ((long long)blockIdx.x + 1) * per_cta < n_tiles ? ((long long)blockIdx.x + 1) * per_cta : n_tiles;
    for (long long tile = (long long)blockIdx.x * per_cta; tile < tile_end; ++tile) {
        const long long i = tile / tiles_per_record;
        const int j0 = (int)(tile - i * tiles_per_record) * kPropTile;
        if (i != staged) {                       // (uniform over the CTA)
            __syncthreads();                     // the previous record's readers are done (and, the first time, the table is staged)
            if (threadIdx.x < KSL_NCONST) s_rec[threadIdx.x] = consts[(long long)threadIdx.x * n + i];
            __syncthreads();
            if (threadIdx.x == 0) {
                ksl::prepare_fast_record(s_rec, 1);
                s_tsmall = ksl::fast_record_small_e_tmax(ld);
            }
            __syncthreads();
            staged = i;
        }
        // thread -> epochs ja = j0 + tid and jb = ja + 128 (clamped for the reads; rows end at m)
        const long long ja = (long long)j0 + threadIdx.x, jb = ja + kPropThreads;
        const long long jac = ja < m ? ja : m - 1, jbc = jb < m ? jb : m - 1;
        const double ta = per_sample_t ? tsince[i * m + jac] : tsince[jac];
        const double tb = per_sample_t ? tsince[i * m + jbc] : tsince[jbc];
        // the near-circular form when every lane of the warp qualifies at both epochs (one record per warp: the lanes differ in
        // |t| only), so that the warp never runs both forms; both evaluations in one straight-line block
        const bool small = __all_sync(0xffffffffu, fmax(fabs(ta), fabs(tb)) < s_tsmall);
        double ra[3], va[3], rb[3], vb[3];
        bool ok_a, ok_b;
        if (small) {
            ok_a = ksl::sgp4_prop_fast_rv<true, true>(ld, ta, ra, va, s_grid);
            ok_b = ksl::sgp4_prop_fast_rv<true, true>(ld, tb, rb, vb, s_grid);
        } else {
            ok_a = ksl::sgp4_prop_fast_rv<true, false>(ld, ta, ra, va, s_grid);
            ok_b = ksl::sgp4_prop_fast_rv<true, false>(ld, tb, rb, vb, s_grid);
        }
        int e = 0;
        if (!ok_a) {
            double rx[3], vx[3];
            e |= ksl::sgp4_state_exact(ld, ta, rx, vx);
            nx += (ja < m) ? 1u : 0u;
            ra[0] = rx[0]; ra[1] = rx[1]; ra[2] = rx[2]; va[0] = vx[0]; va[1] = vx[1]; va[2] = vx[2];
        }
        if (!ok_b) {
            double rx[3], vx[3];
            e |= ksl::sgp4_state_exact(ld, tb, rx, vx);
            nx += (jb < m) ? 1u : 0u;
            rb[0] = rx[0]; rb[1] = rx[1]; rb[2] = rx[2]; vb[0] = vx[0]; vb[1] = vx[1]; vb[2] = vx[2];
        }
        // coalesced stores: each warp's 32 states of a half-tile are consecutive epochs of record i, transposed through smem
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            double *st = &stage[warp][lane * 6];
            const double *r = half ? rb : ra, *v = half ? vb : va;
            st[0] = r[0]; st[1] = r[1]; st[2] = r[2];
            st[3] = v[0]; st[4] = v[1]; st[5] = v[2];
            __syncwarp();
            const long long jw = (long long)j0 + half * kPropThreads + warp * 32;
            const long long valid = (m - jw) * 6;       // doubles of this record left from the warp's first epoch on (<= 0: none)
            double *dst = rv + (i * m + jw) * 6;
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const int o = q * 32 + lane;
                if (o < valid) dst[o] = stage[warp][o];
            }
            __syncwarp();
        }
        e = __reduce_or_sync(0xffffffffu, e);
        if (lane == 0 && err && e) atomicOr(err + i, e);
    }
    nx = __reduce_add_sync(0xffffffffu, nx);
    if (lane == 0 && n_exact && nx) atomicAdd(n_exact, (unsigned long long)nx);
}

__global__ void __launch_bounds__(kPropThreads) k_sgp4_prop(const double *__restrict__ consts, long long n,
                                                            const double *__restrict__ tsince, long long m,
                                                            int per_sample_t, double *__restrict__ rv,
                                                            int *__restrict__ err, unsigned long long *__restrict__ n_exact) {
    __shared__ double stage[kPropThreads / 32][32 * 6];
    const long long total = n * m;
    const long long idx = (long long)blockIdx.x * kPropThreads + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool active = idx < total;
    int e = 0;
    long long i = 0;
    if (active) {
        i = idx / m;
        const long long j = idx - i * m;
        const double t = per_sample_t ? tsince[idx] : tsince[j];
        double r[3], v[3];
        e = ksl::sgp4_state([&](int k) { return __ldg(consts + (long long)k * n + i); }, t, r, v, ksl::kGridSinCos);
        double *st = &stage[warp][lane * 6];
        st[0] = r[0]; st[1] = r[1]; st[2] = r[2];
        st[3] = v[0]; st[4] = v[1]; st[5] = v[2];
    }
    store_states_coalesced(stage, warp, lane, (long long)blockIdx.x * kPropThreads + warp * 32, total, rv);
    if (err && active && e) atomicOr(err + i, e);
    (void)n_exact;
}

// ------------------------------------------------------------------------------------
// util.upsample (+ optional km -> m of util.py:579)
// ------------------------------------------------------------------------------------
template <bool kToMetres>
__global__ void __launch_bounds__(256) k_upsample(const double *__restrict__ x, long long n_in, int ch,
                                                  long long n_out, double scale, double *__restrict__ y) {
    const long long total = n_out * ch;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long j = idx / ch;
        const int c = (int)(idx - j * ch);
        int64_t i0, i1;
        double w0, w1;
        ksl::upsample_index(j, scale, n_in, &i0, &i1, &w0, &w1);
        const double a = x[i0 * ch + c], b = x[i1 * ch + c];
        y[idx] = kToMetres ? ksl::lerp_m(w0, a, w1, b) : ksl::lerp_torch(w0, a, w1, b);
    }
}

#include "screen_kernels.cuh"   // sgp4_init_record, find_conjunction kernels, K3 (k_screen, k_screen_warp), k_target_positions

// one initialised TLE at n_sub epochs given as MJD: [n_sub][6] km, km/s (util.py:575-576)
__global__ void __launch_bounds__(128) k_prop_times(const double *__restrict__ consts, double epoch_mjd,
                                                    const double *__restrict__ times, long long n_sub,
                                                    double *__restrict__ rv, int *__restrict__ err) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_sub) return;
    double r[3], v[3];
    const int e = ksl::sgp4_state([&](int q) { return __ldg(consts + q); }, (times[k] - epoch_mjd) * 1440.0, r, v, ksl::kGridSinCos);
    double *o = rv + 6 * k;
    o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = v[0]; o[4] = v[1]; o[5] = v[2];
    if (e && err) atomicOr(err, e);
}

// ------------------------------------------------------------------------------------
// summary of a screen batch: exact integer counts + deterministic FP64 moments
// ------------------------------------------------------------------------------------
constexpr int kSumBlocks = 296;
__global__ void __launch_bounds__(256) k_summary_partial(const long long *__restrict__ i_conj, const double *__restrict__ d_min,
                                                         const int *__restrict__ err, long long n, double coll_thr,
                                                         unsigned long long *__restrict__ counts,
                                                         double *__restrict__ partial /*[grid][4]*/) {
    unsigned long long c[6] = {0, 0, 0, 0, 0, 0};
    double m0 = 0.0, m1 = 0.0, m2 = 0.0, mn = INFINITY;
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (long long)gridDim.x * blockDim.x) {
        const int e = err[s];
        const bool deep = ((e | (e >> KSL_ERR_CHASER_SHIFT)) & KSL_ERR_DEEPSPACE) != 0;
        const long long ic = i_conj[s];
        const double d = d_min[s];
        c[0] += 1;
        c[1] += deep;
        c[2] += (!deep && ic >= 0);
        c[3] += (!deep && ic > 0);
        c[4] += (!deep && d < coll_thr);
        c[5] += ((e & ~(KSL_ERR_DEEPSPACE | (KSL_ERR_DEEPSPACE << KSL_ERR_CHASER_SHIFT))) != 0);
        if (!deep && d == d) {
            m0 += 1.0;
            m1 += d;
            m2 += d * d;
            mn = fmin(mn, d);
        }
    }
    __shared__ unsigned long long sc[6];
    __shared__ double sm[8][4];
    if (threadIdx.x < 6) sc[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        unsigned long long v = c[q];
        for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sc[q], v);
    }
    for (int d = 16; d > 0; d >>= 1) {
        m0 += __shfl_down_sync(0xffffffffu, m0, d);
        m1 += __shfl_down_sync(0xffffffffu, m1, d);
        m2 += __shfl_down_sync(0xffffffffu, m2, d);
        mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, d));
    }
    if ((threadIdx.x & 31) == 0) {
        double *o = sm[threadIdx.x >> 5];
        o[0] = m0; o[1] = m1; o[2] = m2; o[3] = mn;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 0; q < 6; ++q)
            if (sc[q]) atomicAdd(counts + q, sc[q]);
        double a0 = 0, a1 = 0, a2 = 0, an = INFINITY;
        for (int w = 0; w < 8; ++w) { a0 += sm[w][0]; a1 += sm[w][1]; a2 += sm[w][2]; an = fmin(an, sm[w][3]); }
        double *o = partial + 4 * blockIdx.x;
        o[0] = a0; o[1] = a1; o[2] = a2; o[3] = an;
    }
}

__global__ void k_summary_final(const double *__restrict__ partial, int nparts, double *__restrict__ moments) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double a0 = 0, a1 = 0, a2 = 0, an = INFINITY;
    for (int b = 0; b < nparts; ++b) {
        a0 += partial[4 * b]; a1 += partial[4 * b + 1]; a2 += partial[4 * b + 2]; an = fmin(an, partial[4 * b + 3]);
    }
    moments[0] = a0; moments[1] = a1; moments[2] = a2; moments[3] = an;
}

// ------------------------------------------------------------------------------------
// K4  Monte Carlo states at TCA + shifted moments (model.py:536-550).  Thread per sample: sgp4init (fast form)
// and the fast SGP4 state, all in registers, no call and no stack.  A sample the fast path declines (eccentricity
// above 0.12, decayed, non-finite ...) is only MARKED -- one bit in a per-warp-iteration mask --, and the exact
// route (sgp4_init + sgp4_prop, out of line) is taken for the marked samples by a second, tiny kernel, so the hot
// kernel carries neither that code's registers nor its stack frame.  Both kernels reduce their 28 shifted moments
// in a fixed order into rows of `partial`; k_moments_final adds the rows in order: deterministic.
// ------------------------------------------------------------------------------------
constexpr int kTcaThreads = 128;

__device__ __forceinline__ void moments_add(double acc[KSL_NMOMENT], const double x[6], const double *ref) {
    double d[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) d[q] = x[q] - ref[q];
    acc[0] += 1.0;
    int o = 7;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        acc[1 + a] += d[a];
#pragma unroll
        for (int b = a; b < 6; ++b) { acc[o] = fma(d[a], d[b], acc[o]); ++o; }
    }
}
// the same on accumulators that live in shared memory, one column per thread (stride = threads per CTA: conflict-free)
template <int kStride>
__device__ __forceinline__ void moments_add_strided(double *acc, const double x[6], const double *ref) {
    double d[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) d[q] = x[q] - ref[q];
    acc[0] += 1.0;
    int o = 7;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        acc[(1 + a) * kStride] += d[a];
#pragma unroll
        for (int b = a; b < 6; ++b) { acc[o * kStride] = fma(d[a], d[b], acc[o * kStride]); ++o; }
    }
}
__device__ __forceinline__ bool finite6(const double x[6]) {
    bool f = true;
#pragma unroll
    for (int q = 0; q < 6; ++q) f = f && (fabs(x[q]) < INFINITY);
    return f;
}
// sum of acc[q] over the lanes selected by the xor-distances `first`, first/2, ... 1 ... (all lanes: first = 16)
template <int kFirst, int kLast>
__device__ __forceinline__ double lanes_sum(double v) {
#pragma unroll
    for (int d = kFirst; d >= kLast; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

__global__ void __launch_bounds__(kTcaThreads, 3) k_tca_moments(const double *__restrict__ elems, long long n, double tsince,
                                                                const double *__restrict__ ref, double *__restrict__ states,
                                                                double *__restrict__ partial /*[rows][KSL_NMOMENT]*/,
                                                                unsigned int *__restrict__ masks /*[ceil(n/32)], zeroed*/,
                                                                unsigned long long *__restrict__ err_count) {
    __shared__ __align__(16) double s_grid[2 * ksl::kGridPoints];
    __shared__ double s_ref[6];
    __shared__ double sm[kTcaThreads / 32][KSL_NMOMENT];
    for (int q = threadIdx.x; q < 2 * ksl::kGridPoints; q += kTcaThreads) s_grid[q] = ksl::kGridSinCos[q];
    if (threadIdx.x < 6) s_ref[threadIdx.x] = ref[threadIdx.x];
    __syncthreads();
    double acc[KSL_NMOMENT];
#pragma unroll
    for (int q = 0; q < KSL_NMOMENT; ++q) acc[q] = 0.0;
    unsigned int nerr = 0;
    const int lane = threadIdx.x & 31;
    const long long warps_total = (long long)gridDim.x * (kTcaThreads / 32);
    for (long long it = (long long)blockIdx.x * (kTcaThreads / 32) + (threadIdx.x >> 5); it * 32 < n; it += warps_total) {
        const long long i = it * 32 + lane;
        const bool active = i < n;
        double el[7], x[6];
#pragma unroll
        for (int k = 0; k < 7; ++k) el[k] = elems[k * n + (active ? i : 0)];
        int ierr;
        const bool ok = ksl::sgp4_init_state_fast(el, tsince, x, x + 3, s_grid, &ierr);
#pragma unroll
        for (int q = 0; q < 6; ++q) x[q] *= 1e3;  // model.py:541: *1e3 -> metres
        const unsigned int declined = __ballot_sync(0xffffffffu, active && !ok);
        if (declined && lane == 0) {
            masks[it] = declined;
            masks[-1] = 1u;          // "some sample needs the exact route" (the word in front of the masks)
        }
        if (active && ok) {
            if (states) {
#pragma unroll
                for (int q = 0; q < 6; ++q) states[6 * i + q] = x[q];
            }
            nerr += ierr ? 1u : 0u;
            moments_add(acc, x, s_ref);
        }
    }
#pragma unroll
    for (int q = 0; q < KSL_NMOMENT; ++q) {
        const double v = lanes_sum<16, 1>(acc[q]);
        if (lane == 0) sm[threadIdx.x >> 5][q] = v;
    }
    nerr = __reduce_add_sync(0xffffffffu, nerr);
    if (lane == 0 && nerr) atomicAdd(err_count, (unsigned long long)nerr);
    __syncthreads();
    if (threadIdx.x < KSL_NMOMENT) {
        double v = 0.0;
        for (int w = 0; w < kTcaThreads / 32; ++w) v += sm[w][threadIdx.x];
        partial[(long long)blockIdx.x * KSL_NMOMENT + threadIdx.x] = v;
    }
}

// the samples k_tca_moments marked: exact route.  One thread per mask word; a single CTA row of `partial` per block.
constexpr int kExactThreads = 128;
__global__ void __launch_bounds__(kExactThreads) k_tca_moments_exact(const double *__restrict__ elems, long long n, double tsince,
                                                                     const double *__restrict__ ref, double *__restrict__ states,
                                                                     double *__restrict__ partial, const unsigned int *__restrict__ masks,
                                                                     unsigned long long *__restrict__ err_count) {
    __shared__ double sm[kExactThreads / 32][KSL_NMOMENT];
    double acc[KSL_NMOMENT];
    for (int q = 0; q < KSL_NMOMENT; ++q) acc[q] = 0.0;
    double rf[6];
    for (int q = 0; q < 6; ++q) rf[q] = ref[q];
    unsigned long long nerr = 0;
    const long long n_it = (masks[-1] != 0u) ? (n + 31) / 32 : 0;     // nothing marked: only the (zero) partial row is written
    for (long long it = (long long)blockIdx.x * kExactThreads + threadIdx.x; it < n_it; it += (long long)gridDim.x * kExactThreads) {
        unsigned int m = masks[it];
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const long long i = it * 32 + b;
            double el[7], x[6];
            for (int k = 0; k < 7; ++k) el[k] = elems[k * n + i];
            const int e = ksl::sgp4_init_state_exact(el, tsince, x, x + 3);
            for (int q = 0; q < 6; ++q) x[q] *= 1e3;
            if (states)
                for (int q = 0; q < 6; ++q) states[6 * i + q] = x[q];
            nerr += e ? 1 : 0;
            if (finite6(x)) moments_add(acc, x, rf);
        }
    }
    const int lane = threadIdx.x & 31;
    for (int q = 0; q < KSL_NMOMENT; ++q) {
        const double v = lanes_sum<16, 1>(acc[q]);
        if (lane == 0) sm[threadIdx.x >> 5][q] = v;
    }
    for (int d = 16; d > 0; d >>= 1) nerr += __shfl_xor_sync(0xffffffffu, nerr, d);
    if (lane == 0 && nerr) atomicAdd(err_count, nerr);
    __syncthreads();
    if (threadIdx.x < KSL_NMOMENT) {
        double v = 0.0;
        for (int w = 0; w < kExactThreads / 32; ++w) v += sm[w][threadIdx.x];
        partial[(long long)blockIdx.x * KSL_NMOMENT + threadIdx.x] = v;
    }
}

// rows of partial moments -> their sum, in a fixed order: column q is handled by the 32 lanes of warp q (lane l adds
// rows l, l + 32, ... in order, then a fixed shuffle tree): deterministic, and 32 loads in flight per column
__global__ void k_moments_final(const double *__restrict__ partial, int nparts, int width, double *__restrict__ out) {
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= width) return;
    double v = 0.0;
    for (int b = lane; b < nparts; b += 32) v += partial[(long long)b * width + q];
    v = lanes_sum<16, 1>(v);
    if (lane == 0) out[q] = v;
}

// ------------------------------------------------------------------------------------
// K5  collision count (model.py:432-437): integer result, no FMA contraction
// ------------------------------------------------------------------------------------
constexpr int kPcThreads = 256;
constexpr int kPcTile = 1024;
__global__ void __launch_bounds__(kPcThreads) k_pc_all_pairs(const double *__restrict__ t, long long nt,
                                                             const double *__restrict__ c, long long nc, double thr2,
                                                             long long c_per_block, unsigned long long *__restrict__ count) {
    __shared__ double sx[kPcTile], sy[kPcTile], sz[kPcTile];
    __shared__ unsigned long long s_total;
    const long long i = (long long)blockIdx.x * kPcThreads + threadIdx.x;
    const bool have = i < nt;
    const double tx = have ? t[i] : 0.0, ty = have ? t[nt + i] : 0.0, tz = have ? t[2 * nt + i] : 0.0;
    const long long c_lo = (long long)blockIdx.y * c_per_block;
    const long long c_hi = (c_lo + c_per_block < nc) ? c_lo + c_per_block : nc;
    unsigned int local = 0;
    if (threadIdx.x == 0) s_total = 0;
    for (long long base = c_lo; base < c_hi; base += kPcTile) {
        const int m = (int)((c_hi - base < kPcTile) ? c_hi - base : kPcTile);
        __syncthreads();
        for (int q = threadIdx.x; q < m; q += kPcThreads) {
            sx[q] = c[base + q];
            sy[q] = c[nc + base + q];
            sz[q] = c[2 * nc + base + q];
        }
        __syncthreads();
        if (have) {
#pragma unroll 4
            for (int q = 0; q < m; ++q) {
                const double dx = ksl::add_rn(tx, -sx[q]), dy = ksl::add_rn(ty, -sy[q]), dz = ksl::add_rn(tz, -sz[q]);
                local += (ksl::sqdist3(dx, dy, dz) < thr2) ? 1u : 0u;
            }
        }
    }
    unsigned long long v = local;
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_total, v);
    __syncthreads();
    if (threadIdx.x == 0 && s_total) atomicAdd(count, s_total);
}

__global__ void __launch_bounds__(256) k_pc_elementwise(const double *__restrict__ t, long long nt,
                                                        const double *__restrict__ c, long long nc, long long n,
                                                        double thr2, unsigned long long *__restrict__ count) {
    unsigned long long local = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double dx = ksl::add_rn(t[i], -c[i]), dy = ksl::add_rn(t[nt + i], -c[nc + i]),
                     dz = ksl::add_rn(t[2 * nt + i], -c[2 * nc + i]);
        local += (ksl::sqdist3(dx, dy, dz) < thr2) ? 1u : 0u;
    }
    for (int d = 16; d > 0; d >>= 1) local += __shfl_down_sync(0xffffffffu, local, d);
    __shared__ unsigned long long s_total;
    if (threadIdx.x == 0) s_total = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&s_total, local);
    __syncthreads();
    if (threadIdx.x == 0 && s_total) atomicAdd(count, s_total);
}

// ------------------------------------------------------------------------------------
// K6  scale-out Monte Carlo collision probability (BASELINE configs[2]):
// for every global sample index g: draw (a, e, i, Omega, omega, M) ~ N(mean, L L^T) for target and
// chaser with a counter-based generator keyed by (seed, g, object) -- so the draws do not depend on how the
// samples are sharded over GPUs --, convert as model.py:500,525-533 do, sgp4init + sgp4 at the TCA,
// and reduce (i) the number of samples whose miss distance is below the threshold (exact integer)
// and (ii) the shifted moments of both states.
// Mapping: a PAIR of adjacent lanes per sample -- the even lane is the target, the odd lane the chaser.  Every
// lane runs the same instruction stream on its own object (parameters from a two-entry shared-memory table),
// keeps ONE set of 28 moment accumulators, and the two positions meet through three shuffles for the
// miss-distance test.  sgp4init in its fast form and the fast SGP4 state stay in registers; samples they decline
// are marked in a mask and redone on the exact route by k_mc_pc_exact (see K4).  96 B of moments per CTA and one
// atomic per CTA leave the SM.
// The standard normals are single-precision Box-Muller deviates (24 significant bits, tails to 6.8 sigma) widened to
// FP64: the FP32 / MUFU pipes idle in this kernel while the FP64 pipe is what bounds it; everything from the
// element vector on is FP64.
// ------------------------------------------------------------------------------------
struct McObject {
    double mean[6];   // a [m], e, i, raan, argp, M
    double chol[21];  // lower triangle of L, row-major: L00, L10, L11, L20, ...
    double bstar, tsince;
    double ref[6];    // reference state [m, m/s] for the shifted moments
};
constexpr int kMcObjDoubles = sizeof(McObject) / sizeof(double);   // 35

__device__ __forceinline__ void philox4x32_10(unsigned int c0, unsigned int c1, unsigned int c2, unsigned int c3,
                                              unsigned int k0, unsigned int k1, unsigned int out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 53-bit uniform in (0, 1) from two 32-bit words
__device__ __forceinline__ double u53(unsigned int a, unsigned int b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}

// one Box-Muller pair in single precision from two 32-bit words: radius from u1 = (w0 + 1/2) 2^-32 in (0, 1] (a float
// keeps the small values that make the tails exactly), angle 2 pi w1 2^-32 - pi in [-pi, pi].  log, sin and cos are the
// MUFU-based intrinsics: absolute error 4e-7 (2^-21.4) in the logarithm near u1 = 1 and in sin / cos -- a relative
// 1e-6 distortion of a random deviate, far below the Monte Carlo error of any statistic of 1e8 samples -- and the
// special-function unit they run on is otherwise idle here.  max(., 0) guards the radius against that error at u1 -> 1.
__device__ __forceinline__ void box_muller_f32(unsigned int w0, unsigned int w1, double &z0, double &z1) {
    const float u1 = fmaf(__uint2float_rn(w0), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float rad = sqrtf(fmaxf(-2.0f * __logf(u1), 0.0f));
    const float ang = fmaf(__uint2float_rn(w1), 1.4629180792671596e-09f, -3.14159265358979f);   // 2 pi 2^-32 w1 - pi
    z0 = (double)(rad * __cosf(ang));
    z1 = (double)(rad * __sinf(ang));
}

// six standard normals for (seed, sample g, stream): two Philox blocks, three Box-Muller pairs
__device__ __forceinline__ void normals6(unsigned long long seed, unsigned long long g, unsigned int stream, double z[6]) {
    unsigned int w[4], v[4];
    philox4x32_10((unsigned int)g, (unsigned int)(g >> 32), stream * 2u, 0x4b534c32u, (unsigned int)seed, (unsigned int)(seed >> 32), w);
    philox4x32_10((unsigned int)g, (unsigned int)(g >> 32), stream * 2u + 1u, 0x4b534c32u, (unsigned int)seed, (unsigned int)(seed >> 32), v);
    box_muller_f32(w[0], w[1], z[0], z[1]);
    box_muller_f32(w[2], w[3], z[2], z[3]);
    box_muller_f32(v[0], v[1], z[4], z[5]);
}

// element vector of one draw: q = mean + L z, then a -> mean motion [rad/min] and e < 0 -> 0 (model.py:500,525-533).
// `o` points at the object's 35 parameters (shared memory or a register copy); explicit roundings, so that the fast
// kernel and the exact-route kernel form bit-identical elements from the same draw.
__device__ __forceinline__ void mc_elements(const double *o, const double z[6], double mu, double el[7]) {
    double q[6];
    int idx = 6;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        double acc = o[r];
#pragma unroll
        for (int c = 0; c <= r; ++c) acc = ksl::fma_rn(o[idx++], z[c], acc);
        q[r] = acc;
    }
    const double a3 = ksl::mul_rn(ksl::mul_rn(q[0], q[0]), q[0]);
    el[0] = ksl::mul_rn(sqrt(ksl::mul_rn(mu, ksl::rcp_2(a3))), 60.0);
    el[1] = q[1] < 0.0 ? 0.0 : q[1];
    el[2] = q[2]; el[3] = q[3]; el[4] = q[4]; el[5] = q[5];
    el[6] = o[27];
}

#ifndef KSL_MC_SMEM_ACC
#define KSL_MC_SMEM_ACC 0      // 1: moment accumulators in shared memory (frees 56 registers per thread)
#endif
#ifndef KSL_MC_MIN_CTAS
#define KSL_MC_MIN_CTAS 3
#endif
constexpr int kMcThreads = 128;
// kSmallE: both objects' clouds sit inside the near-circular form of the fast path (mc_cloud_is_small_e, decided on the host
// from the mean elements and the spread of the eccentricity): ~25 FP64 instructions less per object-sample
template <bool kSmallE>
__global__ void __launch_bounds__(kMcThreads, KSL_MC_MIN_CTAS) k_mc_pc(const McObject tgt, const McObject chs, unsigned long long seed,
                                                         long long offset, long long n, double thr2, double mu,
                                                         unsigned long long *__restrict__ count,
                                                         double *__restrict__ partial /*[rows][2][KSL_NMOMENT]*/,
                                                         unsigned int *__restrict__ masks /*[ceil(n/16)], zeroed*/,
                                                         unsigned long long *__restrict__ nerr,
                                                         double *__restrict__ dump /*[n_dump][26] or NULL*/, long long n_dump) {
    __shared__ __align__(16) double s_grid[2 * ksl::kGridPoints];
    __shared__ double s_obj[2][kMcObjDoubles + 1];
    __shared__ double sm[kMcThreads / 32][2 * KSL_NMOMENT];
    __shared__ unsigned long long s_hits, s_errs;
    for (int q = threadIdx.x; q < 2 * ksl::kGridPoints; q += kMcThreads) s_grid[q] = ksl::kGridSinCos[q];
    if (threadIdx.x < kMcObjDoubles) {
        s_obj[0][threadIdx.x] = reinterpret_cast<const double *>(&tgt)[threadIdx.x];
        s_obj[1][threadIdx.x] = reinterpret_cast<const double *>(&chs)[threadIdx.x];
    }
    if (threadIdx.x == 0) { s_hits = 0; s_errs = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, obj = lane & 1;
    const double *o = s_obj[obj];
#if KSL_MC_SMEM_ACC
    __shared__ double s_acc[KSL_NMOMENT * kMcThreads];
    double *acc = s_acc + threadIdx.x;
#pragma unroll
    for (int q = 0; q < KSL_NMOMENT; ++q) acc[q * kMcThreads] = 0.0;
#else
    double acc[KSL_NMOMENT];
#pragma unroll
    for (int q = 0; q < KSL_NMOMENT; ++q) acc[q] = 0.0;
#endif
    unsigned int hits = 0, errs = 0;
    const long long warps_total = (long long)gridDim.x * (kMcThreads / 32);
    for (long long it = (long long)blockIdx.x * (kMcThreads / 32) + (threadIdx.x >> 5); it * 16 < n; it += warps_total) {
        const long long i = it * 16 + (lane >> 1);
        const bool active = i < n;
        const unsigned long long g = (unsigned long long)(offset + (active ? i : 0));
        double z[6], el[7], x[6];
        normals6(seed, g, (unsigned int)obj, z);
        mc_elements(o, z, mu, el);
        int ierr;
        const bool ok = ksl::sgp4_init_state_fast<kSmallE>(el, o[28], x, x + 3, s_grid, &ierr);
#pragma unroll
        for (int q = 0; q < 6; ++q) x[q] *= 1e3;
        // (the shuffle stands on its own line: behind `ok &&` the lanes whose own evaluation declined would skip a
        // full-mask shuffle -- undefined, and a hang with some code generations)
        const int ok_other = __shfl_xor_sync(0xffffffffu, ok ? 1 : 0, 1);
        const bool ok_pair = ok && (ok_other != 0);
        const unsigned int declined = __ballot_sync(0xffffffffu, active && !ok_pair);
        if (declined && lane == 0) {
            masks[it] = declined;
            masks[-1] = 1u;          // "some sample needs the exact route" (the word in front of the masks)
        }
        // miss distance: target minus chaser, in the oracle's arithmetic; both lanes of the pair hold it, the even one counts
        const double dx = ksl::add_rn(x[0], -__shfl_xor_sync(0xffffffffu, x[0], 1));
        const double dy = ksl::add_rn(x[1], -__shfl_xor_sync(0xffffffffu, x[1], 1));
        const double dz = ksl::add_rn(x[2], -__shfl_xor_sync(0xffffffffu, x[2], 1));
        const int e_pair = ierr | __shfl_xor_sync(0xffffffffu, ierr, 1);
        if (active && ok_pair) {
            if (obj == 0) {
                hits += (ksl::sqdist3(dx, dy, dz) < thr2) ? 1u : 0u;
                errs += e_pair ? 1u : 0u;
            }
#if KSL_MC_SMEM_ACC
            moments_add_strided<kMcThreads>(acc, x, o + 29);
#else
            moments_add(acc, x, o + 29);
#endif
            if (dump && i < n_dump) {
                double *d = dump + i * 26;
#pragma unroll
                for (int q = 0; q < 7; ++q) d[obj * 7 + q] = el[q];
#pragma unroll
                for (int q = 0; q < 6; ++q) d[14 + obj * 6 + q] = x[q];
            }
        }
    }
    // even lanes hold target moments, odd lanes chaser moments: reduce over the 16 lanes of each parity
#pragma unroll
    for (int q = 0; q < KSL_NMOMENT; ++q) {
#if KSL_MC_SMEM_ACC
        const double v = lanes_sum<16, 2>(acc[q * kMcThreads]);
#else
        const double v = lanes_sum<16, 2>(acc[q]);
#endif
        if (lane < 2) sm[threadIdx.x >> 5][lane * KSL_NMOMENT + q] = v;
    }
    hits = __reduce_add_sync(0xffffffffu, hits);
    errs = __reduce_add_sync(0xffffffffu, errs);
    if (lane == 0) {
        if (hits) atomicAdd(&s_hits, (unsigned long long)hits);
        if (errs) atomicAdd(&s_errs, (unsigned long long)errs);
    }
    __syncthreads();
    if (threadIdx.x < 2 * KSL_NMOMENT) {
        double v = 0.0;
        for (int w = 0; w < kMcThreads / 32; ++w) v += sm[w][threadIdx.x];
        partial[(long long)blockIdx.x * 2 * KSL_NMOMENT + threadIdx.x] = v;
    }
    if (threadIdx.x == 0) {
        if (s_hits) atomicAdd(count, s_hits);
        if (s_errs) atomicAdd(nerr, s_errs);
    }
}

// the samples k_mc_pc marked: both objects on the exact route (same draws, same elements).  One thread per mask word.
__global__ void __launch_bounds__(kExactThreads) k_mc_pc_exact(const McObject tgt, const McObject chs, unsigned long long seed,
                                                               long long offset, long long n, double thr2, double mu,
                                                               unsigned long long *__restrict__ count, double *__restrict__ partial,
                                                               const unsigned int *__restrict__ masks, unsigned long long *__restrict__ nerr,
                                                               double *__restrict__ dump, long long n_dump) {
    __shared__ double sm[kExactThreads / 32][2 * KSL_NMOMENT];
    double acc_t[KSL_NMOMENT], acc_c[KSL_NMOMENT];
    for (int q = 0; q < KSL_NMOMENT; ++q) { acc_t[q] = 0.0; acc_c[q] = 0.0; }
    unsigned long long hits = 0, errs = 0;
    const long long n_it = (masks[-1] != 0u) ? (n + 15) / 16 : 0;     // nothing marked: only the (zero) partial row is written
    for (long long it = (long long)blockIdx.x * kExactThreads + threadIdx.x; it < n_it; it += (long long)gridDim.x * kExactThreads) {
        unsigned int m = masks[it] & 0x55555555u;     // one bit per pair (both lanes of a declined pair voted)
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const long long i = it * 16 + (b >> 1);
            const unsigned long long g = (unsigned long long)(offset + i);
            double z[6], el_t[7], el_c[7], xt[6], xc[6];
            normals6(seed, g, 0u, z);
            mc_elements(reinterpret_cast<const double *>(&tgt), z, mu, el_t);
            const int e0 = ksl::sgp4_init_state_exact(el_t, tgt.tsince, xt, xt + 3);
            normals6(seed, g, 1u, z);
            mc_elements(reinterpret_cast<const double *>(&chs), z, mu, el_c);
            const int e1 = ksl::sgp4_init_state_exact(el_c, chs.tsince, xc, xc + 3);
            for (int q = 0; q < 6; ++q) { xt[q] *= 1e3; xc[q] *= 1e3; }
            const double dx = ksl::add_rn(xt[0], -xc[0]), dy = ksl::add_rn(xt[1], -xc[1]), dz = ksl::add_rn(xt[2], -xc[2]);
            hits += (ksl::sqdist3(dx, dy, dz) < thr2) ? 1u : 0u;
            errs += (e0 | e1) ? 1u : 0u;
            if (finite6(xt)) moments_add(acc_t, xt, tgt.ref);
            if (finite6(xc)) moments_add(acc_c, xc, chs.ref);
            if (dump && i < n_dump) {
                double *d = dump + i * 26;
                for (int q = 0; q < 7; ++q) { d[q] = el_t[q]; d[7 + q] = el_c[q]; }
                for (int q = 0; q < 6; ++q) { d[14 + q] = xt[q]; d[20 + q] = xc[q]; }
            }
        }
    }
    const int lane = threadIdx.x & 31;
    for (int q = 0; q < KSL_NMOMENT; ++q) {
        const double a = lanes_sum<16, 1>(acc_t[q]), b = lanes_sum<16, 1>(acc_c[q]);
        if (lane == 0) { sm[threadIdx.x >> 5][q] = a; sm[threadIdx.x >> 5][KSL_NMOMENT + q] = b; }
    }
    for (int d = 16; d > 0; d >>= 1) {
        hits += __shfl_xor_sync(0xffffffffu, hits, d);
        errs += __shfl_xor_sync(0xffffffffu, errs, d);
    }
    if (lane == 0) {
        if (hits) atomicAdd(count, hits);
        if (errs) atomicAdd(nerr, errs);
    }
    __syncthreads();
    if (threadIdx.x < 2 * KSL_NMOMENT) {
        double v = 0.0;
        for (int w = 0; w < kExactThreads / 32; ++w) v += sm[w][threadIdx.x];
        partial[(long long)blockIdx.x * 2 * KSL_NMOMENT + threadIdx.x] = v;
    }
}

// ------------------------------------------------------------------------------------
// K7  mean elements from an osculating state (dsgp4.newton_method as Kessler calls it, model.py:399,411):
// find (no_kozai, e, i, raan, argp, M) whose SGP4 state at tsince = 0 equals the target state.  One thread
// per problem; Newton in the non-singular unknowns z = (n, e sin w, e cos w, i, raan, w + M) with a central-
// difference Jacobian (12 SGP4 evaluations), a 6x6 elimination with partial pivoting and a halving line search
// (6 evaluations) per iteration -- everything in registers / local memory, no host round trip per iteration.
// K8  Kepler elements + exact Jacobian by forward-mode dual numbers (kepler_device.cuh).
// ------------------------------------------------------------------------------------
__device__ __noinline__ void newton_residual(const double z[6], double bstar, const double target[6], double F[6]) {
    double el[7];
    const double ecc = sqrt(z[1] * z[1] + z[2] * z[2]);
    const double w = atan2(z[1], z[2]);
    el[0] = z[0]; el[1] = ecc; el[2] = z[3]; el[3] = z[4]; el[4] = w; el[5] = z[5] - w; el[6] = bstar;
    double c[KSL_NCONST], x[6];
    ksl::sgp4_init(el, [&](int k, double v) { c[k] = v; });
    ksl::sgp4_prop<true>([&](int k) { return c[k]; }, 0.0, x, x + 3);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        F[q] = x[q] - target[q];
        F[3 + q] = (x[3 + q] - target[3 + q]) * 1.0e3;   // weigh km/s like km over ~1000 s
    }
}

__device__ __forceinline__ double norm6(const double F[6]) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 6; ++q) s += F[q] * F[q];
    return sqrt(s);
}

__global__ void __launch_bounds__(64) k_newton(const double *__restrict__ target_km, const double *__restrict__ bstar_in,
                                               long long n, int max_iter, double tol, double mu_km,
                                               double *__restrict__ y_out, double *__restrict__ resid, int *__restrict__ iters) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double target[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) target[q] = target_km[6 * i + q];
    const double bstar = bstar_in[i];
    double kep[6];
    const bool ok = ksl::cartesian_to_keplerian<double>(target, target + 3, mu_km, kep);
    double z[6] = {sqrt(mu_km / (kep[0] * kep[0] * kep[0])) * 60.0, kep[1] * sin(kep[4]), kep[1] * cos(kep[4]), kep[2], kep[3],
                   kep[4] + kep[5]};
    const double h[6] = {1e-9, 1e-8, 1e-8, 1e-8, 1e-8, 1e-8};
    double F[6], res = INFINITY;
    int it = 0;
    if (ok) {
        for (; it < max_iter; ++it) {
            newton_residual(z, bstar, target, F);
            res = norm6(F);
            if (!(res >= tol) || !(res < INFINITY)) break;
            double A[6][7];   // [J | -F]
            for (int col = 0; col < 6; ++col) {
                double zp[6], zm[6], Fp[6], Fm[6];
                for (int q = 0; q < 6; ++q) { zp[q] = z[q]; zm[q] = z[q]; }
                zp[col] += h[col];
                zm[col] -= h[col];
                newton_residual(zp, bstar, target, Fp);
                newton_residual(zm, bstar, target, Fm);
                for (int row = 0; row < 6; ++row) A[row][col] = (Fp[row] - Fm[row]) / (2.0 * h[col]);
            }
            for (int row = 0; row < 6; ++row) A[row][6] = -F[row];
            bool singular = false;
            for (int col = 0; col < 6; ++col) {   // elimination with partial pivoting
                int piv = col;
                for (int row = col + 1; row < 6; ++row)
                    if (fabs(A[row][col]) > fabs(A[piv][col])) piv = row;
                if (!(fabs(A[piv][col]) > 0.0)) { singular = true; break; }
                if (piv != col)
                    for (int q = col; q < 7; ++q) { const double tmp = A[col][q]; A[col][q] = A[piv][q]; A[piv][q] = tmp; }
                for (int row = col + 1; row < 6; ++row) {
                    const double m = A[row][col] / A[col][col];
                    for (int q = col; q < 7; ++q) A[row][q] -= m * A[col][q];
                }
            }
            if (singular) break;
            double dz[6];
            for (int row = 5; row >= 0; --row) {
                double acc = A[row][6];
                for (int q = row + 1; q < 6; ++q) acc -= A[row][q] * dz[q];
                dz[row] = acc / A[row][row];
            }
            double best_norm = INFINITY, frac = 1.0, best_frac = 0.0;
            for (int m = 0; m < 6; ++m, frac *= 0.5) {   // damped step
                double zc[6], Fc[6];
                for (int q = 0; q < 6; ++q) zc[q] = z[q] + frac * dz[q];
                newton_residual(zc, bstar, target, Fc);
                const double nn = norm6(Fc);
                if (nn < best_norm) { best_norm = nn; best_frac = frac; }
            }
            if (!(best_norm < res)) { ++it; break; }   // no progress: rounding floor
            for (int q = 0; q < 6; ++q) z[q] += best_frac * dz[q];
        }
        if (it == max_iter) {
            newton_residual(z, bstar, target, F);
            res = norm6(F);
        }
    }
    const double ecc = sqrt(z[1] * z[1] + z[2] * z[2]);
    const double w = atan2(z[1], z[2]);
    double *y = y_out + 6 * i;
    y[0] = z[0]; y[1] = ecc; y[2] = z[3]; y[3] = z[4]; y[4] = w; y[5] = z[5] - w;
    if (!ok) for (int q = 0; q < 6; ++q) y[q] = NAN;
    resid[i] = ok ? res : NAN;
    iters[i] = it;
}

__global__ void __launch_bounds__(128) k_kepler_partials(const double *__restrict__ states, long long n, double mu,
                                                         double *__restrict__ elements, double *__restrict__ jac) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double st[6], el[6], J[36];
#pragma unroll
    for (int q = 0; q < 6; ++q) st[q] = states[6 * i + q];
    const bool ok = ksl::kepler_elements_and_partials(st, mu, el, J);
    for (int q = 0; q < 6; ++q) elements[6 * i + q] = ok ? el[q] : NAN;
    for (int q = 0; q < 36; ++q) jac[36 * i + q] = ok ? J[q] : NAN;
}

// ------------------------------------------------------------------------------------
// Batched ground segment, helpers for many events at once.
// K9   states_at_fine: the rows `t_states[i_cdm]` that forward() hands to the observation models
//      (model.py:658,670) for many (record, fine index) pairs: both bracketing coarse SGP4 states and the
//      torch interpolation arithmetic of util.upsample, in metres.
// K10  collision probability of many CDMs (model.py:424-437): Gaussian position samples of target and chaser
//      drawn on the device (Philox keyed by (seed, CDM id, object, point)), then the all-pairs count.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_states_at_fine(const double *__restrict__ consts, long long n_rec,
                                                        const long long *__restrict__ rec, const double *__restrict__ epoch,
                                                        const long long *__restrict__ fine_idx, long long n_items,
                                                        const double *__restrict__ times_coarse, long long n_coarse, double scale,
                                                        double *__restrict__ out_m, int *__restrict__ err) {
    const long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= n_items) return;
    const long long r = rec[it];
    int64_t i0, i1;
    double w0, w1;
    ksl::upsample_index(fine_idx[it], scale, n_coarse, &i0, &i1, &w0, &w1);
    double a[6], b[6];
    auto ld = [&](int k) { return __ldg(consts + (long long)k * n_rec + r); };
    int e = ksl::sgp4_state(ld, (times_coarse[i0] - epoch[it]) * 1440.0, a, a + 3, ksl::kGridSinCos);
    e |= ksl::sgp4_state(ld, (times_coarse[i1] - epoch[it]) * 1440.0, b, b + 3, ksl::kGridSinCos);
#pragma unroll
    for (int q = 0; q < 6; ++q) out_m[6 * it + q] = ksl::lerp_m(w0, a[q], w1, b[q]);
    if (err) err[it] = e;
}

// points [n_cdm][2][3][n_pts] (SoA per object): p = mean + L z, L lower-triangular 3x3 (6 numbers, row-major)
__global__ void __launch_bounds__(256) k_mvn_points(const double *__restrict__ mean /*[n_cdm][2][3]*/,
                                                    const double *__restrict__ chol /*[n_cdm][2][6]*/, long long n_cdm,
                                                    long long n_pts, unsigned long long seed, long long cdm_id0,
                                                    double *__restrict__ pts) {
    const long long total = n_cdm * 2 * n_pts;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long cdm = idx / (2 * n_pts);
        const long long rem = idx - cdm * 2 * n_pts;
        const int obj = (int)(rem / n_pts);
        const long long p = rem - (long long)obj * n_pts;
        unsigned int w[4], w2[4];
        const unsigned long long g = (unsigned long long)(cdm_id0 + cdm);
        philox4x32_10((unsigned int)g, (unsigned int)(g >> 32), (unsigned int)p, 0x50630000u + (unsigned int)obj + (unsigned int)((p >> 32) << 8),
                      (unsigned int)seed, (unsigned int)(seed >> 32), w);
        philox4x32_10((unsigned int)g, (unsigned int)(g >> 32), (unsigned int)p, 0x50640000u + (unsigned int)obj + (unsigned int)((p >> 32) << 8),
                      (unsigned int)seed, (unsigned int)(seed >> 32), w2);
        double z[4], sn, cs;
        double rad = sqrt(-2.0 * log(u53(w[0], w[1])));
        sincospi(2.0 * u53(w[2], w[3]), &sn, &cs);
        z[0] = rad * cs; z[1] = rad * sn;
        rad = sqrt(-2.0 * log(u53(w2[0], w2[1])));
        sincospi(2.0 * u53(w2[2], w2[3]), &sn, &cs);
        z[2] = rad * cs;
        const double *m = mean + (cdm * 2 + obj) * 3;
        const double *L = chol + (cdm * 2 + obj) * 6;
        double *o = pts + ((cdm * 2 + obj) * 3) * n_pts + p;
        o[0] = m[0] + L[0] * z[0];
        o[n_pts] = m[1] + L[1] * z[0] + L[2] * z[1];
        o[2 * n_pts] = m[2] + L[3] * z[0] + L[4] * z[1] + L[5] * z[2];
    }
}

// axis-aligned bounding box of every point cloud: bbox [n_cdm][2][6] = (min x,y,z, max x,y,z); one CTA per cloud
__global__ void __launch_bounds__(256) k_points_bbox(const double *__restrict__ pts, long long n_pts, double *__restrict__ bbox) {
    const double *p = pts + (long long)blockIdx.x * 3 * n_pts;
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (long long i = threadIdx.x; i < n_pts; i += blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double v = p[a * n_pts + i];
            lo[a] = fmin(lo[a], v);
            hi[a] = fmax(hi[a], v);
        }
    }
    __shared__ double s[8][6];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int d = 16; d > 0; d >>= 1) {
            lo[a] = fmin(lo[a], __shfl_down_sync(0xffffffffu, lo[a], d));
            hi[a] = fmax(hi[a], __shfl_down_sync(0xffffffffu, hi[a], d));
        }
        if ((threadIdx.x & 31) == 0) { s[threadIdx.x >> 5][a] = lo[a]; s[threadIdx.x >> 5][3 + a] = hi[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = s[0][threadIdx.x];
        for (int w = 1; w < 8; ++w) v = threadIdx.x < 3 ? fmin(v, s[w][threadIdx.x]) : fmax(v, s[w][threadIdx.x]);
        bbox[(long long)blockIdx.x * 6 + threadIdx.x] = v;
    }
}

// all-pairs count per CDM: grid = (target tiles, n_cdm); counts[cdm] accumulated.  A CDM whose two clouds are
// further apart than the threshold box-to-box (most of them: the clouds sit at the two objects' own radii,
// model.py:426-427) contributes 0 without visiting a single pair -- the count stays exact.
__global__ void __launch_bounds__(kPcThreads) k_pc_all_pairs_batched(const double *__restrict__ pts, long long n_pts, double thr2,
                                                                     const double *__restrict__ bbox,
                                                                     unsigned long long *__restrict__ counts) {
    __shared__ double sx[kPcTile], sy[kPcTile], sz[kPcTile];
    __shared__ unsigned long long s_total;
    const long long cdm = blockIdx.y;
    {
        const double *bt = bbox + (cdm * 2 + 0) * 6, *bc = bbox + (cdm * 2 + 1) * 6;
        double gap2 = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double g = fmax(0.0, fmax(bt[a] - bc[3 + a], bc[a] - bt[3 + a]));
            gap2 += g * g;
        }
        if (gap2 > thr2 * (1.0 + 1e-9) + 1e-6) return;   // uniform over the CTA
    }
    const double *t = pts + (cdm * 2 + 0) * 3 * n_pts, *c = pts + (cdm * 2 + 1) * 3 * n_pts;
    const long long i = (long long)blockIdx.x * kPcThreads + threadIdx.x;
    const bool have = i < n_pts;
    const double tx = have ? t[i] : 0.0, ty = have ? t[n_pts + i] : 0.0, tz = have ? t[2 * n_pts + i] : 0.0;
    unsigned int local = 0;
    if (threadIdx.x == 0) s_total = 0;
    for (long long base = 0; base < n_pts; base += kPcTile) {
        const int m = (int)((n_pts - base < kPcTile) ? n_pts - base : kPcTile);
        __syncthreads();
        for (int q = threadIdx.x; q < m; q += kPcThreads) {
            sx[q] = c[base + q];
            sy[q] = c[n_pts + base + q];
            sz[q] = c[2 * n_pts + base + q];
        }
        __syncthreads();
        if (have) {
#pragma unroll 4
            for (int q = 0; q < m; ++q) {
                const double dx = ksl::add_rn(tx, -sx[q]), dy = ksl::add_rn(ty, -sy[q]), dz = ksl::add_rn(tz, -sz[q]);
                local += (ksl::sqdist3(dx, dy, dz) < thr2) ? 1u : 0u;
            }
        }
    }
    unsigned long long v = local;
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_total, v);
    __syncthreads();
    if (threadIdx.x == 0 && s_total) atomicAdd(counts + cdm, s_total);
}

// ------------------------------------------------------------------------------------
// K11  prior sampling on the device: the eight TLE fields that
// Conjunction.make_target / make_chaser draw (model.py:245-319) from the default prior's families --
// mixtures of truncated normals by inverse CDF (util.py:134-160), a mixture of normals, uniforms --
// followed by the TLE-text quantisation that TLE(dict) applies, so that the
// rejection loop needs no host RNG and no H2D copy.  Draws are keyed by (seed, global index, field).
// Entries whose decimal rounding is within binary noise of a tie (probability ~1e-13) are flagged and
// re-quantised on the host by the text formatter.
// ------------------------------------------------------------------------------------
constexpr int kPriorMaxComp = 16;
struct PriorField {
    int kind;    // 0 uniform(lo, hi); 1 mixture of normals; 2 mixture of truncated normals on [lo, hi]
    int ncomp;
    double lo, hi;
    double cum[kPriorMaxComp];     // cumulative mixture weights
    double loc[kPriorMaxComp], scale[kPriorMaxComp];
    double cdf_a[kPriorMaxComp], cdf_w[kPriorMaxComp];   // Phi(alpha), Phi(beta) - Phi(alpha)
};
// field order of `raw`: mean_motion [rad/s], mean_anomaly, eccentricity, inclination, argument_of_perigee, raan,
// mean_motion_first_derivative, b_star
constexpr int kPriorFields = 8;

__constant__ double kPow10[19] = {1e-9, 1e-8, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, 1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9};

__device__ __forceinline__ double round_decimals_dev(double x, double p10, bool &risky) {
    const double scaled = x * p10;
    const double frac = fabs(scaled - floor(scaled) - 0.5);
    if (!(frac > 1e-14 * fmax(1.0, fabs(scaled)))) risky = true;
    return rint(scaled) / p10;
}

__device__ __forceinline__ double quantize_exp_field_dev(double x, bool &risky) {
    if (x == 0.0 || !(fabs(x) < INFINITY)) return 0.0;
    const double ax = fabs(x);
    const int ex = (int)floor(log10(ax)) + 1;
    if (ex < -8 || ex > 8) { risky = true; return x; }
    const double scaled = ax / kPow10[ex + 9] * 1e5;
    const double mant = rint(scaled);
    const double frac = fabs(scaled - floor(scaled) - 0.5);
    if (frac < 1e-9 || mant >= 100000.0 || mant < 10000.0) risky = true;
    return copysign(1.0, x) * (mant / 1e5) * kPow10[ex + 9];
}

__global__ void __launch_bounds__(256) k_sample_prior(const PriorField *__restrict__ prior, unsigned long long seed, long long offset,
                                                      long long n, double *__restrict__ raw /*[8][n]*/, double *__restrict__ elems /*[7][n]*/,
                                                      int *__restrict__ flag) {
    __shared__ PriorField s_p[kPriorFields];
    for (int i = threadIdx.x; i < (int)(sizeof(PriorField) * kPriorFields / sizeof(int)); i += blockDim.x)
        reinterpret_cast<int *>(s_p)[i] = reinterpret_cast<const int *>(prior)[i];
    __syncthreads();
    const double deg = 3.141592653589793 / 180.0, twopi = 2.0 * 3.141592653589793;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long g = (unsigned long long)(offset + i);
        double v[kPriorFields];
#pragma unroll 1
        for (int f = 0; f < kPriorFields; ++f) {
            unsigned int w[4];
            philox4x32_10((unsigned int)g, (unsigned int)(g >> 32), (unsigned int)f, 0x50524931u, (unsigned int)seed,
                          (unsigned int)(seed >> 32), w);
            const double u_comp = u53(w[0], w[1]), u_val = u53(w[2], w[3]);
            const PriorField &p = s_p[f];
            if (p.kind == 0) {
                v[f] = p.lo + u_val * (p.hi - p.lo);
            } else {
                int c = 0;
                while (c < p.ncomp - 1 && u_comp >= p.cum[c]) ++c;
                if (p.kind == 1) {
                    v[f] = p.loc[c] + p.scale[c] * normcdfinv(u_val);
                } else {
                    double x = p.loc[c] + p.scale[c] * normcdfinv(p.cdf_a[c] + u_val * p.cdf_w[c]);
                    v[f] = fmin(fmax(x, p.lo), p.hi);   // the reference redraws outside the support; reachable only by rounding
                }
            }
            raw[(long long)f * n + i] = v[f];
        }
        // TLE(dict): every SGP4 input goes through the text fields (tle.py / model.quantize_sgp4_inputs)
        bool risky = false;
        const double rev = round_decimals_dev(v[0] * 86400.0 / twopi, 1e8, risky);
        double e = round_decimals_dev(v[2], 1e7, risky);
        if (e >= 1.0) e = 0.9999999;
        elems[0 * n + i] = rev * twopi / 86400.0 * 60.0;
        elems[1 * n + i] = e;
        elems[2 * n + i] = round_decimals_dev(v[3] / deg, 1e4, risky) * deg;
        elems[3 * n + i] = round_decimals_dev(v[5] / deg, 1e4, risky) * deg;
        elems[4 * n + i] = round_decimals_dev(v[4] / deg, 1e4, risky) * deg;
        elems[5 * n + i] = round_decimals_dev(v[1] / deg, 1e4, risky) * deg;
        elems[6 * n + i] = quantize_exp_field_dev(v[7], risky);
        flag[i] = risky ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------
// FP64 peak probe: 8 independent DFMA chains per thread
// ------------------------------------------------------------------------------------
__global__ void k_fp64_peak(long long iters, double *__restrict__ sink) {
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-9, a2 = a0 + 2e-9, a3 = a0 + 3e-9, a4 = a0 + 4e-9, a5 = a0 + 5e-9,
           a6 = a0 + 6e-9, a7 = a0 + 7e-9;
    const double m = 0.999999999, b = 1e-12;
    for (long long i = 0; i < iters; ++i) {
        a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
        a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123456.789) sink[0] = r;  // never true; keeps the chains alive
}

// the same with a selectable number of independent chains per thread: how much instruction-level parallelism the FP64
// pipe needs at a given number of resident warps (the screen kernel runs 16 warps per SM with two chains per warp)
// kMode -- where the operands of the chain's instruction come from:
//   0  a = fma(a, M, B)   M, B compile-time constants (constant-bank operands): one 64-bit REGISTER operand per DFMA
//   1  a = fma(a, m, b)   m, b per-thread registers that differ from chain to chain: THREE register operands (real code)
//   2  a = fma(a, m, B)   two register operands
//   3  a = a * m          DMUL, two register operands
//   4  a = a + b          DADD, two register operands
//   5  a = fma(a, u, b)   u warp-uniform (read from shared memory at a uniform address, so ptxas moves it to a UNIFORM
//                         register: DFMA R, R, UR, R), b a per-thread register
//   6  a = fma(a, m, u)   the uniform operand in the addend slot (DFMA R, R, R, UR)
template <int kChains, int kMode>
__global__ void k_fp64_chains(long long iters, double *__restrict__ sink) {
    double a[kChains], m[kChains], b[kChains];
    __shared__ double uni[2 * kChains];
    if (kMode >= 5) {
        if (threadIdx.x < 2 * kChains)
            uni[threadIdx.x] = sink[0] * 1e-300 + (threadIdx.x < kChains ? 0.999999999 - 7e-13 * threadIdx.x : 1e-12 + 3e-21 * threadIdx.x);
        __syncthreads();
    }
#pragma unroll
    for (int c = 0; c < kChains; ++c) {
        a[c] = 1.0 + threadIdx.x * 1e-9 + c * 1e-9;
        m[c] = kMode == 5 ? uni[c] : 0.999999999 - (threadIdx.x + 7 * c) * 1e-13;
        b[c] = kMode == 6 ? uni[kChains + c] : 1e-12 + (threadIdx.x + 3 * c) * 1e-21;
    }
    for (long long i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) {
            if (kMode == 0) a[c] = fma(a[c], 0.999999999, 1e-12);
            else if (kMode == 1 || kMode >= 5) a[c] = fma(a[c], m[c], b[c]);
            else if (kMode == 2) a[c] = fma(a[c], m[c], 1e-12);
            else if (kMode == 3) a[c] = __dmul_rn(a[c], m[c]);
            else a[c] = __dadd_rn(a[c], b[c]);
        }
    }
    double r = 0.0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) r += a[c];
    if (r == 123456.789) sink[0] = r;
}

}  // namespace

// ======================================================================================
// C ABI
// ======================================================================================
extern "C" {

int ksl_version(void) { return KSL_VERSION; }
int ksl_nconst(void) { return KSL_NCONST; }
const char *ksl_last_error(void) { return g_err; }
int64_t ksl_launch_count(void) { return (int64_t)g_launches.load(); }

int ksl_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int ksl_sgp4_init(const double *elems, int64_t n, double *consts, int32_t *err, void *stream) {
    if (n < 0 || (n > 0 && (!elems || !consts))) return fail(KSL_E_ARG, "ksl_sgp4_init: bad argument");
    if (n == 0) return KSL_OK;
    const long long blocks = (n + 127) / 128;
    k_sgp4_init<<<(unsigned)blocks, 128, 0, S(stream)>>>(elems, n, consts, err);
    KSL_LAUNCH_CHECK();
    return KSL_OK;
}

int ksl_sgp4_prop_counted(const double *consts, int64_t n, const double *tsince, int64_t m, int per_sample_t, double *rv_km,
                          int32_t *err, unsigned long long *n_exact, void *stream) {
    if (n < 0 || m < 0 || ((n > 0 && m > 0) && (!consts || !tsince || !rv_km)))
        return fail(KSL_E_ARG, "ksl_sgp4_prop: bad argument");
    if (n == 0 || m == 0) return KSL_OK;
    if (m >= 32) {       // a CTA per (record, 128 epochs): the record staged in shared memory
        const long long tiles = (m + kPropTile - 1) / kPropTile;
        if (tiles > 0x7fffffffLL) return fail(KSL_E_ARG, "ksl_sgp4_prop: m too large");
        int sms = 0;
        const int rc = sm_count(&sms);
        if (rc) return rc;
        const long long n_tiles = tiles * n;
        long long grid = (long long)sms * 4;            // resident CTAs of 128 threads per SM at <= 128 registers
        if (grid > n_tiles) grid = n_tiles;
        k_sgp4_prop_tile<<<(unsigned)grid, kPropThreads, 0, S(stream)>>>(consts, n, tsince, m, per_sample_t, (int)tiles, n_tiles, rv_km, err,
                                                                         n_exact);
    } else {
        const long long total = (long long)n * m;
        const long long blocks = (total + kPropThreads - 1) / kPropThreads;
        if (blocks > 0x7fffffffLL) return fail(KSL_E_ARG, "ksl_sgp4_prop: n*m too large for one launch");
        k_sgp4_prop<<<(unsigned)blocks, kPropThreads, 0, S(stream)>>>(consts, n, tsince, m, per_sample_t, rv_km, err, n_exact);
    }
    KSL_LAUNCH_CHECK();
    return KSL_OK;
}

int ksl_sgp4_prop(const double *consts, int64_t n, const double *tsince, int64_t m, int per_sample_t, double *rv_km,
                  int32_t *err, void *stream) {
    return ksl_sgp4_prop_counted(consts, n, tsince, m, per_sample_t, rv_km, err, nullptr, stream);
}

static double up_scale(int64_t n_in, int64_t n_out) { return n_out > 1 ? (double)(n_in - 1) / (double)(n_out - 1) : 0.0; }

int ksl_upsample(const double *x, int64_t n_in, int64_t ch, int64_t n_out, double *y, void *stream) {
    if (n_in <= 0 || ch <= 0 || n_out < 0 || !x || (n_out > 0 && !y)) return fail(KSL_E_ARG, "ksl_upsample: bad argument");
    if (n_out == 0) return KSL_OK;
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc) return rc;
    long long blocks = (n_out * ch + 255) / 256;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    k_upsample<false><<<(unsigned)blocks, 256, 0, S(stream)>>>(x, n_in, (int)ch, n_out, up_scale(n_in, n_out), y);
    KSL_LAUNCH_CHECK();
    return KSL_OK;
}

int64_t ksl_find_conjunction_ws(int64_t n) {
    (void)n;
    return (int64_t)sizeof(Best) * 1024;
}

int ksl_find_conjunction(const double *tr0, const double *tr1, int64_t n, int64_t stride, double thr, int64_t *out_i,
                         double *out_d, void *workspace, void *stream) {
    if (n <= 0 || stride < 3 || !tr0 || !tr1 || !out_i || !out_d || !workspace)
        return fail(KSL_E_ARG, "ksl_find_conjunction: bad argument");
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc) return rc;
    long long blocks = (n + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    if (blocks > (long long)sms * 4) blocks = (long long)sms * 4;
    Best *partial = reinterpret_cast<Best *>(workspace);
    k_findconj_partial<<<(unsigned)blocks, 256, 0, S(stream)>>>(tr0, tr1, n, stride, thr * thr, partial);
    KSL_LAUNCH_CHECK();
    k_findconj_final<<<1, 256, 0, S(stream)>>>(partial, (int)blocks, reinterpret_cast<long long *>(out_i), out_d);
    KSL_LAUNCH_CHECK();
    return KSL_OK;
}

}  // extern "C" (the screen launcher is C++)

namespace {
// workspace of the screen entry points: [target positions 3 x n_coarse][group counter][shared-target error flag]
// (+ for ksl_screen_elems below kFuseMinTraces: initialised records and flags of chasers and targets)
constexpr long long kFuseMinTraces = 65536;   // from here on sgp4init runs inside the screen kernel (one lane per record)
size_t screen_ws_core(int64_t n_coarse) { return sizeof(double) * 3 * (size_t)(n_coarse > 0 ? n_coarse : 1) + 256; }
size_t pad256(size_t b) { return ((b + 255) / 256) * 256; }

int screen_impl(const double *consts_t, const double *elems_t, const double *epoch_t, const int32_t *init_err_t, int64_t n_t,
                const double *consts_c, const double *elems_c, const double *epoch_c, const int32_t *init_err_c, int64_t n,
                const double *times_coarse, int64_t n_coarse, int64_t n_fine, double thr_m, int mode, int64_t *i_min,
                double *d_min, int64_t *i_conj, double *d_conj, int32_t *err, void *workspace, void *stream, const char *who) {
    if (n < 0 || n_coarse <= 0 || n_fine <= 0 || (n_t != 1 && n_t != n))
        return fail(KSL_E_ARG, "%s: need n >= 0, n_coarse > 0, n_fine > 0, n_t in {1, n}", who);
    const int force = mode & (KSL_SCREEN_FORCE_CTA | KSL_SCREEN_FORCE_WARP);
    mode &= ~(KSL_SCREEN_FORCE_CTA | KSL_SCREEN_FORCE_WARP);
    if (mode != KSL_SCREEN_ANALYTIC && mode != KSL_SCREEN_BRUTE) return fail(KSL_E_ARG, "%s: unknown mode %d", who, mode);
    if (n == 0) return KSL_OK;
    const bool from_elems_in = elems_c != nullptr;
    if ((!from_elems_in && (!consts_t || !consts_c)) || (from_elems_in && !elems_t) || !epoch_t || !epoch_c || !times_coarse ||
        !i_min || !d_min || !i_conj || !d_conj || !err)
        return fail(KSL_E_ARG, "%s: NULL pointer", who);
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc) return rc;
    const bool shared = (n_t == 1 && n > 1);
    // enough samples to give (three quarters of) the resident warps their own trace -> warp-per-trace kernel, else
    // CTA-per-trace (measured crossover on B200: ~1700 of 2368 warps; one warp needs 0.7 ms for a default trace, a CTA 0.18 ms)
    // (the warp kernel keeps 32-bit epoch / fine-grid indices)
    const bool fits32 = n_coarse < 0x40000000LL && n_fine < 0x7fffffffLL;   // (chunk bases advance by 32 in an int)
    const bool per_warp = fits32 && ((force & KSL_SCREEN_FORCE_WARP) ||
                                     (!(force & KSL_SCREEN_FORCE_CTA) && 4 * n >= 3LL * sms * KSL_SCREEN_MIN_CTAS * kWarpsPerCta));
    // sgp4init inside the warp kernel pays once a warp has whole groups of records to initialise (one lane each)
    const bool fuse_init = from_elems_in && per_warp && n >= kFuseMinTraces;
    char *ws = reinterpret_cast<char *>(workspace);
    if ((shared || (from_elems_in && !fuse_init)) && !ws) return fail(KSL_E_NOMEM, "%s: workspace required", who);
    double *pos = reinterpret_cast<double *>(ws);
    unsigned long long *counter = ws ? reinterpret_cast<unsigned long long *>(ws + sizeof(double) * 3 * (size_t)n_coarse) : nullptr;
    int *terr = ws ? reinterpret_cast<int *>(counter + 1) : nullptr;
    if (from_elems_in && (!fuse_init || shared)) {
        // records through k_sgp4_init: all of them below the fusing size, the shared target's single one always
        char *q = ws + screen_ws_core(n_coarse);
        double *ct = reinterpret_cast<double *>(q);
        q += pad256(sizeof(double) * KSL_NCONST * (size_t)n_t);
        int32_t *it = reinterpret_cast<int32_t *>(q);
        q += pad256(sizeof(int32_t) * (size_t)n_t);
        if (shared || !fuse_init) {
            rc = ksl_sgp4_init(elems_t, n_t, ct, it, stream);
            if (rc) return rc;
            consts_t = ct;
            init_err_t = it;
        }
        if (!fuse_init) {
            double *cc = reinterpret_cast<double *>(q);
            q += pad256(sizeof(double) * KSL_NCONST * (size_t)n);
            int32_t *ic = reinterpret_cast<int32_t *>(q);
            rc = ksl_sgp4_init(elems_c, n, cc, ic, stream);
            if (rc) return rc;
            consts_c = cc;
            init_err_c = ic;
        }
    }
    ScreenParams p;
    p.consts_t = consts_t; p.epoch_t = epoch_t; p.init_err_t = init_err_t; p.n_t = n_t;
    p.consts_c = consts_c; p.epoch_c = epoch_c; p.init_err_c = init_err_c; p.n = n;
    p.elems_t = elems_t; p.elems_c = elems_c;
    p.times_coarse = times_coarse; p.n_coarse = n_coarse; p.n_fine = n_fine;
    p.scale = up_scale(n_coarse, n_fine);
    p.thr2 = thr_m * thr_m;
    p.target_pos = nullptr; p.target_err = nullptr;
    p.i_min = reinterpret_cast<long long *>(i_min); p.i_conj = reinterpret_cast<long long *>(i_conj);
    p.d_min = d_min; p.d_conj = d_conj; p.err = err;
    p.counter = nullptr;
    p.group = 1;
    if (shared) {
        KSL_CUDA(cudaMemsetAsync(terr, 0, sizeof(int), S(stream)));
        k_target_positions<<<(unsigned)((n_coarse + 127) / 128), 128, 0, S(stream)>>>(consts_t, epoch_t, times_coarse, n_coarse,
                                                                                       pos, terr);
        KSL_LAUNCH_CHECK();
        p.target_pos = pos;
        p.target_err = terr;
    }
    const size_t dyn_grid = ((KSL_SCREEN_GRID_SMEM && n_coarse <= kMaxGridInSmem) ? sizeof(double) * (size_t)n_coarse : 0) +
                            ((n_coarse <= kMaxGridInSmem && n_fine < 0x7fffffffLL) ? sizeof(int) * (size_t)(n_coarse + 2) : 0);
    const size_t dyn = dyn_grid + (per_warp ? sizeof(double) * kWarpsPerCta * kGroupRecs * ksl::KSL_NFAST : 0);
    long long grid = (long long)sms * KSL_SCREEN_MIN_CTAS;
    if (!per_warp && grid > n) grid = n;
    if (per_warp) {
        // group size: as large as the per-warp record slice allows once every warp still gets >= 8 groups
        const long long warps_total = grid * kWarpsPerCta;
        long long g = n / (warps_total * 8);
        const long long gmax = shared ? kGroupRecs : kMaxGroup;
        p.group = (int)(g < 1 ? 1 : (g > gmax ? gmax : g));
        if (fuse_init && p.group < 4) p.group = 4;
        if (counter) {
            KSL_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), S(stream)));
            p.counter = counter;
        }
    }
#define KSL_SCREEN_LAUNCH(SH, MD)                                                                       \
    do {                                                                                                \
        if (per_warp && fuse_init) {                                                                    \
            KSL_CUDA(cudaFuncSetAttribute(k_screen_warp<SH, MD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)); \
            k_screen_warp<SH, MD, true><<<(unsigned)grid, kScreenThreads, dyn, S(stream)>>>(p);         \
        } else if (per_warp) {                                                                          \
            KSL_CUDA(cudaFuncSetAttribute(k_screen_warp<SH, MD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)); \
            k_screen_warp<SH, MD, false><<<(unsigned)grid, kScreenThreads, dyn, S(stream)>>>(p);        \
        } else {                                                                                        \
            KSL_CUDA(cudaFuncSetAttribute(k_screen<SH, MD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)); \
            k_screen<SH, MD><<<(unsigned)grid, kScreenThreads, dyn, S(stream)>>>(p);                    \
        }                                                                                               \
    } while (0)
    if (shared) {
        if (mode == KSL_SCREEN_ANALYTIC) KSL_SCREEN_LAUNCH(true, KSL_SCREEN_ANALYTIC);
        else KSL_SCREEN_LAUNCH(true, KSL_SCREEN_BRUTE);
    } else {
        if (mode == KSL_SCREEN_ANALYTIC) KSL_SCREEN_LAUNCH(false, KSL_SCREEN_ANALYTIC);
        else KSL_SCREEN_LAUNCH(false, KSL_SCREEN_BRUTE);
    }
#undef KSL_SCREEN_LAUNCH
    KSL_LAUNCH_CHECK();
    return KSL_OK;
}
}  // namespace

extern "C" {

int64_t ksl_screen_ws(int64_t n_t, int64_t n, int64_t n_coarse) {
    (void)n;
    (void)n_t;
    return (int64_t)screen_ws_core(n_coarse);
}

int ksl_screen(const double *consts_t, const double *epoch_t, const int32_t *init_err_t, int64_t n_t,
               const double *consts_c, const double *epoch_c, const int32_t *init_err_c, int64_t n,
               const double *times_coarse, int64_t n_coarse, int64_t n_fine, double thr_m, int mode, int64_t *i_min,
               double *d_min, int64_t *i_conj, double *d_conj, int32_t *err, void *workspace, void *stream) {
    return screen_impl(consts_t, nullptr, epoch_t, init_err_t, n_t, consts_c, nullptr, epoch_c, init_err_c, n, times_coarse, n_coarse,
                       n_fine, thr_m, mode, i_min, d_min, i_conj, d_conj, err, workspace, stream, "ksl_screen");
}

int64_t ksl_screen_elems_ws(int64_t n_t, int64_t n, int64_t n_coarse) {
    size_t b = screen_ws_core(n_coarse);
    const size_t nt = (size_t)(n_t > 0 ? n_t : 1), nn = (size_t)(n > 0 ? n : 1);
    b += pad256(sizeof(double) * KSL_NCONST * nt) + pad256(sizeof(int32_t) * nt);
    if (n < kFuseMinTraces) b += pad256(sizeof(double) * KSL_NCONST * nn) + pad256(sizeof(int32_t) * nn);
    return (int64_t)b;
}

int ksl_screen_elems(const double *elems_t, const double *epoch_t, int64_t n_t, const double *elems_c, const double *epoch_c,
                     int64_t n, const double *times_coarse, int64_t n_coarse, int64_t n_fine, double thr_m, int mode,
                     int64_t *i_min, double *d_min, int64_t *i_conj, double *d_conj, int32_t *err, void *workspace, void *stream) {
    if (n > 0 && (!elems_t || !elems_c)) return fail(KSL_E_ARG, "ksl_screen_elems: NULL pointer");
    return screen_impl(nullptr, elems_t, epoch_t, nullptr, n_t, nullptr, elems_c, epoch_c, nullptr, n, times_coarse, n_coarse, n_fine,
                       thr_m, mode, i_min, d_min, i_conj, d_conj, err, workspace, stream, "ksl_screen_elems");
}

int64_t ksl_screen_summary_ws(int64_t n) {
    (void)n;
    return (int64_t)sizeof(double) * 4 * kSumBlocks;
}

int ksl_screen_summary(const int64_t *i_conj, const double *d_min, const int32_t *err, int64_t n, double collision_thr_m,
                       uint64_t *counts, double *moments, void *workspace, void *stream) {
    if (n < 0 || !counts || !moments || !workspace || (n > 0 && (!i_conj || !d_min || !err)))
        return fail(KSL_E_ARG, "ksl_screen_summary: bad argument");
    KSL_CUDA(cudaMemsetAsync(counts, 0, sizeof(uint64_t) * KSL_SUM_NCOUNT, S(stream)));
    long long blocks = (n + 255) / 256;
    if (blocks > kSumBlocks) blocks = kSumBlocks;
    if (blocks < 1) blocks = 1;
    double *partial = reinterpret_cast<double *>(workspace);
    k_summary_partial<<<(unsigned)blocks, 256, 0, S(stream)>>>(reinterpret_cast<const long long *>(i_conj), d_min, err, n,
                                                               collision_thr_m,
                                                               reinterpret_cast<unsigned long long *>(counts), partial);
    KSL_LAUNCH_CHECK();
    k_summary_final<<<1, 32, 0, S(stream)>>>(partial, (int)blocks, moments);
    KSL_LAUNCH_CHECK();
    return KSL_OK;
}

// Monte Carlo kernels: rows of partial moments (hot kernel: <= kMcMaxRows CTAs, exact-route kernel: kMcExactRows) + masks
constexpr int kMcMaxRows = 1024, kMcExactRows = 64;

int64_t ksl_tca_moments_ws(int64_t n) {
    const size_t words = (size_t)((n > 0 ? n : 0) + 31) / 32 + 8;
    return (int64_t)(sizeof(double) * KSL_NMOMENT * (kMcMaxRows + kMcExactRows) + sizeof(unsigned int) * words + 256);
}

int ksl_tca_moments(const double *elems, int64_t n, double tsince, const double *ref_state_m, double *states_m,
                    double *moments, int64_t *err_count, void *workspace, void *stream) {
    if (n < 0 || !ref_state_m || !moments || !err_count || !workspace || (n > 0 && !elems))
        return fail(KSL_E_ARG, "ksl_tca_moments: bad argument");
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc) return rc;
    long long blocks = (n + kTcaThreads - 1) / kTcaThreads;
    if (blocks > (long long)sms * 3) blocks = (long long)sms * 3;
    if (blocks > kMcMaxRows) blocks = kMcMaxRows;
    if (blocks < 1) blocks = 1;
    const long long words = (n + 31) / 32;
    long long xblocks = (words + kExactThreads - 1) / kExactThreads;
    if (xblocks > kMcExactRows) xblocks = kMcExactRows;
    if (xblocks < 1) xblocks = 1;
    double *partial = reinterpret_cast<double *>(workspace);
    unsigned int *masks = reinterpret_cast<unsigned int *>(partial + (size_t)KSL_NMOMENT * (kMcMaxRows + kMcExactRows)) + 4;
    KSL_CUDA(cudaMemsetAsync(err_count, 0, sizeof(int64_t), S(stream)));
    KSL_CUDA(cudaMemsetAsync(masks - 4, 0, sizeof(unsigned int) * (size_t)(words + 4), S(stream)));
    k_tca_moments<<<(unsigned)blocks, kTcaThreads, 0, S(stream)>>>(elems, n, tsince, ref_state_m, states_m, partial, masks,
                                                                   reinterpret_cast<unsigned long long *>(err_count));
    KSL_LAUNCH_CHECK();
    k_tca_moments_exact<<<(unsigned)xblocks, kExactThreads, 0, S(stream)>>>(elems, n, tsince, ref_state_m, states_m,
                                                                            partial + (size_t)KSL_NMOMENT * blocks, masks,
                                                                            reinterpret_cast<unsigned long long *>(err_count));
    KSL_LAUNCH_CHECK();
    k_moments_final<<<1, 32 * KSL_NMOMENT, 0, S(stream)>>>(partial, (int)(blocks + xblocks), KSL_NMOMENT, moments);
    KSL_LAUNCH_CHECK();
    return KSL_OK;
}

int ksl_pc_count(const double *t_xyz, int64_t nt, const double *c_xyz, int64_t nc, double thr_m, int pairing,
                 unsigned long long *count, void *stream) {
    if (nt < 0 || nc < 0 || !count || ((nt > 0 && nc > 0) && (!t_xyz || !c_xyz)))
        return fail(KSL_E_ARG, "ksl_pc_count: bad argument");
    if (nt == 0 || nc == 0) return KSL_OK;
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc) return rc;
    const double thr2 = thr_m * thr_m;
    if (pairing == KSL_PAIR_ELEMENTWISE) {
        const long long n = nt < nc ? nt : nc;
        long long blocks = (n + 255) / 256;
        if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
        k_pc_elementwise<<<(unsigned)blocks, 256, 0, S(stream)>>>(t_xyz, nt, c_xyz, nc, n, thr2, count);
    } else if (pairing == KSL_PAIR_ALL) {
        const long long bx = (nt + kPcThreads - 1) / kPcThreads;
        if (bx > 0x7fffffffLL) return fail(KSL_E_ARG, "ksl_pc_count: nt too large");
        // split the chaser range so that the grid covers the SMs a few times over
        long long by = ((long long)sms * 4 + bx - 1) / bx;
        const long long max_by = (nc + kPcTile - 1) / kPcTile;
        if (by > max_by) by = max_by;
        if (by > 65535) by = 65535;
        if (by < 1) by = 1;
        long long per = (nc + by - 1) / by;
        per = ((per + kPcTile - 1) / kPcTile) * kPcTile;
        by = (nc + per - 1) / per;
        dim3 grid((unsigned)bx, (unsigned)by);
        k_pc_all_pairs<<<grid, kPcThreads, 0, S(stream)>>>(t_xyz, nt, c_xyz, nc, thr2, per, count);
    } else {
        return fail(KSL_E_ARG, "ksl_pc_count: unknown pairing %d", pairing);
    }
    KSL_LAUNCH_CHECK();
    return KSL_OK;
}

int ksl_propagate_upsample(const double *consts, double epoch_mjd, const double *times_sub, int64_t n_sub, int64_t n_out,
                           double *out_m, double *workspace, int32_t *err, void *stream) {
    if (n_sub <= 0 || n_out < 0 || !consts || !times_sub || !workspace || (n_out > 0 && !out_m))
        return fail(KSL_E_ARG, "ksl_propagate_upsample: bad argument");
    if (n_out == 0) return KSL_OK;
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc) return rc;
    k_prop_times<<<(unsigned)((n_sub + 127) / 128), 128, 0, S(stream)>>>(consts, epoch_mjd, times_sub, n_sub, workspace, err);
    KSL_LAUNCH_CHECK();
    long long blocks = (n_out * 6 + 255) / 256;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    k_upsample<true><<<(unsigned)blocks, 256, 0, S(stream)>>>(workspace, n_sub, 6, n_out, up_scale(n_sub, n_out), out_m);
    KSL_LAUNCH_CHECK();
    return KSL_OK;
}

int64_t ksl_mc_pc_ws(int64_t n) {
    const size_t words = (size_t)((n > 0 ? n : 0) + 15) / 16 + 8;
    return (int64_t)(sizeof(double) * 2 * KSL_NMOMENT * (kMcMaxRows + kMcExactRows) + sizeof(unsigned int) * words + 256);
}

// Does the whole cloud of one object fit the near-circular form of the fast path at its tsince?  The mean elements go through
// sgp4init on the HOST (the device code is __host__ __device__), the eccentricity is widened by six standard deviations
// of its marginal (row 1 of the Cholesky factor), and the record's own test decides.  A performance hint only: a sample
// outside the form is declined by the kernel and redone by the exact route.
struct HostRecStore {   // (functors rather than host lambdas: the callees are __host__ __device__ templates)
    double *rec;
    __host__ __device__ void operator()(int k, double v) const { rec[k] = v; }
};
struct HostRecLoad {
    const double *rec;
    __host__ __device__ double operator()(int k) const { return rec[k]; }
};
static bool mc_cloud_is_small_e(const McObject &o, double mu) {
    const double a = o.mean[0];
    if (!(a > 0.0) || !(mu > 0.0)) return false;
    double el[7] = {sqrt(mu / (a * a * a)) * 60.0, o.mean[1] < 0.0 ? 0.0 : o.mean[1], o.mean[2], o.mean[3], o.mean[4], o.mean[5], o.bstar};
    double rec[ksl::KSL_NFAST] = {0.0};
    if (ksl::sgp4_init(el, HostRecStore{rec}) != 0) return false;
    ksl::prepare_fast_record(rec, 1);
    // six standard deviations up for the form's upper bound, six down for its lower one (it declines em < 1e-6, where the
    // general form clamps: a cloud that reaches e = 0 -- draws below zero are clamped there -- belongs to the general form)
    const double spread = 6.0 * sqrt(o.chol[1] * o.chol[1] + o.chol[2] * o.chol[2]);
    rec[KSL_C_ECCO] = o.mean[1] - spread;
    if (!ksl::fast_record_is_small_e(HostRecLoad{rec}, fabs(o.tsince))) return false;
    rec[KSL_C_ECCO] = o.mean[1] + spread;
    return ksl::fast_record_is_small_e(HostRecLoad{rec}, fabs(o.tsince));
}

int ksl_mc_pc(const double *mean_t, const double *chol_t, double bstar_t, double tsince_t, const double *ref_t,
              const double *mean_c, const double *chol_c, double bstar_c, double tsince_c, const double *ref_c,
              uint64_t seed, int64_t sample_offset, int64_t n, double thr_m, double mu_m3s2, unsigned long long *count,
              double *moments_t, double *moments_c, unsigned long long *err_count, double *dump, int64_t n_dump,
              void *workspace, void *stream) {
    if (n < 0 || !mean_t || !chol_t || !ref_t || !mean_c || !chol_c || !ref_c || !count || !moments_t || !moments_c ||
        !err_count || !workspace || n_dump < 0 || (n_dump > 0 && !dump))
        return fail(KSL_E_ARG, "ksl_mc_pc: bad argument");
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc) return rc;
    McObject t, c;   // host pointers: the 2 x 35 parameters travel as kernel arguments
    for (int i = 0; i < 6; ++i) { t.mean[i] = mean_t[i]; c.mean[i] = mean_c[i]; t.ref[i] = ref_t[i]; c.ref[i] = ref_c[i]; }
    for (int i = 0; i < 21; ++i) { t.chol[i] = chol_t[i]; c.chol[i] = chol_c[i]; }
    t.bstar = bstar_t; t.tsince = tsince_t; c.bstar = bstar_c; c.tsince = tsince_c;
    long long blocks = (2 * n + kMcThreads - 1) / kMcThreads;     // two lanes per sample
    if (blocks > (long long)sms * KSL_MC_MIN_CTAS) blocks = (long long)sms * KSL_MC_MIN_CTAS;
    if (blocks > kMcMaxRows) blocks = kMcMaxRows;
    if (blocks < 1) blocks = 1;
    const long long words = (n + 15) / 16;
    long long xblocks = (words + kExactThreads - 1) / kExactThreads;
    if (xblocks > kMcExactRows) xblocks = kMcExactRows;
    if (xblocks < 1) xblocks = 1;
    double *partial = reinterpret_cast<double *>(workspace);
    unsigned int *masks = reinterpret_cast<unsigned int *>(partial + (size_t)2 * KSL_NMOMENT * (kMcMaxRows + kMcExactRows)) + 4;
    KSL_CUDA(cudaMemsetAsync(masks - 4, 0, sizeof(unsigned int) * (size_t)(words + 4), S(stream)));
    if (mc_cloud_is_small_e(t, mu_m3s2) && mc_cloud_is_small_e(c, mu_m3s2))
        k_mc_pc<true><<<(unsigned)blocks, kMcThreads, 0, S(stream)>>>(t, c, (unsigned long long)seed, sample_offset, n, thr_m * thr_m,
                                                                       mu_m3s2, count, partial, masks, err_count, dump, n_dump);
    else
        k_mc_pc<false><<<(unsigned)blocks, kMcThreads, 0, S(stream)>>>(t, c, (unsigned long long)seed, sample_offset, n, thr_m * thr_m,
                                                                        mu_m3s2, count, partial, masks, err_count, dump, n_dump);
    KSL_LAUNCH_CHECK();
    k_mc_pc_exact<<<(unsigned)xblocks, kExactThreads, 0, S(stream)>>>(t, c, (unsigned long long)seed, sample_offset, n, thr_m * thr_m,
                                                                      mu_m3s2, count, partial + (size_t)2 * KSL_NMOMENT * blocks, masks,
                                                                      err_count, dump, n_dump);
    KSL_LAUNCH_CHECK();
    k_moments_final<<<2, 32 * KSL_NMOMENT, 0, S(stream)>>>(partial, (int)(blocks + xblocks), 2 * KSL_NMOMENT, moments_t);
    KSL_LAUNCH_CHECK();
    if (moments_c != moments_t + KSL_NMOMENT)
        KSL_CUDA(cudaMemcpyAsync(moments_c, moments_t + KSL_NMOMENT, sizeof(double) * KSL_NMOMENT, cudaMemcpyDeviceToDevice, S(stream)));
    return KSL_OK;
}

int ksl_newton_elements(const double *target_km, const double *bstar, int64_t n, int max_iter, double tol, double *y,
                        double *resid, int32_t *iters, void *stream) {
    if (n < 0 || max_iter < 0 || (n > 0 && (!target_km || !bstar || !y || !resid || !iters)))
        return fail(KSL_E_ARG, "ksl_newton_elements: bad argument");
    if (n == 0) return KSL_OK;
    k_newton<<<(unsigned)((n + 63) / 64), 64, 0, S(stream)>>>(target_km, bstar, n, max_iter, tol, 398600.5, y, resid, iters);
    KSL_LAUNCH_CHECK();
    return KSL_OK;
}

int ksl_kepler_partials(const double *states, int64_t n, double mu, double *elements, double *jac, void *stream) {
    if (n < 0 || (n > 0 && (!states || !elements || !jac))) return fail(KSL_E_ARG, "ksl_kepler_partials: bad argument");
    if (n == 0) return KSL_OK;
    k_kepler_partials<<<(unsigned)((n + 127) / 128), 128, 0, S(stream)>>>(states, n, mu, elements, jac);
    KSL_LAUNCH_CHECK();
    return KSL_OK;
}

int ksl_states_at_fine(const double *consts, int64_t n_rec, const int64_t *rec, const double *epoch, const int64_t *fine_idx,
                       int64_t n_items, const double *times_coarse, int64_t n_coarse, int64_t n_fine, double *out_m, int32_t *err,
                       void *stream) {
    if (n_items < 0 || n_rec <= 0 || n_coarse <= 0 || n_fine <= 0 ||
        (n_items > 0 && (!consts || !rec || !epoch || !fine_idx || !times_coarse || !out_m)))
        return fail(KSL_E_ARG, "ksl_states_at_fine: bad argument");
    if (n_items == 0) return KSL_OK;
    k_states_at_fine<<<(unsigned)((n_items + 127) / 128), 128, 0, S(stream)>>>(consts, n_rec, reinterpret_cast<const long long *>(rec), epoch,
                                                                              reinterpret_cast<const long long *>(fine_idx), n_items,
                                                                              times_coarse, n_coarse, up_scale(n_coarse, n_fine), out_m, err);
    KSL_LAUNCH_CHECK();
    return KSL_OK;
}

int ksl_pc_mvn_batched(const double *mean, const double *chol, int64_t n_cdm, int64_t n_pts, uint64_t seed, int64_t cdm_id0,
                       double thr_m, unsigned long long *counts, double *points_ws, void *stream) {
    if (n_cdm < 0 || n_pts <= 0 || (n_cdm > 0 && (!mean || !chol || !counts || !points_ws)))
        return fail(KSL_E_ARG, "ksl_pc_mvn_batched: bad argument");
    if (n_cdm == 0) return KSL_OK;
    if (n_cdm > 32767) return fail(KSL_E_ARG, "ksl_pc_mvn_batched: at most 32767 CDMs per call");
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc) return rc;
    const long long total = n_cdm * 2 * n_pts;
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    k_mvn_points<<<(unsigned)blocks, 256, 0, S(stream)>>>(mean, chol, n_cdm, n_pts, (unsigned long long)seed, cdm_id0, points_ws);
    KSL_LAUNCH_CHECK();
    KSL_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * (size_t)n_cdm, S(stream)));
    double *bbox = points_ws + (size_t)n_cdm * 2 * 3 * (size_t)n_pts;
    k_points_bbox<<<(unsigned)(n_cdm * 2), 256, 0, S(stream)>>>(points_ws, n_pts, bbox);
    KSL_LAUNCH_CHECK();
    dim3 grid((unsigned)((n_pts + kPcThreads - 1) / kPcThreads), (unsigned)n_cdm);
    k_pc_all_pairs_batched<<<grid, kPcThreads, 0, S(stream)>>>(points_ws, n_pts, thr_m * thr_m, bbox, counts);
    KSL_LAUNCH_CHECK();
    return KSL_OK;
}

int64_t ksl_prior_param_bytes(void) { return (int64_t)(sizeof(PriorField) * kPriorFields); }

int ksl_prior_pack(int field, int kind, int ncomp, double lo, double hi, const double *probs, const double *loc, const double *scale,
                   const double *cdf_a, const double *cdf_b, void *host_params) {
    if (field < 0 || field >= kPriorFields || kind < 0 || kind > 2 || ncomp < 0 || ncomp > kPriorMaxComp || !host_params ||
        (kind > 0 && (ncomp < 1 || !probs || !loc || !scale)) || (kind == 2 && (!cdf_a || !cdf_b)))
        return fail(KSL_E_ARG, "ksl_prior_pack: bad argument");
    PriorField &p = reinterpret_cast<PriorField *>(host_params)[field];
    memset(&p, 0, sizeof(p));
    p.kind = kind; p.ncomp = ncomp; p.lo = lo; p.hi = hi;
    double acc = 0.0, tot = 0.0;
    for (int c = 0; c < ncomp; ++c) tot += probs[c];
    for (int c = 0; c < ncomp; ++c) {
        acc += probs[c] / tot;
        p.cum[c] = acc;
        p.loc[c] = loc[c];
        p.scale[c] = scale[c];
        if (kind == 2) { p.cdf_a[c] = cdf_a[c]; p.cdf_w[c] = cdf_b[c] - cdf_a[c]; }
    }
    return KSL_OK;
}

int ksl_sample_prior(const void *params, uint64_t seed, int64_t sample_offset, int64_t n, double *raw, double *elems, int32_t *flag,
                     void *stream) {
    if (n < 0 || !params || (n > 0 && (!raw || !elems || !flag))) return fail(KSL_E_ARG, "ksl_sample_prior: bad argument");
    if (n == 0) return KSL_OK;
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc) return rc;
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
    k_sample_prior<<<(unsigned)blocks, 256, 0, S(stream)>>>(reinterpret_cast<const PriorField *>(params), (unsigned long long)seed,
                                                            sample_offset, n, raw, elems, flag);
    KSL_LAUNCH_CHECK();
    return KSL_OK;
}

int ksl_fp64_chain_kernel(int blocks, int threads, int64_t iters, int chains, double *sink, void *stream) {
    if (blocks <= 0 || threads <= 0 || threads > 1024 || iters < 0 || !sink) return fail(KSL_E_ARG, "ksl_fp64_chain_kernel: bad argument");
    const int mode = chains / 16, nch = chains % 16;
#define KSL_CHAIN_CASE(C, M) \
    if (nch == C && mode == M) { k_fp64_chains<C, M><<<blocks, threads, 0, S(stream)>>>(iters, sink); launched = true; }
    bool launched = false;
    KSL_CHAIN_CASE(1, 0) KSL_CHAIN_CASE(2, 0) KSL_CHAIN_CASE(4, 0) KSL_CHAIN_CASE(8, 0)
    KSL_CHAIN_CASE(1, 1) KSL_CHAIN_CASE(2, 1) KSL_CHAIN_CASE(4, 1) KSL_CHAIN_CASE(8, 1)
    KSL_CHAIN_CASE(2, 2) KSL_CHAIN_CASE(8, 2) KSL_CHAIN_CASE(2, 3) KSL_CHAIN_CASE(8, 3) KSL_CHAIN_CASE(2, 4) KSL_CHAIN_CASE(8, 4)
    KSL_CHAIN_CASE(2, 5) KSL_CHAIN_CASE(8, 5) KSL_CHAIN_CASE(2, 6) KSL_CHAIN_CASE(8, 6)
#undef KSL_CHAIN_CASE
    if (!launched) return fail(KSL_E_ARG, "ksl_fp64_chain_kernel: unsupported chains / mode combination %d", chains);
    KSL_LAUNCH_CHECK();
    return KSL_OK;
}

int ksl_fp64_peak_kernel(int blocks, int threads, int64_t iters, double *sink, void *stream) {
    if (blocks <= 0 || threads <= 0 || threads > 1024 || iters < 0 || !sink) return fail(KSL_E_ARG, "ksl_fp64_peak_kernel: bad argument");
    k_fp64_peak<<<blocks, threads, 0, S(stream)>>>(iters, sink);
    KSL_LAUNCH_CHECK();
    return KSL_OK;
}

}  // extern "C"

// ---- end-to-end entry point: host buffers in, host buffers out ----------------------
namespace {
struct HostCache {
    std::mutex mu;            // one call at a time PER DEVICE (the arena and the two streams are per device);
                              // calls on different devices run concurrently
    char *arena = nullptr;
    size_t cap = 0;
    cudaStream_t stream = nullptr, stream2 = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
};
HostCache g_cache[64];

struct Bump {
    char *base;
    size_t off = 0;
    template <typename T>
    T *take(size_t count) {
        T *p = reinterpret_cast<T *>(base + off);
        off += ((count * sizeof(T) + 255) / 256) * 256;
        return p;
    }
};

// the body of ksl_screen_host once the device is selected and its cache locked; any error leaves work in flight,
// the caller drains both streams before it returns
int screen_host_locked(HostCache &hc, const double *elems_t, const double *epoch_t, int64_t n_t, const double *elems_c,
                       const double *epoch_c, int64_t n, const double *times_coarse, int64_t n_coarse, int64_t n_fine,
                       double thr_m, int mode, double collision_thr_m, int64_t *i_min, double *d_min, int64_t *i_conj,
                       double *d_conj, int32_t *err, uint64_t *summary_counts, double *summary_moments) {
    if (!hc.stream) {
        KSL_CUDA(cudaStreamCreateWithFlags(&hc.stream, cudaStreamNonBlocking));
        KSL_CUDA(cudaStreamCreateWithFlags(&hc.stream2, cudaStreamNonBlocking));
        KSL_CUDA(cudaEventCreateWithFlags(&hc.ev_ready, cudaEventDisableTiming));
        KSL_CUDA(cudaEventCreateWithFlags(&hc.ev_done, cudaEventDisableTiming));
    }
    // Large batches go through in chunks on two streams: the host->device copy of chunk j+1 and the device->host copy of
    // chunk j-1 overlap the kernels of chunk j (samples are independent, so chunking cannot change a result).  sgp4init
    // runs inside the screen kernel, so what is staged per chunk is 2 x 8 doubles per trace and nothing else.
    const bool shared = (n_t == 1);
    const int nchunk = (n >= (1 << 18)) ? 4 : 1;
    const size_t un = (size_t)n, unc = (size_t)n_coarse;
    // chunk boundaries: a SHORT first chunk (1/16 of the batch) -- its host->device copy is the one transfer nothing can
    // hide --, the rest in three equal parts
    size_t bound[5] = {0, un, un, un, un};
    if (nchunk > 1) {
        const size_t first = un / 16, per = (un - first + 2) / 3;
        for (int j = 1; j < nchunk; ++j) bound[j] = first + (size_t)(j - 1) * per < un ? first + (size_t)(j - 1) * per : un;
    }
    size_t m = 0, ws_bytes = 0;                                            // the longest chunk; the largest workspace any chunk needs
    for (int j = 0; j < nchunk; ++j) {                                     // (a chunk below the fusing threshold stages its records)
        const size_t len = bound[j + 1] - bound[j];
        if (len == 0) continue;
        m = len > m ? len : m;
        const size_t w = (size_t)ksl_screen_elems_ws((int64_t)(shared ? 1 : len), (int64_t)len, n_coarse);
        ws_bytes = w > ws_bytes ? w : ws_bytes;
    }
    const int nslot = nchunk > 1 ? 2 : 1;
    const size_t mt = shared ? 1 : m;
    auto pad = [](size_t b) { return ((b + 255) / 256) * 256; };
    const size_t slot_bytes = pad(8 * 7 * m) + pad(8 * m) + pad(8 * 7 * mt) + pad(8 * mt) + pad(ws_bytes);
    const size_t need = pad(8 * unc) + 4 * pad(8 * un) + pad(4 * un) + (size_t)nslot * slot_bytes +
                        pad((size_t)ksl_screen_summary_ws(n)) + pad(8 * KSL_SUM_NCOUNT) + pad(8 * KSL_SUM_NMOMENT) + 4096;
    if (hc.cap < need) {
        if (hc.arena) KSL_CUDA(cudaFree(hc.arena));
        hc.arena = nullptr;
        hc.cap = 0;
        KSL_CUDA(cudaMalloc(reinterpret_cast<void **>(&hc.arena), need));
        hc.cap = need;
    }
    Bump b{hc.arena};
    double *d_times = b.take<double>(unc);
    int64_t *d_imin = b.take<int64_t>(un), *d_iconj = b.take<int64_t>(un);
    double *d_dmin = b.take<double>(un), *d_dconj = b.take<double>(un);
    int32_t *d_err = b.take<int32_t>(un);
    struct Slot {
        double *el_c, *ep_c, *el_t, *ep_t;
        void *ws;
    } slot[2];
    for (int q = 0; q < nslot; ++q) {
        slot[q].el_c = b.take<double>(7 * m); slot[q].ep_c = b.take<double>(m);
        slot[q].el_t = b.take<double>(7 * mt); slot[q].ep_t = b.take<double>(mt);
        slot[q].ws = b.take<char>(ws_bytes);
    }
    void *d_sws = b.take<char>((size_t)ksl_screen_summary_ws(n));
    uint64_t *d_counts = b.take<uint64_t>(KSL_SUM_NCOUNT);
    double *d_mom = b.take<double>(KSL_SUM_NMOMENT);
    cudaStream_t streams[2] = {hc.stream, hc.stream2};
    KSL_CUDA(cudaMemcpyAsync(d_times, times_coarse, 8 * unc, cudaMemcpyHostToDevice, streams[0]));
    KSL_CUDA(cudaEventRecord(hc.ev_ready, streams[0]));
    // The results of chunk j go back AFTER chunk j+1 has been queued: with pageable host buffers a device->host copy blocks
    // the calling thread until the chunk's kernel has finished, and issued straight after its kernel it would keep the next
    // chunk's host->device copy from overlapping anything (measured: 189 ms per 1e6 traces that way; with pinned buffers
    // the order makes no difference).
    auto copy_back = [&](int j) -> int {
        if (j < 0) return KSL_OK;
        const size_t off = bound[j], len = bound[j + 1] - bound[j];
        if (len == 0) return KSL_OK;
        cudaStream_t st = streams[j % nslot];
        KSL_CUDA(cudaMemcpyAsync(i_min + off, d_imin + off, 8 * len, cudaMemcpyDeviceToHost, st));
        KSL_CUDA(cudaMemcpyAsync(d_min + off, d_dmin + off, 8 * len, cudaMemcpyDeviceToHost, st));
        KSL_CUDA(cudaMemcpyAsync(i_conj + off, d_iconj + off, 8 * len, cudaMemcpyDeviceToHost, st));
        KSL_CUDA(cudaMemcpyAsync(d_conj + off, d_dconj + off, 8 * len, cudaMemcpyDeviceToHost, st));
        if (err) KSL_CUDA(cudaMemcpyAsync(err + off, d_err + off, 4 * len, cudaMemcpyDeviceToHost, st));
        return KSL_OK;
    };
    int j_last = -1;
    for (int j = 0; j < nchunk; ++j) {
        const size_t off = bound[j], len = bound[j + 1] - bound[j];
        if (len == 0) break;
        Slot &sl = slot[j % nslot];
        cudaStream_t st = streams[j % nslot];
        if (j == 1) KSL_CUDA(cudaStreamWaitEvent(st, hc.ev_ready, 0));       // the time grid, copied on the other stream
        // element SoA [7][n] -> the chunk's own SoA [7][len]
        KSL_CUDA(cudaMemcpy2DAsync(sl.el_c, 8 * len, elems_c + off, 8 * un, 8 * len, 7, cudaMemcpyHostToDevice, st));
        KSL_CUDA(cudaMemcpyAsync(sl.ep_c, epoch_c + off, 8 * len, cudaMemcpyHostToDevice, st));
        const size_t lt = shared ? 1 : len;
        if (shared) {
            KSL_CUDA(cudaMemcpyAsync(sl.el_t, elems_t, 8 * 7, cudaMemcpyHostToDevice, st));
            KSL_CUDA(cudaMemcpyAsync(sl.ep_t, epoch_t, 8, cudaMemcpyHostToDevice, st));
        } else {
            KSL_CUDA(cudaMemcpy2DAsync(sl.el_t, 8 * len, elems_t + off, 8 * un, 8 * len, 7, cudaMemcpyHostToDevice, st));
            KSL_CUDA(cudaMemcpyAsync(sl.ep_t, epoch_t + off, 8 * len, cudaMemcpyHostToDevice, st));
        }
        const int rc = ksl_screen_elems(sl.el_t, sl.ep_t, (int64_t)lt, sl.el_c, sl.ep_c, (int64_t)len, d_times, n_coarse, n_fine, thr_m,
                                        mode, d_imin + off, d_dmin + off, d_iconj + off, d_dconj + off, d_err + off, sl.ws, st);
        if (rc) return rc;
        const int rb = copy_back(j - 1);
        if (rb) return rb;
        j_last = j;
    }
    {
        const int rb = copy_back(j_last);
        if (rb) return rb;
    }
    if (nslot > 1) {
        KSL_CUDA(cudaEventRecord(hc.ev_done, streams[1]));
        KSL_CUDA(cudaStreamWaitEvent(streams[0], hc.ev_done, 0));
    }
    if (summary_counts || summary_moments) {
        const int rc = ksl_screen_summary(d_iconj, d_dmin, d_err, n, collision_thr_m, d_counts, d_mom, d_sws, streams[0]);
        if (rc) return rc;
        if (summary_counts) KSL_CUDA(cudaMemcpyAsync(summary_counts, d_counts, 8 * KSL_SUM_NCOUNT, cudaMemcpyDeviceToHost, streams[0]));
        if (summary_moments) KSL_CUDA(cudaMemcpyAsync(summary_moments, d_mom, 8 * KSL_SUM_NMOMENT, cudaMemcpyDeviceToHost, streams[0]));
    }
    return KSL_OK;
}
}  // namespace

extern "C" int ksl_host_cache_free(void) {
    int cur = 0;
    cudaGetDevice(&cur);
    for (int d = 0; d < 64; ++d) {
        std::lock_guard<std::mutex> lk(g_cache[d].mu);
        HostCache &hc = g_cache[d];
        if (!hc.arena && !hc.stream) continue;
        cudaSetDevice(d);
        if (hc.arena) cudaFree(hc.arena);
        if (hc.stream) cudaStreamDestroy(hc.stream);
        if (hc.stream2) cudaStreamDestroy(hc.stream2);
        if (hc.ev_ready) cudaEventDestroy(hc.ev_ready);
        if (hc.ev_done) cudaEventDestroy(hc.ev_done);
        hc.arena = nullptr; hc.cap = 0;
        hc.stream = hc.stream2 = nullptr;
        hc.ev_ready = hc.ev_done = nullptr;
    }
    cudaSetDevice(cur);
    return KSL_OK;
}

extern "C" int ksl_screen_host(int device, const double *elems_t, const double *epoch_t, int64_t n_t, const double *elems_c,
                               const double *epoch_c, int64_t n, const double *times_coarse, int64_t n_coarse,
                               int64_t n_fine, double thr_m, int mode, double collision_thr_m, int64_t *i_min,
                               double *d_min, int64_t *i_conj, double *d_conj, int32_t *err, uint64_t *summary_counts,
                               double *summary_moments) {
    if (device < 0 || device >= 64) return fail(KSL_E_ARG, "ksl_screen_host: bad device %d", device);
    if (n < 0 || n_coarse <= 0 || n_fine <= 0 || (n_t != 1 && n_t != n)) return fail(KSL_E_ARG, "ksl_screen_host: bad sizes");
    if (n == 0) return KSL_OK;
    if (!elems_t || !epoch_t || !elems_c || !epoch_c || !times_coarse || !i_min || !d_min || !i_conj || !d_conj)
        return fail(KSL_E_ARG, "ksl_screen_host: NULL pointer");
    if (ksl_device_count() <= device) return fail(KSL_E_CUDA, "ksl_screen_host: CUDA device %d not available (no CPU fallback)", device);
    HostCache &hc = g_cache[device];
    std::lock_guard<std::mutex> lk(hc.mu);   // per device: a process driving 8 GPUs from 8 threads is not serialised
    int prev = 0;
    KSL_CUDA(cudaGetDevice(&prev));          // the current device is per host thread
    KSL_CUDA(cudaSetDevice(device));
    int rc = screen_host_locked(hc, elems_t, epoch_t, n_t, elems_c, epoch_c, n, times_coarse, n_coarse, n_fine, thr_m, mode,
                                collision_thr_m, i_min, d_min, i_conj, d_conj, err, summary_counts, summary_moments);
    // success or not, nothing of this call may still be in flight when the arena becomes reusable and the caller's
    // buffers are handed back
    const cudaError_t e0 = hc.stream ? cudaStreamSynchronize(hc.stream) : cudaSuccess;
    const cudaError_t e1 = hc.stream2 ? cudaStreamSynchronize(hc.stream2) : cudaSuccess;
    if (rc == KSL_OK && (e0 != cudaSuccess || e1 != cudaSuccess))
        rc = fail(KSL_E_CUDA, "ksl_screen_host: %s", cudaGetErrorString(e0 != cudaSuccess ? e0 : e1));
    cudaSetDevice(prev);
    return rc;
}
