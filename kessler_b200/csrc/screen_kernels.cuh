// screen_kernels.cuh -- the fused-trace kernels of libkessler_b200.so (K3: k_screen, k_screen_warp), model.find_conjunction
// on resident trajectories, and the helpers they share.  Textually included by kessler_b200.cu INSIDE its anonymous
// namespace (one translation unit, internal linkage); split out so that the evidence under profiles/ can be tied to
// the sources of the kernel it describes (bench.kernel_sources_sha hashes this file and the device headers it uses).
// Needs: sgp4_device.cuh, screen_device.cuh, Best / best_* (using-declarations of kessler_b200.cu).
#pragma once

// sgp4init of one element set into a private fast record (the 36 public constants, then the derived products).
// Out of line on purpose: (1) one body of machine code serves k_sgp4_init and the fused kernel's prologue, so the
// two paths produce bit-identical records; (2) its registers do not count against the epoch loop.
__device__ __noinline__ int sgp4_init_record(const double *el, double *rec) {
    double e7[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) e7[k] = el[k];
    return ksl::sgp4_init(e7, [&](int k, double v) { rec[k] = v; });
}

// ------------------------------------------------------------------------------------
// arg-min / first-below bookkeeping shared by find_conjunction and the fused screen.
// torch semantics (model.py:23-52): first index wins ties, NaN is the minimum.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ Best best_shfl_down(const Best &b, int delta) {
    Best o;
    o.sq = __shfl_down_sync(0xffffffffu, b.sq, delta);
    o.j = __shfl_down_sync(0xffffffffu, b.j, delta);
    o.nan_j = __shfl_down_sync(0xffffffffu, b.nan_j, delta);
    o.conj_j = __shfl_down_sync(0xffffffffu, b.conj_j, delta);
    o.conj_sq = __shfl_down_sync(0xffffffffu, b.conj_sq, delta);
    return o;
}

// all threads of the CTA call; result valid in thread 0.  `scratch` holds >= 32 Best.
__device__ __forceinline__ Best best_block_reduce(Best b, Best *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        Best o = best_shfl_down(b, d);
        best_merge(b, o);
    }
    if (lane == 0) scratch[warp] = b;
    __syncthreads();
    if (warp == 0) {
        if (lane < nwarp) b = scratch[lane];
        else best_init(b);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            Best o = best_shfl_down(b, d);
            best_merge(b, o);
        }
    }
    __syncthreads();
    return b;
}

// ------------------------------------------------------------------------------------
// model.find_conjunction on two resident trajectories (two-stage reduction)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_findconj_partial(const double *__restrict__ tr0,
                                                          const double *__restrict__ tr1, long long n,
                                                          long long stride, double thr2, Best *__restrict__ partial) {
    __shared__ Best scratch[32];
    Best b;
    best_init(b);
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
        const double dx = ksl::add_rn(tr0[j * stride], -tr1[j * stride]);
        const double dy = ksl::add_rn(tr0[j * stride + 1], -tr1[j * stride + 1]);
        const double dz = ksl::add_rn(tr0[j * stride + 2], -tr1[j * stride + 2]);
        best_visit(b, ksl::sqdist3(dx, dy, dz), j, thr2);
    }
    b = best_block_reduce(b, scratch);
    if (threadIdx.x == 0) partial[blockIdx.x] = b;
}

__global__ void __launch_bounds__(256) k_findconj_final(const Best *__restrict__ partial, int nparts,
                                                        long long *__restrict__ out_i, double *__restrict__ out_d) {
    __shared__ Best scratch[32];
    Best b;
    best_init(b);
    for (int p = threadIdx.x; p < nparts; p += blockDim.x) best_merge(b, partial[p]);
    b = best_block_reduce(b, scratch);
    if (threadIdx.x == 0) ksl::best_result(b, out_i, out_d, out_i + 1, out_d + 1);
}

// ------------------------------------------------------------------------------------
// K3  the fused trace.  One CTA per sample (persistent grid, samples strided over
// CTAs); threads run along the coarse epochs: thread -> (sample, epoch), as the
// north star asks.  Per chunk of kScreenThreads epochs the CTA
//   1. evaluates SGP4 for chaser (and target unless it is shared) into shared memory,
//   2. lets thread i own coarse segment [k, k+1]: the ~100 fine-grid points whose
//      torch source index floors to k.
// KSL_SCREEN_BRUTE visits every fine point with the reference's arithmetic.
// KSL_SCREEN_ANALYTIC uses that the squared distance is a quadratic in the
// interpolation weight on a segment: it locates the vertex and the threshold
// crossing in closed form and re-evaluates only the neighbouring fine points with
// the reference's arithmetic, falling back to the brute loop for a segment that is
// non-finite or ill-conditioned.  Segments whose analytic minimum cannot beat the
// thread's running best (nor the threshold) are skipped.
// ------------------------------------------------------------------------------------
// Shape of the fused-trace kernels, tuned on B200 (profiles/README.md, round 1): one 512-thread CTA per SM
// (16 warps, 128 registers/thread) beat 2 x 256, 3 x 256 (80 regs, spills) and 5-7 x 128.
#ifndef KSL_SCREEN_THREADS
#define KSL_SCREEN_THREADS 512
#endif
#ifndef KSL_SCREEN_MIN_CTAS
#define KSL_SCREEN_MIN_CTAS 1
#endif
#ifndef KSL_SCREEN_GRID_SMEM
#define KSL_SCREEN_GRID_SMEM 1   // stage the time grid in shared memory (north star); 0 = read it through L1
#endif
constexpr int kScreenThreads = KSL_SCREEN_THREADS;
constexpr int kMaxGridInSmem = 6144;  // time grid staged in shared memory up to this many epochs

struct ScreenParams {
    const double *consts_t, *epoch_t;   // consts_*: initialised records SoA [KSL_NCONST][n_*] (k_screen, k_screen_warp<.., false>)
    const int *init_err_t;
    long long n_t;
    const double *consts_c, *epoch_c;
    const int *init_err_c;
    long long n;
    const double *elems_t, *elems_c;    // SGP4 inputs SoA [7][n_*]: k_screen_warp<.., true> runs sgp4init itself
    unsigned long long *counter;        // next group of traces to hand out (zeroed before the launch); NULL = static striding
    int group;                          // traces per group, 1..kMaxGroup
    const double *times_coarse;
    long long n_coarse, n_fine;
    double scale, thr2;
    const double *target_pos;  // [n_coarse][3] km when the target is shared, else NULL
    const int *target_err;     // [1] when shared
    long long *i_min, *i_conj;
    double *d_min, *d_conj;
    int *err;
};

template <bool kSharedTarget, int kMode>
__global__ void __launch_bounds__(kScreenThreads, KSL_SCREEN_MIN_CTAS) k_screen(const ScreenParams p) {
    extern __shared__ double s_times[];  // [min(n_coarse, kMaxGridInSmem)]
    __shared__ double s_ct[ksl::KSL_NFAST], s_cc[ksl::KSL_NFAST];
    __shared__ double s_pt[kScreenThreads][3], s_pc[kScreenThreads][3];
    __shared__ Best s_scratch[32];
    __shared__ double s_epoch[2];
    __shared__ int s_flags[2];
    __shared__ __align__(16) double s_grid[2 * ksl::kGridPoints];   // ksl::kGridSinCos, for sincos_grid

    const int tid = threadIdx.x;
    for (int q = tid; q < 2 * ksl::kGridPoints; q += kScreenThreads) s_grid[q] = ksl::kGridSinCos[q];
    const bool grid_in_smem = KSL_SCREEN_GRID_SMEM && p.n_coarse <= kMaxGridInSmem;
    int *s_ja = reinterpret_cast<int *>(s_times + (grid_in_smem ? p.n_coarse : 0));
    if (grid_in_smem)
        for (long long k = tid; k < p.n_coarse; k += kScreenThreads) s_times[k] = p.times_coarse[k];
    const bool ja_in_smem = p.n_coarse <= kMaxGridInSmem && p.n_fine < 0x7fffffffLL;
    if (ja_in_smem)
        for (long long k = tid; k <= p.n_coarse; k += kScreenThreads)
            s_ja[k] = (k == p.n_coarse) ? (int)p.n_fine : (int)ksl::first_fine_of(k, p.scale, p.n_coarse, p.n_fine);
    const int *ja_tab = ja_in_smem ? s_ja : nullptr;

    for (long long s = blockIdx.x; s < p.n; s += gridDim.x) {
        __syncthreads();  // previous sample's readers are done with s_ct/s_cc/s_pt/s_pc
        const long long it = (p.n_t == 1) ? 0 : s;
        if (tid < KSL_NCONST) s_cc[tid] = p.consts_c[(long long)tid * p.n + s];
        if (!kSharedTarget && tid >= 64 && tid < 64 + KSL_NCONST) s_ct[tid - 64] = p.consts_t[(long long)(tid - 64) * p.n_t + it];
        if (tid == kScreenThreads - 1) {
            s_epoch[0] = p.epoch_t[it];
            s_epoch[1] = p.epoch_c[s];
            s_flags[0] = p.init_err_t ? p.init_err_t[it] : 0;
            s_flags[1] = p.init_err_c ? p.init_err_c[s] : 0;
        }
        __syncthreads();
        if (tid == 0) ksl::prepare_fast_record(s_cc, 1);
        if (!kSharedTarget && tid == 32) ksl::prepare_fast_record(s_ct, 1);
        __syncthreads();
        const int ierr_t = s_flags[0], ierr_c = s_flags[1];
        if ((ierr_t | ierr_c) & KSL_ERR_DEEPSPACE) {  // dsgp4.initialize_tle raises -> prop error (model.py:584,595)
            if (tid == 0) {
                p.i_min[s] = -1;
                p.d_min[s] = NAN;
                p.i_conj[s] = -1;
                p.d_conj[s] = NAN;
                p.err[s] = ierr_t | (ierr_c << KSL_ERR_CHASER_SHIFT);
            }
            continue;
        }
        const double epoch_t = s_epoch[0], epoch_c = s_epoch[1];
        const long long last = p.n_coarse - 1;
        // the same choice of SGP4 form as k_screen_warp makes (near-circular records, decided per trace), so that the two
        // work decompositions return bit-identical results
        bool small_c, small_t = false;
        {
            const double t_first = p.times_coarse[0], t_last = p.times_coarse[last];
            small_c = ksl::fast_record_is_small_e([&](int q) { return s_cc[q]; }, 1440.0 * fmax(fabs(t_first - epoch_c), fabs(t_last - epoch_c)));
            if (!kSharedTarget) {
                small_t = ksl::fast_record_is_small_e([&](int q) { return s_ct[q]; }, 1440.0 * fmax(fabs(t_first - epoch_t), fabs(t_last - epoch_t)));
                small_c = small_t = (small_c && small_t);
            }
        }

        Best best;
        best_init(best);
        int e_t = 0, e_c = 0;
        for (long long base = 0; base < last || base == 0; base += kScreenThreads - 1) {
            const long long k = base + tid;
            if (k <= last) {
                const double tm = grid_in_smem ? s_times[k] : p.times_coarse[k];
                double r[3];
                e_c |= ksl::sgp4_position(s_cc, 1, (tm - epoch_c) * 1440.0, r, s_grid, small_c);
                s_pc[tid][0] = r[0]; s_pc[tid][1] = r[1]; s_pc[tid][2] = r[2];
                if (kSharedTarget) {
                    r[0] = p.target_pos[3 * k]; r[1] = p.target_pos[3 * k + 1]; r[2] = p.target_pos[3 * k + 2];
                } else {
                    e_t |= ksl::sgp4_position(s_ct, 1, (tm - epoch_t) * 1440.0, r, s_grid, small_t);
                }
                s_pt[tid][0] = r[0]; s_pt[tid][1] = r[1]; s_pt[tid][2] = r[2];
            }
            __syncthreads();
            // thread tid owns segment k when its right neighbour is in this chunk, or when k is the very
            // last coarse epoch (degenerate segment: torch clamps i1 to n_in-1)
            const bool owner = (k < last && tid < kScreenThreads - 1) || (k == last);
            if (owner) {
                const int nb = (k == last) ? tid : tid + 1;
                const double *t0 = s_pt[tid], *t1 = s_pt[nb], *c0 = s_pc[tid], *c1 = s_pc[nb];
                const double d0[3] = {t0[0] - c0[0], t0[1] - c0[1], t0[2] - c0[2]};
                const double d1[3] = {t1[0] - c1[0], t1[1] - c1[1], t1[2] - c1[2]};
                if (kMode == KSL_SCREEN_BRUTE || !ksl::segment_skippable(d0, d1, best.sq, best.conj_j == ksl::kNoIndex, p.thr2))
                    ksl::segment_search<kMode>(k, last, p.scale, p.n_coarse, p.n_fine, p.thr2, t0, t1, c0, c1, best, ja_tab);
            }
            __syncthreads();
        }
        // OR the error bits over the CTA
        e_t = __reduce_or_sync(0xffffffffu, e_t);
        e_c = __reduce_or_sync(0xffffffffu, e_c);
        __shared__ int s_err[2];
        if (tid == 0) { s_err[0] = 0; s_err[1] = 0; }
        __syncthreads();
        if ((tid & 31) == 0) {
            if (e_t) atomicOr(&s_err[0], e_t);
            if (e_c) atomicOr(&s_err[1], e_c);
        }
        best = best_block_reduce(best, s_scratch);  // contains __syncthreads
        if (tid == 0) {
            ksl::best_result(best, p.i_min + s, p.d_min + s, p.i_conj + s, p.d_conj + s);
            int et = s_err[0] | ierr_t;
            if (kSharedTarget && p.target_err) et |= *p.target_err;
            p.err[s] = et | ((s_err[1] | ierr_c) << KSL_ERR_CHASER_SHIFT);
        }
    }
}

// ------------------------------------------------------------------------------------
// K3w  the fused trace, one WARP per sample.  Same work decomposition as k_screen (lane <->
// coarse epoch, lane i owns segment [k, k+1]) but nothing is shared between warps after the
// prologue: the neighbour's positions come from warp shuffles, the per-trace constants sit in a
// per-warp slice of shared memory, and the only block-wide barrier is the one after the time
// grid is staged.  Warps therefore drift apart and the FP64 pipe sees a mix of the FP64-dense
// and the integer/selection phases of different warps instead of CTA-wide lock-step phases
// (ncu, round 1: FP64 pipe 57 % active with `wait` + `barrier` on top).  Used when there are
// enough samples to fill the machine with warps; k_screen (CTA per trace) serves small batches.
// ------------------------------------------------------------------------------------
constexpr int kWarpsPerCta = kScreenThreads / 32;

__device__ __forceinline__ double shfl_up_d(double v, int delta) { return __shfl_up_sync(0xffffffffu, v, delta); }

// Best with 32-bit indices: what k_screen_warp keeps live across its epoch loop (the launcher only picks that kernel
// when n_fine and n_coarse fit; register pressure, not arithmetic, is what limits that loop's scheduling)
struct BestW {
    double sq, conj_sq;
    int j, nan_j, conj_j;
};
constexpr int kNoIndexW = 0x7fffffff;
__device__ __forceinline__ Best best_expand(const BestW &w) {
    Best b;
    b.sq = w.sq;
    b.conj_sq = w.conj_sq;
    b.j = (w.j == kNoIndexW) ? ksl::kNoIndex : (long long)w.j;
    b.nan_j = (w.nan_j == kNoIndexW) ? ksl::kNoIndex : (long long)w.nan_j;
    b.conj_j = (w.conj_j == kNoIndexW) ? ksl::kNoIndex : (long long)w.conj_j;
    return b;
}
__device__ __forceinline__ BestW best_compact(const Best &b) {
    BestW w;
    w.sq = b.sq;
    w.conj_sq = b.conj_sq;
    w.j = (b.j == ksl::kNoIndex) ? kNoIndexW : (int)b.j;
    w.nan_j = (b.nan_j == ksl::kNoIndex) ? kNoIndexW : (int)b.nan_j;
    w.conj_j = (b.conj_j == ksl::kNoIndex) ? kNoIndexW : (int)b.conj_j;
    return w;
}

// Traces are handed to warps in GROUPS: the warp first fills its shared-memory slice with the group's fast records --
// one lane per record, from the raw elements (kFromElems: sgp4init runs here, nothing but 2 x 7 doubles per trace is
// read from HBM) or from records k_sgp4_init wrote --, then runs the traces of the group one after the other.  Groups
// come from an atomic counter when the caller provides one (tail balance: a trace on the exact SGP4 path costs 4-5
// fast ones), else by static striding.
constexpr int kMaxGroup = 8;                    // traces per group with per-sample targets (2 x 8 records per warp)
constexpr int kGroupRecs = 2 * kMaxGroup;       // records per warp slice; a shared target leaves all 16 to chasers
constexpr int kRecEpoch = ksl::KSL_F_PAD;       // spare slots of a fast record: the record's epoch (MJD) ...
constexpr int kRecErr = KSL_C_PAD;              // ... and its sgp4init error flag (as a double)

// sgp4init flag of a record in the warp's slice (a shared target has none there: its flag is ScreenParams::target_err)
template <bool kAbsent>
__device__ __forceinline__ int record_err(const double *rec) { return kAbsent ? 0 : (int)rec[kRecErr]; }

// The group's records into the warp's slice: chaser of trace q in slot q, target in slot kMaxGroup + q.  Called by the whole
// warp; out of line so that the prologue's registers (sgp4init's arguments, the copy loops) stay out of the epoch loop's budget.
template <bool kSharedTarget, bool kFromElems>
__device__ __noinline__ void fill_group_records(const double *src_t, const double *epoch_t, const int *ierr_t, long long n_t,
                                                const double *src_c, const double *epoch_c, const int *ierr_c, long long n,
                                                long long s0, int cnt, double *s_rec) {
    const int lane = threadIdx.x & 31;
    const int nrec = kSharedTarget ? cnt : 2 * cnt;
    if (!kFromElems) {
        for (int r = 0; r < nrec; ++r) {
            const bool tgt = r >= cnt;
            const long long idx = s0 + (tgt ? r - cnt : r);
            const double *src = tgt ? src_t : src_c;
            const long long stride = tgt ? n_t : n;
            double *rec = s_rec + (tgt ? kMaxGroup + (r - cnt) : r) * ksl::KSL_NFAST;
            for (int q = lane; q < KSL_NCONST; q += 32) rec[q] = src[(long long)q * stride + idx];
        }
        __syncwarp();
    }
    if (lane < nrec) {
        const bool tgt = lane >= cnt;
        const int q = tgt ? lane - cnt : lane;
        const long long idx = s0 + q;
        double *rec = s_rec + (tgt ? kMaxGroup + q : q) * ksl::KSL_NFAST;
        int ierr;
        if (kFromElems) {
            const double *src = tgt ? src_t : src_c;
            const long long stride = tgt ? n_t : n;
            double el[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) el[k] = src[(long long)k * stride + idx];
            ierr = sgp4_init_record(el, rec);
        } else {
            const int *ie = tgt ? ierr_t : ierr_c;
            ierr = ie ? ie[idx] : 0;
        }
        ksl::prepare_fast_record(rec, 1);
        rec[kRecErr] = (double)ierr;
        rec[kRecEpoch] = (tgt ? epoch_t : epoch_c)[idx];
    }
    __syncwarp();
}

template <bool kSharedTarget, int kMode, bool kFromElems>
__global__ void __launch_bounds__(kScreenThreads, KSL_SCREEN_MIN_CTAS) k_screen_warp(const ScreenParams p) {
    extern __shared__ double s_dyn[];
    __shared__ __align__(16) double s_grid[2 * ksl::kGridPoints];          // ksl::kGridSinCos, for sincos_grid
    __shared__ double s_carry_all[kWarpsPerCta][8];   // lane 31's (target, chaser, float d) for the next chunk's lane 0

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool grid_in_smem = KSL_SCREEN_GRID_SMEM && p.n_coarse <= kMaxGridInSmem;
    // dynamic shared memory: [records: warps x kGroupRecs x KSL_NFAST][time grid][first fine index per segment (+ n_fine)]
    double *s_times = s_dyn + kWarpsPerCta * kGroupRecs * ksl::KSL_NFAST;
    int *s_ja = reinterpret_cast<int *>(s_times + (grid_in_smem ? p.n_coarse : 0));
    if (grid_in_smem)
        for (long long k = tid; k < p.n_coarse; k += kScreenThreads) s_times[k] = p.times_coarse[k];
    const bool ja_in_smem = p.n_coarse <= kMaxGridInSmem && p.n_fine < 0x7fffffffLL;
    if (ja_in_smem)
        for (long long k = tid; k <= p.n_coarse; k += kScreenThreads)
            s_ja[k] = (k == p.n_coarse) ? (int)p.n_fine : (int)ksl::first_fine_of(k, p.scale, p.n_coarse, p.n_fine);
    const int *ja_tab = ja_in_smem ? s_ja : nullptr;
    for (int q = tid; q < 2 * ksl::kGridPoints; q += kScreenThreads) s_grid[q] = ksl::kGridSinCos[q];
    __syncthreads();

    double *s_rec = s_dyn + warp * kGroupRecs * ksl::KSL_NFAST;
    double *s_carry = s_carry_all[warp];
    float *s_carry_f = reinterpret_cast<float *>(s_carry + 6);
    const int last = (int)p.n_coarse - 1;
    // first epoch of the pilot chunk (see below): >= 0 use it, kNoHint none, otherwise -1 - epoch = known but the last two
    // traces of this warp were unrelated
    constexpr int kNoHint = -0x7fffffff - 1;
    int hint = kNoHint;
    unsigned int g_static = blockIdx.x * kWarpsPerCta + warp;   // (the launcher keeps the number of groups below 2^32)
    for (;;) {
        unsigned int g;
        if (p.counter) {
            unsigned int t = 0;
            if (lane == 0) t = (unsigned int)atomicAdd(p.counter, 1ULL);
            g = __shfl_sync(0xffffffffu, t, 0);
        } else {
            g = g_static;
            g_static += gridDim.x * kWarpsPerCta;
        }
        const long long s0 = (long long)g * p.group;
        if (s0 >= p.n) break;
        const int cnt = (p.n - s0 < p.group) ? (int)(p.n - s0) : p.group;
        __syncwarp();
        fill_group_records<kSharedTarget, kFromElems>(kFromElems ? p.elems_t : p.consts_t, p.epoch_t, p.init_err_t, p.n_t,
                                                      kFromElems ? p.elems_c : p.consts_c, p.epoch_c, p.init_err_c, p.n,
                                                      (long long)g * p.group, cnt, s_rec);
        for (int tq = 0; tq < cnt; ++tq) {
        const double *s_cc = s_rec + tq * ksl::KSL_NFAST, *s_ct = s_rec + (kMaxGroup + tq) * ksl::KSL_NFAST;
        const double epoch_c = s_cc[kRecEpoch], epoch_t = kSharedTarget ? 0.0 : s_ct[kRecEpoch];
        if ((record_err<kSharedTarget>(s_ct) | record_err<false>(s_cc)) & KSL_ERR_DEEPSPACE) {
            if (lane == 0) {
                const long long s = (long long)g * p.group + tq;
                const int ierr_c = record_err<false>(s_cc), ierr_t = record_err<kSharedTarget>(s_ct);
                p.i_min[s] = -1;
                p.d_min[s] = NAN;
                p.i_conj[s] = -1;
                p.d_conj[s] = NAN;
                p.err[s] = ierr_t | (ierr_c << KSL_ERR_CHASER_SHIFT);
            }
            continue;
        }
        // near-circular records take the shorter Kepler solve (sgp4_device.cuh): decided once per trace from a bound on
        // the eccentricity over the window, uniform over the warp
        bool small_c, small_t = false;
        {
            const double t_first = grid_in_smem ? s_times[0] : __ldg(p.times_coarse), t_last = grid_in_smem ? s_times[last] : __ldg(p.times_coarse + last);
            small_c = ksl::fast_record_is_small_e([&](int q) { return s_cc[q]; }, 1440.0 * fmax(fabs(t_first - epoch_c), fabs(t_last - epoch_c)));
            if (!kSharedTarget) {
                small_t = ksl::fast_record_is_small_e([&](int q) { return s_ct[q]; }, 1440.0 * fmax(fabs(t_first - epoch_t), fabs(t_last - epoch_t)));
                small_c = small_t = (small_c && small_t);     // a mixed pair takes the general form for both (two code paths, not four)
            }
        }
        BestW best = {INFINITY, 0.0, kNoIndexW, kNoIndexW, kNoIndexW};
        int e_t = 0, e_c = 0;
        double cutoff = INFINITY;   // warp-wide running minimum (m^2), refreshed whenever a lane visited a segment
        float want_f = INFINITY;    // what a segment's lower bound has to exceed to be skipped (km^2), follows cutoff
        // Lazy search (one object against a catalog: consecutive traces are unrelated, the pilot chunk is off, and a
        // trace sets a new running minimum at almost every orbit -- a third of the chunks would enter the exact-candidate
        // stage).  A segment that may hold the minimum but cannot hold a threshold crossing is not searched at once: its
        // lane keeps its index and lower bound, and the warp keeps `cut_f`, an UPPER bound of the trace's final minimum
        // (every such segment holds a fine point below its ub).  Segments whose lower bound exceeds cut_f are skipped
        // like those that exceed the running minimum; a lane that meets a second candidate drops the first when it is
        // dead (lb above cut_f) and otherwise searches the new one on the spot.  After the last chunk the surviving
        // candidates -- typically one or two per trace -- are searched: their two end points are evaluated again (same
        // function of (record, epoch): same positions).  The result is exactly the eager one.
        constexpr bool kLazy = kSharedTarget && kMode == KSL_SCREEN_ANALYTIC;
        const bool lazy = kLazy && ja_tab != nullptr && p.scale <= 0.125 && (double)p.n_coarse * 1.2e-7 <= 0.25 * p.scale;
        const float dl_f = (float)(2.0 * p.scale), thr2_f = (float)(p.thr2 * 1e-6), thr2_want_f = ksl::segment_want_f32(0.0, true, p.thr2);
        int kd = -1;                // the lane's deferred segment ends at epoch kd (-1: none)
        float lbd = 0.0f;           // its lower bound (km^2)
        float cut_f = INFINITY;     // warp-wide upper bound of the final minimum from deferred segments (km^2)
        // lane i of a chunk evaluates epoch k = base + i and owns the segment [k-1, k] that ENDS there; lane 0
        // finds the left end, lane 31 of the previous chunk, in the warp's carry slot (no epoch evaluated twice).
        // -- SGP4 of one epoch per lane: both objects (the shared target's position is looked up) ----------------------
        auto evaluate = [&](int k, double rt[3], double rc[3]) {
            if (k > last) return;
            const double tm = grid_in_smem ? s_times[k] : __ldg(p.times_coarse + k);
            // both fast evaluations back to back: one straight-line block with two independent
            // dependency chains for the scheduler; the (rare) exact fallbacks come after
            const double ts_c = (tm - epoch_c) * 1440.0, ts_t = (tm - epoch_t) * 1440.0;
            bool ok_c, ok_t = true;
            auto ld_c = [&](int q) { return s_cc[q]; };
            auto ld_t = [&](int q) { return s_ct[q]; };
            if (kSharedTarget) {
                rt[0] = __ldg(p.target_pos + 3 * k); rt[1] = __ldg(p.target_pos + 3 * k + 1); rt[2] = __ldg(p.target_pos + 3 * k + 2);
                ok_c = small_c ? ksl::sgp4_prop_fast<true>(ld_c, ts_c, rc, s_grid) : ksl::sgp4_prop_fast<false>(ld_c, ts_c, rc, s_grid);
            } else if (small_c && small_t) {     // one uniform branch per chunk; each arm is one straight-line block
                ok_c = ksl::sgp4_prop_fast<true>(ld_c, ts_c, rc, s_grid);
                ok_t = ksl::sgp4_prop_fast<true>(ld_t, ts_t, rt, s_grid);
            } else {
                ok_c = ksl::sgp4_prop_fast<false>(ld_c, ts_c, rc, s_grid);
                ok_t = ksl::sgp4_prop_fast<false>(ld_t, ts_t, rt, s_grid);
            }
            // (through a scratch array: the out-of-line call takes an address, and rc / rt must stay in registers)
            if (!ok_c) {
                double rx[3];
                e_c |= ksl::sgp4_position_exact(s_cc, 1, ts_c, rx);
                rc[0] = rx[0]; rc[1] = rx[1]; rc[2] = rx[2];
            }
            if (!ok_t) {
                double rx[3];
                e_t |= ksl::sgp4_position_exact(s_ct, 1, ts_t, rx);
                rt[0] = rx[0]; rt[1] = rx[1]; rt[2] = rx[2];
            }
        };
        // -- the segments that end at the chunk's epochs -----------------------------------------------------------------
        auto search = [&](int k, bool pilot, const double rt[3], const double rc[3]) {
            // stage 1, in FP32 (conservative, see segment_skippable_f32): relative position at this epoch (d1) and
            // at the left-hand neighbour (d0: 3 shuffles)
            const bool owner = k >= 1 && k <= last && !(pilot && lane == 0);   // the pilot's lane 0 has no left neighbour
            float d0[3], d1[3];
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                d1[x] = (float)(rt[x] - rc[x]);
                d0[x] = __shfl_up_sync(0xffffffffu, d1[x], 1);
                if (lane == 0) d0[x] = s_carry_f[x];
            }
            // what a segment's lower bound has to exceed to be skipped: the running minimum -- or the upper bound of the
            // final one the deferred segments give --, but never less than the threshold while this lane still looks for
            // its first crossing
            const float eff_f = (kLazy && lazy) ? fmaxf(fminf(want_f, cut_f), best.conj_j == kNoIndexW ? thr2_want_f : 0.0f) : want_f;
            bool need = owner && (kMode == KSL_SCREEN_BRUTE || !ksl::segment_skippable_f32(d0, d1, eff_f));
            if (kLazy && lazy && __any_sync(0xffffffffu, need)) {
                bool tightened = false;
                float lb, ub;
                if (need && ksl::segment_bounds_f32(d0, d1, dl_f, &lb, &ub) &&
                    !(best.conj_j == kNoIndexW && lb < thr2_f * 1.0001f + 1e-6f)) {     // no crossing possible in it (or none needed)
                    const float cut_new = fminf(cut_f, ub);
                    if (kd < 0 || lbd > cut_new) {       // nothing deferred yet, or what is deferred is dead: keep this one instead
                        kd = k;
                        lbd = lb;
                        need = false;
                    }
                    tightened = cut_new < cut_f;
                    cut_f = cut_new;
                }
                if (__any_sync(0xffffffffu, tightened)) {
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) cut_f = fminf(cut_f, __shfl_xor_sync(0xffffffffu, cut_f, d));
                }
            }
            // stage 2 (rare): the neighbour's full positions, then the exact candidate evaluation
            if (__any_sync(0xffffffffu, need)) {
                double t0[3], c0[3];
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    t0[x] = shfl_up_d(rt[x], 1);
                    c0[x] = shfl_up_d(rc[x], 1);
                    if (lane == 0) { t0[x] = s_carry[x]; c0[x] = s_carry[3 + x]; }
                }
                if (need) {
                    Best b = best_expand(best);
                    ksl::segment_search<kMode>(k - 1, last, p.scale, p.n_coarse, p.n_fine, p.thr2, t0, rt, c0, rc, b, ja_tab);
                    best = best_compact(b);
                }
                // the smallest squared distance any lane of this trace has seen prunes every lane's later segments
                cutoff = best.sq;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) cutoff = fmin(cutoff, __shfl_xor_sync(0xffffffffu, cutoff, d));
                want_f = ksl::segment_want_f32(cutoff, best.conj_j == kNoIndexW, p.thr2);
            }
            // The pilot only seeds the running MINIMUM.  A crossing it found must not close this lane's search for the
            // FIRST crossing: "crossing seen -> later segments need not be visited for the threshold" is only valid
            // while segments are visited in increasing order, and the pilot's segments are late ones.  They are
            // visited again in their turn, so forgetting the pilot's crossing loses nothing.
            if (pilot) {
                best.conj_j = kNoIndexW;
                best.conj_sq = 0.0;
                want_f = ksl::segment_want_f32(cutoff, true, p.thr2);
            }
            __syncwarp();
            if (lane == 31) {
#pragma unroll
                for (int x = 0; x < 3; ++x) { s_carry[x] = rt[x]; s_carry[3 + x] = rc[x]; s_carry_f[x] = d1[x]; }
            }
            __syncwarp();
            // torch's clamped tail segment [last, last]: both ends are the last epoch (its owner is the lane that holds it)
            if (k == last && !pilot) {
                const double dl[3] = {rt[0] - rc[0], rt[1] - rc[1], rt[2] - rc[2]};
                if (kMode == KSL_SCREEN_BRUTE || !ksl::segment_skippable(dl, dl, cutoff, best.conj_j == kNoIndexW, p.thr2)) {
                    Best b = best_expand(best);
                    ksl::segment_search<kMode>(last, last, p.scale, p.n_coarse, p.n_fine, p.thr2, rt, rt, rc, rc, b, ja_tab);
                    best = best_compact(b);
                }
            }
        };
        double rt[3] = {0.0, 0.0, 0.0}, rc[3] = {0.0, 0.0, 0.0};
        // Pilot chunk: the 32 epochs around the PREVIOUS trace's closest approach are searched first.  In a
        // Monte Carlo batch consecutive traces are perturbations of one pair, so this seeds the warp-wide running
        // minimum (`cutoff`) close to the final one and the skip test then prunes nearly every other segment.  It only
        // changes the order in which segments are visited (the pilot's epochs are evaluated again in their turn),
        // never the result: arg-min and first crossing are order-independent (ties go to the smaller index).
        // It is used once two consecutive traces of this warp bottomed out within 64 epochs of each other (related
        // traces); for unrelated ones (a catalog) it would be a wasted chunk.
        if (hint >= 0) {
            evaluate(hint + lane, rt, rc);
            search(hint + lane, true, rt, rc);
        }
        for (int base = 0; base <= last; base += 32) {
            evaluate(base + lane, rt, rc);
            search(base + lane, false, rt, rc);
        }
        if (kLazy && lazy) {
            // the deferred segments that can still hold the minimum: lower bound not above what is known by now
            const float known = fminf(cut_f, ksl::segment_want_f32(cutoff, false, p.thr2));
            const bool alive = kd >= 1 && lbd <= known;
            if (__any_sync(0xffffffffu, alive)) {
                double t0[3] = {0.0, 0.0, 0.0}, c0[3] = {0.0, 0.0, 0.0};
                const int ke = alive ? kd : 1;
                evaluate(ke - 1, t0, c0);
                evaluate(ke, rt, rc);
                if (alive) {
                    Best b = best_expand(best);
                    ksl::segment_search<kMode>(ke - 1, last, p.scale, p.n_coarse, p.n_fine, p.thr2, t0, rt, c0, rc, b, ja_tab);
                    best = best_compact(b);
                }
            }
        }
        e_t = __reduce_or_sync(0xffffffffu, e_t);
        e_c = __reduce_or_sync(0xffffffffu, e_c);
        Best full = best_expand(best);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            Best o = best_shfl_down(full, d);
            best_merge(full, o);
        }
        if (lane == 0) {
            const long long s = (long long)g * p.group + tq;
            ksl::best_result(full, p.i_min + s, p.d_min + s, p.i_conj + s, p.d_conj + s);
            int et = e_t | record_err<kSharedTarget>(s_ct);
            if (kSharedTarget && p.target_err) et |= *p.target_err;
            p.err[s] = et | ((e_c | record_err<false>(s_cc)) << KSL_ERR_CHASER_SHIFT);
        }
        // next trace's pilot chunk: centred on the coarse segment of this trace's minimum -- if this trace and the one
        // before it bottomed out within 64 epochs of each other
        const long long j_min = __shfl_sync(0xffffffffu, full.j, 0);
        if (kMode == KSL_SCREEN_BRUTE || j_min == ksl::kNoIndex) {
            hint = kNoHint;
        } else {
            const int now = max(0, (int)ksl::up_i0(j_min, p.scale, p.n_coarse) - 15);
            const int before = (hint >= 0) ? hint : -1 - hint;      // (kNoHint decodes to a huge epoch: unrelated)
            hint = (abs(now - before) <= 64) ? now : -1 - now;
        }
        }   // traces of the group
    }
}

// target positions for the shared-target case: thread per coarse epoch
__global__ void __launch_bounds__(128) k_target_positions(const double *__restrict__ consts,
                                                          const double *__restrict__ epoch, const double *__restrict__ times,
                                                          long long n_coarse, double *__restrict__ pos, int *__restrict__ err) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_coarse) return;
    double r[3], v[3];
    // the same fast path as the per-sample targets of k_screen*, so that both forms of a batch see identical positions
    auto ld = [&](int q) { return __ldg(consts + q); };
    const ksl::FastView<decltype(ld)> view{ld, ld(KSL_C_ISIMP) != 0.0};
    const double ts = (times[k] - epoch[0]) * 1440.0;
    int e = 0;
    if (!ksl::sgp4_prop_fast(view, ts, r, ksl::kGridSinCos)) e = ksl::sgp4_prop<false>(ld, ts, r, v);
    pos[3 * k] = r[0]; pos[3 * k + 1] = r[1]; pos[3 * k + 2] = r[2];
    if (e) atomicOr(err, e);
}

