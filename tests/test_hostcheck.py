"""CPU check of the DEVICE code's algebra: kessler_b200/csrc/{sgp4,screen}_device.cuh are
__host__ __device__, so tests/hostcheck/hostcheck.cpp compiles them with g++ and the
restructured SGP4 (4 sincos + small-angle rotations instead of 16 trig calls + atan2) and
the closed-form segment search are compared with the oracle here, without a GPU.  The GPU
tests (-m gpu) repeat the comparison through the C ABI on the real kernels."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

import common
from oracle import c_oracle as C
from oracle import kessler_path as K
from oracle.sgp4_ref import CONST_NAMES as S_NAMES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = ctypes.POINTER(ctypes.c_double)
I32 = ctypes.POINTER(ctypes.c_int32)
I64 = ctypes.POINTER(ctypes.c_int64)


def _p(a, t):
    return a.ctypes.data_as(t)


@pytest.fixture(scope="module")
def hc(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hostcheck") / "libhostcheck.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-mfma", "-ffp-contract=off", "-x", "c++",
                           os.path.join(ROOT, "tests", "hostcheck", "hostcheck.cpp"), "-o", so])
    return ctypes.CDLL(so)


def hc_init(hc, el):
    el = np.ascontiguousarray(el)
    n = el.shape[1]
    c = np.zeros((36, n))
    e = np.zeros(n, np.int32)
    hc.hc_sgp4_init(_p(el, D), ctypes.c_int64(n), _p(c, D), _p(e, I32))
    return c, e


def hc_prop(hc, c, ts, vel=1):
    c, ts = np.ascontiguousarray(c), np.ascontiguousarray(ts)
    n, m = c.shape[1], ts.shape[-1]
    rv = np.zeros((n, m, 6))
    e = np.zeros(n, np.int32)
    hc.hc_sgp4_prop(_p(c, D), ctypes.c_int64(n), _p(ts, D), ctypes.c_int64(m), 0, vel, _p(rv, D), _p(e, I32))
    return rv, e


def hc_screen(hc, ct, ept, cc, epc, tc, n_fine, thr, mode):
    ct, cc, tc = np.ascontiguousarray(ct), np.ascontiguousarray(cc), np.ascontiguousarray(tc)
    ept, epc = np.ascontiguousarray(ept, dtype=np.float64), np.ascontiguousarray(epc, dtype=np.float64)
    n_t, n = ct.shape[1], cc.shape[1]
    out = dict(i_min=np.zeros(n, np.int64), d_min=np.zeros(n), i_conj=np.zeros(n, np.int64), d_conj=np.zeros(n),
               err=np.zeros(n, np.int32))
    hc.hc_screen(_p(ct, D), _p(ept, D), ctypes.c_int64(n_t), _p(cc, D), _p(epc, D), ctypes.c_int64(n), _p(tc, D),
                 ctypes.c_int64(len(tc)), ctypes.c_int64(n_fine), ctypes.c_double(thr), mode,
                 _p(out["i_min"], I64), _p(out["d_min"], D), _p(out["i_conj"], I64), _p(out["d_conj"], D), _p(out["err"], I32))
    return out


def test_device_sgp4init_matches_oracle(hc):
    el, _ = common.population_elems()
    el = np.concatenate([el, common.synthetic_catalog(300)], axis=1)
    c1, e1 = C.sgp4_init(el)
    c2, e2 = hc_init(hc, el)
    assert (e1 == e2).all()
    ok = e1 == 0
    floor = 1e-6 * np.median(np.abs(c1[:, ok]), axis=1, keepdims=True)   # sin(pi) and what is proportional to it
    for name in ("sinio", "xlcof", "aycof", "omgcof"):                         # (xlcof divides it by the 1.5e-12 guard)
        floor[list(S_NAMES).index(name)] = max(floor[list(S_NAMES).index(name)], 1e-3)
    rel = np.abs(c1[:, ok] - c2[:, ok]) / np.maximum(np.maximum(np.abs(c1[:, ok]), floor), 1e-300)
    # (sincos_grid and cbrt instead of the oracle's libm sincos / pow: an ulp of cos(i) is 8e-13 of 1 - cos^2 at i = 1 deg)
    assert rel.max() < 2e-12, {n: float(r) for n, r in zip(S_NAMES, rel.max(axis=1)) if r > 1e-14}
    for k in (0, 7, 34):   # no, mdot, ao: what accumulates phase
        assert rel[k].max() < 1e-15, S_NAMES[k]


def test_device_fast_sgp4init_matches_oracle(hc):
    """sgp4_init_fast (Monte Carlo kernels): the oracle's constants to a few ulp per factor, the phase-carrying ones to
    1e-15, identical branch flags; and fast init + fast state (sgp4_init_state_fast) against the oracle's init + state
    one and five days from the epoch."""
    el, _ = common.population_elems()
    gt, _, gc, _, _ = common.golden_pair()
    el = np.concatenate([gt[:, None], gc[:, None], el, common.synthetic_catalog(2000, seed=17)], axis=1)
    rng = np.random.default_rng(3)
    mc = common.perturbed_samples(gt, 3000, 4)          # what the kernels actually see: a cloud around one TLE
    el = np.ascontiguousarray(np.concatenate([el, mc], axis=1))
    n = el.shape[1]
    c1, e1 = C.sgp4_init(el)
    c2, e2 = np.zeros((36, n)), np.zeros(n, np.int32)
    hc.hc_sgp4_init_fast(_p(el, D), ctypes.c_int64(n), _p(c2, D), _p(e2, I32))
    assert (e1 == e2).all()
    ok = e1 == 0
    assert np.array_equal(c1[31, ok], c2[31, ok])       # isimp
    floor = 1e-6 * np.median(np.abs(c1[:, ok]), axis=1, keepdims=True)
    for name in ("sinio", "xlcof", "aycof", "omgcof"):
        floor[list(S_NAMES).index(name)] = max(floor[list(S_NAMES).index(name)], 1e-3)
    rel = np.abs(c1[:, ok] - c2[:, ok]) / np.maximum(np.maximum(np.abs(c1[:, ok]), floor), 1e-300)
    # (ao - s4 amplifies the two ulp of cbrt(.)^2 up to 1e3-fold in tsi^4 for perigees just above the s4 altitude)
    assert rel.max() < 2e-11, {nm: float(r) for nm, r in zip(S_NAMES, rel.max(axis=1)) if r > 1e-13}
    for k in (0, 7, 34):                                # no, mdot, ao
        assert rel[k].max() < 1e-15, S_NAMES[k]
    for k in (8, 9):                                    # argpdot, nodedot: absolute, as phase after 5e4 min
        assert np.abs(c1[k, ok] - c2[k, ok]).max() * 5.0e4 < 1e-11, S_NAMES[k]
    for ts in (1440.0, -7200.0):
        ref, ea = C.tca_states(el, ts)                  # metres
        rv, err, used = np.zeros((n, 6)), np.zeros(n, np.int32), np.zeros(n, np.int32)
        hc.hc_init_state_fast(_p(el, D), ctypes.c_int64(n), ctypes.c_double(ts), _p(rv, D), _p(err, I32), _p(used, I32))
        good = ok & (ea == 0) & (used == 1)
        assert good.sum() > 0.85 * ok.sum() and used[-3000:].all()
        # metres.  Tolerance of record: 1e-6 km = 1e-3 m.  The catalog's extremes (perigee 150 km: drag terms amplify the
        # ulps of ao 1e3-fold; e ~ 0.12) reach 1e-4 m after five days; the cloud around a real TLE stays below 5e-6 m
        assert np.abs(rv[good, :3] * 1e3 - ref[good, :3]).max() < 2e-4
        cloud = good & (np.arange(n) >= n - 3000)
        assert np.abs(rv[cloud, :3] * 1e3 - ref[cloud, :3]).max() < 5e-6
        assert np.median(np.abs(rv[good, :3] * 1e3 - ref[good, :3])) < 5e-8
        assert np.abs(rv[good, 3:] * 1e3 - ref[good, 3:]).max() < 2e-7        # m/s (same extremes)
        assert np.abs(rv[cloud, 3:] * 1e3 - ref[cloud, 3:]).max() < 5e-9
        assert (err == e1).all()


def test_device_sgp4_matches_oracle(hc):
    el, _ = common.population_elems()
    el = np.concatenate([el, common.synthetic_catalog(300)], axis=1)
    c1, e1 = C.sgp4_init(el)
    ts = np.linspace(-5.0e4, 5.0e4, 401)
    rv1, ea = C.sgp4_prop(c1, ts)
    rv2, eb = hc_prop(hc, c1, ts)          # same constants: isolates the propagation algebra
    good = (e1 == 0) & (ea == 0)
    assert (ea == eb)[e1 == 0].all()
    assert good.sum() > 200
    assert np.abs(rv1[good][..., :3] - rv2[good][..., :3]).max() < 1e-9     # km  (tolerance of record: 1e-6 km)
    assert np.abs(rv1[good][..., 3:] - rv2[good][..., 3:]).max() < 1e-12    # km/s
    rv3, _ = hc_prop(hc, c1, ts, vel=0)
    assert np.array_equal(rv3[..., :3][good], rv2[..., :3][good])           # position-only path is the same arithmetic
    # bad-physics lanes: same non-finite pattern
    bad = (e1 == 0) & (ea != 0)
    assert np.array_equal(np.isnan(rv1[bad][..., :3]), np.isnan(rv2[bad][..., :3]))


def test_device_fast_path_matches_oracle(hc):
    """sgp4_prop_fast (what the screening kernels run) against the exact path, same constants: the share of epochs it
    accepts and its deviation -- typically 4e-12 km; a one-ulp flip of the ~1e3 rad mean longitude costs 2e-9 km."""
    el, _ = common.population_elems()
    el = np.concatenate([el, common.synthetic_catalog(600)], axis=1)
    c1, e1 = C.sgp4_init(el)
    ts = np.linspace(-2.0e4, 2.0e4, 401)
    rv1, ea = C.sgp4_prop(c1, ts)
    n, m = c1.shape[1], len(ts)
    pos, err, used = np.zeros((n, m, 3)), np.zeros(n, np.int32), np.zeros((n, m), np.int32)
    hc.hc_sgp4_position(_p(np.ascontiguousarray(c1), D), ctypes.c_int64(n), _p(ts, D), ctypes.c_int64(m), _p(pos, D),
                        _p(err, I32), _p(used, I32))
    good = (e1 == 0) & (ea == 0)
    assert (err == ea)[e1 == 0].all()
    dev = np.abs(pos - rv1[..., :3]).max(-1)
    fast = (used == 1) & good[:, None]
    assert fast.sum() > 0.9 * good.sum() * m
    assert dev[fast].max() < 1e-8 and np.median(dev[fast]) < 2e-11        # km
    slow = (used == 0) & good[:, None]                                       # fallback = the exact path
    assert dev[slow].max() < 1e-9
    # the three-sweep Kepler solve at the top of its validity range (e <= 0.12), and low-drag records only: accepted,
    # and as accurate as everywhere else
    hi_e = good & (el[1] > 0.05) & (el[1] < 0.115) & (np.abs(el[6]) < 1e-3)
    assert hi_e.sum() >= 10
    assert fast[hi_e].mean() > 0.95
    assert dev[hi_e][fast[hi_e]].max() < 1e-8 and np.median(dev[hi_e][fast[hi_e]]) < 2e-11


def test_device_near_circular_form_matches_oracle(hc):
    """The two-sweep Kepler solve k_screen_warp selects for near-circular records (e + drag growth < 0.0095): selected for
    the real LEO TLEs and their Monte Carlo clouds, accepted at (nearly) every epoch, and as accurate as the general form."""
    el, _ = common.population_elems()
    gt, _, gc, _, _ = common.golden_pair()
    cat = common.synthetic_catalog(1500, seed=23)
    rng = np.random.default_rng(9)
    edge = np.repeat(gt[:, None], 400, 1)                   # around the class boundary: e from 0.004 to 0.012
    edge[1] = rng.uniform(0.004, 0.012, 400)
    edge[5] = rng.uniform(0, 2 * np.pi, 400)
    calm = el[:, 10].copy()                                 # a low-drag member of the population (B* = 2.5e-4)
    el = np.ascontiguousarray(np.concatenate([gt[:, None], gc[:, None], el, common.perturbed_samples(calm, 500, 1), cat, edge], axis=1))
    c1, e1 = C.sgp4_init(el)
    ts = np.linspace(-1.2e4, 1.2e4, 301)
    rv1, ea = C.sgp4_prop(c1, ts)
    n, m = c1.shape[1], len(ts)
    pos, used, small = np.zeros((n, m, 3)), np.zeros((n, m), np.int32), np.zeros(n, np.int32)
    hc.hc_sgp4_position_small(_p(np.ascontiguousarray(c1), D), ctypes.c_int64(n), _p(ts, D), ctypes.c_int64(m), ctypes.c_double(1.2e4),
                              _p(pos, D), _p(used, I32), _p(small, I32))
    good = (e1 == 0) & (ea == 0)
    # most real TLEs (not the high-drag ones over this +-8 day window: their drag angle needs the general form's kernels)
    # and the whole cloud around a calm one
    assert small[:22].sum() >= 12 and small[22:522].all()
    sm = good & (small == 1)
    assert sm.sum() > 800
    dev = np.abs(pos - rv1[..., :3]).max(-1)
    assert used[sm].mean() > 0.995                                      # the hint is right: (almost) no fallback evaluations
    acc = (used == 1) & sm[:, None]
    assert dev[acc].max() < 1e-8 and np.median(dev[acc]) < 2e-11        # km, as the general form
    assert dev[good].max() < 1e-8
    # the boundary cloud: whatever the hint said, every epoch is either accepted within tolerance or redone exactly
    assert 50 < small[-400:].sum() < 350


def test_device_fast_state_matches_oracle(hc):
    """sgp4_state (K2 / K4 / K6 / K9: fast path with velocity over the 36 public constants, exact fallback)."""
    el, _ = common.population_elems()
    el = np.concatenate([el, common.synthetic_catalog(600)], axis=1)
    c1, e1 = C.sgp4_init(el)
    ts = np.linspace(-2.0e4, 2.0e4, 201)
    rv1, ea = C.sgp4_prop(c1, ts)
    n, m = c1.shape[1], len(ts)
    rv, err, used = np.zeros((n, m, 6)), np.zeros(n, np.int32), np.zeros((n, m), np.int32)
    hc.hc_sgp4_state(_p(np.ascontiguousarray(c1), D), ctypes.c_int64(n), _p(ts, D), ctypes.c_int64(m), _p(rv, D), _p(err, I32),
                     _p(used, I32))
    good = (e1 == 0) & (ea == 0)
    assert (err == ea)[e1 == 0].all()
    fast = (used == 1) & good[:, None]
    assert fast.sum() > 0.9 * good.sum() * m
    dp = np.abs(rv[..., :3] - rv1[..., :3]).max(-1)
    dv = np.abs(rv[..., 3:] - rv1[..., 3:]).max(-1)
    assert dp[fast].max() < 1e-8 and np.median(dp[fast]) < 2e-11          # km
    assert dv[fast].max() < 1e-11 and np.median(dv[fast]) < 3e-14         # km/s (one ulp of a 1e3 rad angle: 2e-12)
    slow = (used == 0) & good[:, None]
    assert dp[slow].max() < 1e-9 and dv[slow].max() < 1e-12


def test_f32_skip_test_is_conservative(hc):
    """segment_skippable_f32 (k_screen_warp's inner loop) may only skip what the FP64 bound skips: random segments,
    segments whose bound sits within 1e-3 relative of the running minimum, huge and non-finite coordinates."""
    rng = np.random.default_rng(5)
    n = 400000
    d0 = rng.normal(size=(n, 3)) * rng.choice([1.0, 50.0, 3000.0, 14000.0], size=(n, 1))
    d1 = d0 + rng.normal(size=(n, 3)) * rng.choice([0.01, 1.0, 300.0, 1500.0], size=(n, 1))
    b = d1 - d0
    a0, a1, a2 = (d0 * d0).sum(1), (d0 * b).sum(1), (b * b).sum(1)
    lam = np.clip(-a1 / np.maximum(a2, 1e-300), 0, 1)
    lb = (a0 + 2 * a1 * lam + a2 * lam * lam) * 1e6                       # m^2
    cur = lb * (1 + rng.choice([-1e-3, -1e-5, -1e-7, 0, 1e-7, 1e-5, 1e-3, 0.5, -0.5], size=n))
    d0[:50] *= 1e200
    d1[50:100] = np.nan
    d0[100:150] = np.inf
    cur[150:200] = np.inf
    hc.hc_skip_compare.restype = ctypes.c_int64
    for conj_open in (0, 1):
        s32, s64 = ctypes.c_int64(), ctypes.c_int64()
        bad = hc.hc_skip_compare(_p(np.ascontiguousarray(d0), D), _p(np.ascontiguousarray(d1), D), _p(cur, D), ctypes.c_int64(n),
                                 ctypes.c_double(25e6), conj_open, ctypes.byref(s32), ctypes.byref(s64))
        assert bad == 0
        assert 0 < s32.value <= s64.value and s32.value > 0.3 * s64.value    # ... and it still prunes (7 of 9 cases sit on the boundary)


def test_near_circular_selection_by_record_and_by_time_limit_agree(hc):
    """fast_record_is_small_e(record, tmax) (the screen kernels: once per trace) and tmax < fast_record_small_e_tmax(record)
    (K2: once per record, compared per tile) are the same decision away from the boundary; records that are circular to
    the last digit of a TLE (e < 2e-6) never select the form -- it declines em < 1e-6, the general form clamps there."""
    el = common.synthetic_catalog(4000, seed=23)
    el[1, 100:140] = np.linspace(0.0, 4.0e-6, 40)
    c, e0 = C.sgp4_init(el)
    tmax = np.array([0.0, 1.0, 720.0, 5040.0, 2.0e4, 6.0e4])
    n, m = el.shape[1], len(tmax)
    is_small, t_limit = np.zeros((n, m), np.int32), np.zeros(n)
    hc.hc_small_e_selection(_p(np.ascontiguousarray(c), D), ctypes.c_int64(n), _p(tmax, D), ctypes.c_int64(m), _p(is_small, I32),
                            _p(t_limit, D))
    ok = e0 == 0
    by_limit = tmax[None, :] < t_limit[:, None]
    near = np.abs(tmax[None, :] - t_limit[:, None]) <= 1e-9 * np.maximum(tmax[None, :], 1.0)
    assert np.array_equal(is_small[ok].astype(bool) | near[ok], by_limit[ok] | near[ok])
    assert 0.2 < is_small[ok, 3].mean() < 0.9                       # a week: most of the LEO catalog qualifies
    assert not is_small[100:120].any() and (t_limit[100:120] <= 0).all()          # e <= 2e-6
    assert is_small[130:140, 1].all()                                             # e >= 3e-6, one minute from the epoch


def test_lazy_search_bounds_are_bounds_at_any_magnitude(hc):
    """segment_bounds_f32: lb <= min of the segment's quadratic <= ub whenever it claims to have bounded the segment -- also
    for separations of 1e10 .. 1e15 km (the garbage of a decayed, flagged object), where a1^2 overflows FP32: it used to
    return ub = -inf there, which pruned the rest of the trace (benchmarks/fuzz_screen.py, seed 508877)."""
    rng = np.random.default_rng(9)
    n = 300000
    mag = 10.0 ** rng.uniform(-1, 16, size=(n, 1))
    d0 = rng.normal(size=(n, 3)) * mag
    d1 = d0 + rng.normal(size=(n, 3)) * mag * 10.0 ** rng.uniform(-4, 1, size=(n, 1))
    hc.hc_bounds_check.restype = ctypes.c_int64
    nb = ctypes.c_int64()
    for dl in (0.02, 0.25):
        bad = hc.hc_bounds_check(_p(np.ascontiguousarray(d0), D), _p(np.ascontiguousarray(d1), D), ctypes.c_int64(n), ctypes.c_double(dl),
                                 ctypes.byref(nb))
        assert bad == 0
        assert nb.value > 0.5 * n                 # (it declines only where FP32 cannot hold the squares at all)
    big = np.abs(d0).max(axis=1) > 1e10
    assert big.sum() > 0.2 * n


def test_sincos_grid_table_is_the_generated_one():
    """kessler_b200/csrc/sincos_grid_table.cuh is the output of tools/make_sincos_grid.py (correctly rounded with mpmath),
    and its entries are sin / cos of the grid angles (libm on the rounded angle agrees to the angle's rounding, 1e-15)."""
    import re
    path = os.path.join(ROOT, "kessler_b200", "csrc", "sincos_grid_table.cuh")
    src = open(path).read()
    a = src.index("kGridSinCos[2 * kGridPoints] = {", src.index("#else"))
    vals = np.array([float(x) for x in re.findall(r"-?\d+\.\d+(?:e-?\d+)?", src[a:src.index("};", a)])])
    assert vals.shape == (1024,)
    ang = 2 * np.pi * np.arange(512) / 512
    assert np.abs(vals[0::2] - np.sin(ang)).max() < 1e-15 and np.abs(vals[1::2] - np.cos(ang)).max() < 1e-15
    pytest.importorskip("mpmath")
    gen = subprocess.check_output([sys.executable, os.path.join(ROOT, "tools", "make_sincos_grid.py")]).decode()
    assert gen == src


def test_device_wrap_is_python_mod(hc):
    hc.hc_wrap_twopi.restype = ctypes.c_double
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-4e3, 4e3, 20000), [0.0, -0.0, 2 * np.pi, -2 * np.pi, 6.283185307179586 * 7, 1e-300, -1e-300]])
    for x in xs:
        assert hc.hc_wrap_twopi(ctypes.c_double(x)) == float(np.mod(x, 2 * np.pi)), x


def _golden_batch(n, seed):
    et, ept, ec, epc, s = common.golden_pair()
    times = K.time_grid(s["time0"], 7.0, 6e5)
    rng = np.random.default_rng(seed)
    elt, elc = np.repeat(et[:, None], n, 1), np.repeat(ec[:, None], n, 1)
    for arr in (elt, elc):
        m = np.round(rng.uniform(0, 360, n), 4) * np.pi / 180     # TLE-text quantised mean anomalies (model.py:269,308)
        m[0] = arr[5, 0]
        arr[5] = m
    for i in range(1, n // 3):                                     # near-golden: conjunctions persist
        elt[5, i] = et[5] + rng.normal() * 2e-5
        elc[5, i] = ec[5] + rng.normal() * 2e-5
    return elt, np.full(n, ept), elc, np.full(n, epc), times, s


@pytest.mark.parametrize("mode", [1, 0])
def test_device_screen_matches_oracle_default_grid(hc, mode):
    elt, ept, elc, epc, times, s = _golden_batch(24, 5)
    tc = np.ascontiguousarray(times[::100])
    ref = C.screen(elt, ept, elc, epc, tc, len(times), 5e3)
    ct, _ = C.sgp4_init(elt)
    cc, _ = C.sgp4_init(elc)
    r = hc_screen(hc, ct, ept, cc, epc, tc, len(times), 5e3, mode)
    assert r["i_min"][0] == 305768 and r["i_conj"][0] == s["i_conj"]
    assert np.array_equal(r["i_min"], ref["i_min"]) and np.array_equal(r["i_conj"], ref["i_conj"])
    # metres.  The fast path keeps the reference's roundings of the ~1e3 rad mean longitude but not every rounding
    # of the small terms added to it, so a sum can land one ulp (2e-13 rad = 1.6e-6 m) away: 1e-5 m = 1e-8 km.
    assert np.nanmax(np.abs(r["d_min"] - ref["d_min"])) < 1e-5
    assert np.nanmedian(np.abs(r["d_min"] - ref["d_min"])) < 1e-7
    both = ref["i_conj"] >= 0
    assert both.sum() >= 5
    assert np.abs(r["d_conj"][both] - ref["d_conj"][both]).max() < 1e-5
    assert np.isnan(r["d_conj"][~both]).all()


@pytest.mark.parametrize("n_coarse,n_fine", [(61, 6001), (7, 61), (2, 9), (300, 29901), (257, 25601), (256, 2551), (1, 5), (40, 13), (5, 1)])
def test_device_screen_ragged_grids(hc, n_coarse, n_fine):
    """Analytic == brute == oracle on odd grid shapes (chunk boundaries at 255/256 epochs,
    single epoch, down-sampling, single output)."""
    el, ep = common.population_elems()
    rng = np.random.default_rng(n_coarse * 1000 + n_fine)
    pairs = [(i, j) for i in range(20) for j in range(20) if i != j]
    sel = [pairs[k] for k in rng.choice(len(pairs), 12, replace=False)]
    elt = np.stack([el[:, i] for i, _ in sel], 1)
    elc = np.stack([el[:, j] for _, j in sel], 1)
    ept, epc = ep[[i for i, _ in sel]], ep[[j for _, j in sel]]
    t0 = ep.mean() - 3.5
    tc = t0 + np.arange(n_coarse) * (7.0 / max(n_coarse, 1))
    thr = 4.0e6                                     # large threshold so that crossings exist on these random pairs
    ref = C.screen(elt, ept, elc, epc, tc, n_fine, thr)
    ct, _ = C.sgp4_init(elt)
    cc, _ = C.sgp4_init(elc)
    for mode in (1, 0):
        r = hc_screen(hc, ct, ept, cc, epc, tc, n_fine, thr, mode)
        assert np.array_equal(r["i_min"], ref["i_min"]), mode
        assert np.array_equal(r["i_conj"], ref["i_conj"]), mode
        assert np.allclose(r["d_min"], ref["d_min"], rtol=0, atol=1e-5)


def test_device_screen_degenerate_and_nan(hc):
    """Identical objects (d == 0 everywhere -> first index), and NaN states (torch.argmin
    returns the first NaN index; NaN < thr is False)."""
    el, ep = common.population_elems()
    n_coarse, n_fine = 300, 29901
    tc = ep[0] + np.arange(n_coarse) * 0.001
    same_t, same_c = el[:, :1].copy(), el[:, :1].copy()
    bad = common.synthetic_catalog(16)[:, [5, 9]]       # lanes that go non-finite (see tests/common.py)
    elt = np.concatenate([same_t, el[:, 1:3]], 1)
    elc = np.concatenate([same_c, bad], 1)
    ept = np.array([ep[0], ep[1], ep[2]])
    epc = np.array([ep[0], ep[1] + 30.0, ep[2] + 30.0])
    ref = C.screen(elt, ept, elc, epc, tc, n_fine, 5e3)
    ct, _ = C.sgp4_init(elt)
    cc, _ = C.sgp4_init(elc)
    assert ref["i_min"][0] == 0 and ref["d_min"][0] == 0.0 and ref["i_conj"][0] == 0
    assert np.isnan(ref["d_min"][1:]).any()
    for mode in (1, 0):
        r = hc_screen(hc, ct, ept, cc, epc, tc, n_fine, 5e3, mode)
        assert np.array_equal(r["i_min"], ref["i_min"]) and np.array_equal(r["i_conj"], ref["i_conj"])
        assert np.array_equal(np.isnan(r["d_min"]), np.isnan(ref["d_min"]))
