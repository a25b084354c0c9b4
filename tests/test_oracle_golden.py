"""Pins the oracle (oracle/) against the reference's golden vectors BEFORE it is
trusted as the checker: docs/notebooks/trace.pickle (via trace_golden.json),
tests/test_util.py:105-131 (TLE 34427 -> poliastro elements), and outputs of the
reference's own non-dsgp4 functions (reference_functions.npz)."""
import numpy as np
import pytest
import torch

import common
from oracle import c_oracle as C
from oracle import kessler_path as K
from oracle import sgp4_ref as S

COEFFS = [k for k in S.CONST_NAMES if k not in ("isimp", "cosio", "sinio", "aobase", "pad")] + ["a", "alta", "altp"]


def _golden_name(k):
    return "_no_unkozai" if k == "no_unkozai" else "_" + k


@pytest.mark.parametrize("who", ["t_tle", "c_tle"])
def test_tle_text_to_elements_bit_exact(golden, who):
    t = golden["tles"][who]
    p = S.parse_tle_lines(t["line1"], t["line2"])
    for k, gk in zip(("no_kozai", "ecco", "inclo", "nodeo", "argpo", "mo", "bstar"), common.EL_KEYS):
        assert p[k] == t[gk], (k, p[k], t[gk])
    assert p["epoch_mjd"] == t["date_mjd"]
    assert S.datetime_to_mjd(__import__("datetime").datetime(*t["epoch_datetime"])) == t["date_mjd"]


def test_gravity_constants(golden):
    t = golden["tles"]["t_tle"]
    assert S.XKE == t["_xke"] and S.TUMIN == t["_tumin"] and S.MU == t["_mu"] and S.RE == t["_radiusearthkm"]
    assert S.J2 == t["_j2"] and S.J3 == t["_j3"] and S.J4 == t["_j4"] and S.J3OJ2 == t["_j3oj2"]


@pytest.mark.parametrize("who", ["t_tle", "c_tle"])
@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_sgp4init_coefficients(golden, who, impl):
    t = golden["tles"][who]
    el = [t[k] for k in common.EL_KEYS]
    if impl == "numpy":
        c, err = S.sgp4_init(*el)
        c = {k: float(v) for k, v in c.items()}
    else:
        arr, err = C.sgp4_init(np.array(el)[:, None])
        c = {k: float(arr[i, 0]) for i, k in enumerate(S.CONST_NAMES)}
    assert int(np.asarray(err).reshape(-1)[0]) == 0
    checked = 0
    for k in COEFFS:
        if k in c and _golden_name(k) in t:
            assert c[k] == pytest.approx(t[_golden_name(k)], rel=2e-15, abs=0), k
            checked += 1
    assert checked >= (28 if impl == "numpy" else 25)
    assert int(c["isimp"]) == t["_isimp"]


@pytest.mark.parametrize("who", ["t_tle", "c_tle"])
def test_mean_elements_at_last_epoch_bit_exact(golden, who):
    """_am,_em,_Om,_om,_mm,_nm at the pickled _t -- also proves floor-mod wrapping."""
    t = golden["tles"][who]
    c, _ = S.sgp4_init(*[t[k] for k in common.EL_KEYS])
    _, _, _, mean = S.sgp4_propagate(c, np.array([t["_t"]]), return_mean=True)
    for name, val in zip(("_am", "_em", "_Om", "_om", "_mm", "_nm"), mean):
        assert float(val[0]) == t[name], name


@pytest.mark.parametrize("use_torch", [True, False])
def test_full_trace_reproduces_golden(golden, use_torch):
    et, ept, ec, epc, s = common.golden_pair()
    times = K.time_grid(s["time0"], s["max_duration_days"], 6e5)
    assert len(times) == 600000
    r = K.screen_trace(et, ept, ec, epc, times, 100, 5e3, use_torch=use_torch)
    assert r["i_conj"] == s["i_conj"] == 305767
    assert times[r["i_conj"]] == s["time_conj"]
    assert times[r["i_min"]] == s["time_min"] and r["i_min"] == 305768
    assert abs(r["d_min"] - s["d_min"]) < 1e-6          # metres  (1e-9 km)
    assert abs(r["d_conj"] - s["d_conj"]) < 1e-6
    max_ncdm, _, tt = K.cdm_schedule(times, s["time0"], times[r["i_min"]])
    assert max_ncdm == s["max_ncdm"] == 10
    for i in range(max_ncdm):
        assert tt[i] == s[f"time_cdm_{i}"]


def test_c_oracle_trace_reproduces_golden():
    et, ept, ec, epc, s = common.golden_pair()
    times = K.time_grid(s["time0"], 7.0, 6e5)
    r = C.screen(et[:, None], [ept], ec[:, None], [epc], times[::100], len(times), 5e3)
    assert r["i_min"][0] == 305768 and r["i_conj"][0] == s["i_conj"]
    assert abs(r["d_min"][0] - s["d_min"]) < 1e-6 and abs(r["d_conj"][0] - s["d_conj"]) < 1e-6


def test_reference_unit_test_tle_34427():
    """tests/test_util.py:105-131: sgp4(tle, 0) -> Kepler elements vs poliastro, 5 places."""
    p = S.parse_tle_lines(*common.TLE_34427)
    c, _ = S.sgp4_init(*[p[k] for k in ("no_kozai", "ecco", "inclo", "nodeo", "argpo", "mo", "bstar")])
    r, v, _ = S.sgp4_propagate(c, np.array([0.0]))
    r, v = r[0] * 1e3, v[0] * 1e3
    a, e, i, Om, om, M = K.cartesian_to_keplerian(r, v, common.MU_POLIASTRO)
    for got, want in zip((a, e, i, Om, om, M), common.KEPLER_34427):
        assert abs(got - want) < 5e-6          # assertAlmostEqual(places=5)
    assert abs(a - common.KEPLER_34427[0]) < 1e-7   # 1.9e-9 m in the scratch check


def test_c_vs_numpy_sgp4_on_population_and_catalog():
    el, _ = common.population_elems()
    el = np.concatenate([el, common.synthetic_catalog(200)], axis=1)
    c1, e1 = C.sgp4_init(el)
    c2, e2 = S.sgp4_init(*el)
    a2 = S.consts_to_array(c2)
    ok = e1 == 0
    assert (e1 == e2).all()
    assert np.allclose(c1[:, ok], a2[:, ok], rtol=1e-13, atol=0)
    ts = np.linspace(-3.0e4, 3.0e4, 241)
    rv, ea = C.sgp4_prop(c1, ts)
    r, v, eb = S.sgp4_propagate({k: a2[i][:, None] for i, k in enumerate(S.CONST_NAMES)}, ts[None, :], couple_kepler=False)
    eb = np.bitwise_or.reduce(eb, axis=1)
    assert (ea[ok] == eb[ok]).all()
    good = ok & (ea == 0)
    assert good.sum() > 150
    assert np.abs(rv[good][..., :3] - r[good]).max() < 2e-8      # km; stale-sin/cos semantics differ by <= 7e-9
    assert np.abs(rv[good][..., 3:] - v[good]).max() < 1e-10


def test_isimp_branch_in_fixtures():
    el, _ = common.population_elems()
    c, _ = C.sgp4_init(el)
    assert c[S.CONST_NAMES.index("isimp")].sum() == 3     # 3 of the 20 population TLEs


# ---- Kessler's own functions, against outputs of the reference code itself ---------------
@pytest.mark.parametrize("tag", list("abcde"))
def test_upsample_matches_reference(refnpz, tag):
    """util.upsample outputs produced by the reference's own code.  The C oracle (and torch
    here) must be BIT-equal; the NumPy variant rounds w0*x0 separately (<= 1 ulp off)."""
    x, n_f = refnpz[f"upsample_{tag}_in"], int(refnpz[f"upsample_{tag}_nf"])
    want = refnpz[f"upsample_{tag}_out"]
    for exact, y in ((False, K.upsample_numpy(x, n_f)), (True, C.upsample(x, n_f)),
                     (True, K.upsample_torch(torch.from_numpy(x), n_f).numpy())):
        if f"upsample_{tag}_idx" in refnpz:
            assert np.allclose(y.sum(axis=0), refnpz[f"upsample_{tag}_sum"], rtol=1e-12)
            y = y[refnpz[f"upsample_{tag}_idx"]]
        if exact:
            assert np.array_equal(y, want)
        else:
            assert np.allclose(y, want, rtol=0, atol=8e-12)


def test_upsample_index_quirk_is_float32_floor():
    """79 of the 600 000 default fine indices take torch's float32-floor path;
    random data makes a wrong source index a gross error, and the C oracle is bit-equal to torch."""
    i0, i1, w0, w1 = K.upsample_index_weights(6000, 600000)
    j = np.arange(600000, dtype=np.float64)
    naive = np.floor(j * (5999 / 599999)).astype(np.int64)
    assert (i0 != naive).sum() == 79
    x = np.random.default_rng(0).standard_normal((6000, 2)) * 7000.0
    want = K.upsample_torch(torch.from_numpy(x), 600000).numpy()
    assert np.array_equal(C.upsample(x, 600000), want)
    assert np.abs(K.upsample_numpy(x, 600000) - want).max() < 8e-12
    for n_in, n_out in ((50, 13), (5, 1), (3, 3), (1, 4)):      # down-sampling, single output, identity, single input
        x = np.random.default_rng(n_in).standard_normal((n_in, 3))
        assert np.array_equal(C.upsample(x, n_out), K.upsample_torch(torch.from_numpy(x), n_out).numpy())


def test_find_conjunction_matches_reference(refnpz):
    for a, b, res in zip(refnpz["fc_tr0"], refnpz["fc_tr1"], refnpz["fc_res"]):
        for fn in (K.find_conjunction, K.find_conjunction_numpy, C.find_conjunction):
            i_min, d_min, i_conj, d_conj = fn(a, b, 5e3)
            i_conj = -1 if i_conj is None else i_conj
            assert i_min == int(res[0]) and i_conj == int(res[2])
            assert d_min == pytest.approx(res[1], rel=1e-12)
            if i_conj >= 0:
                assert d_conj == pytest.approx(res[3], rel=1e-12)


def test_find_conjunction_nan_is_minimum():
    a = np.random.default_rng(0).standard_normal((50, 3)) * 1e4
    b = a + 1e5
    b[17, 1] = np.nan
    b[30, 0] = np.nan
    for fn in (K.find_conjunction, K.find_conjunction_numpy, C.find_conjunction):
        i_min, d_min, i_conj, _ = fn(a, b, 5e3)
        assert i_min == 17 and np.isnan(d_min) and (i_conj is None or i_conj == -1)


def test_time_grid_and_find_closest(refnpz):
    times = K.time_grid(58991.90384230018)
    assert len(times) == int(refnpz["times_len"])
    assert np.array_equal(times[refnpz["times_probe_idx"]], refnpz["times_probe"])
    for q, (idx, val) in zip(refnpz["find_closest_q"], refnpz["find_closest_res"]):
        k = int(np.argmin(abs(times - q)))
        assert k == int(idx) and times[k] == val


def test_rotation_and_mc_cov(refnpz):
    for s, R in zip(refnpz["conv_states"], refnpz["conv_rotation"]):
        assert np.allclose(K.rotation_matrix(s), R, rtol=0, atol=1e-15)
    rng = np.random.default_rng(1)
    st = refnpz["conv_states"][0].reshape(1, 6) + rng.standard_normal((100, 6)) * np.array([50, 80, 20, .05, .08, .02])
    mean, cov = K.mc_covariance(st.reshape(-1, 2, 3))
    mean_c, cov_c = C.mc_cov(st)
    assert np.allclose(mean.reshape(-1), mean_c, rtol=1e-14)
    assert np.allclose(cov, cov_c, rtol=1e-9, atol=1e-12)


def test_pc_count_oracles_agree():
    rng = np.random.default_rng(2)
    t = rng.standard_normal((3, 700)) * 60.0
    c = rng.standard_normal((3, 900)) * 60.0 + np.array([[30.0], [0.0], [0.0]])
    n_all = K.pc_count_all_pairs(t.T, c.T, 70.0)
    assert n_all == C.pc_count(t, c, 70.0, 0)
    assert 0 < n_all < 700 * 900
    assert K.pc_count_elementwise(t.T[:700], c.T[:700], 70.0) == C.pc_count(t, c[:, :700].copy(), 70.0, 1)
