"""The drop-in host layer on the GPU: Conjunction.forward()/get_conjunction() and the dsgp4-style
calls, against the reference's golden trace (docs/notebooks/trace.pickle) and the oracle."""
import json
import os

import numpy as np
import pytest
import torch

import common
from oracle import c_oracle as C
from oracle import kessler_path as K

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def kb():
    import kessler_b200.model as M
    from kessler_b200 import ppl, sgp4, util
    from kessler_b200.tle import TLE, load_from_lines
    torch.cuda.set_device(0)
    return dict(M=M, ppl=ppl, sgp4=sgp4, util=util, TLE=TLE, load=load_from_lines)


def _golden_tles(kb, golden):
    t, c = golden["tles"]["t_tle"], golden["tles"]["c_tle"]
    return kb["TLE"](["0 " + t["name"], t["line1"], t["line2"]]), kb["TLE"](["0 " + c["name"], c["line1"], c["line2"]])


def test_dsgp4_surface(kb, golden):
    sgp4, TLE = kb["sgp4"], kb["TLE"]
    t_tle, c_tle = _golden_tles(kb, golden)
    sgp4.initialize_tle(t_tle)
    g = golden["tles"]["t_tle"]
    for name in ("_cc1", "_cc4", "_mdot", "_no_unkozai", "_t5cof", "_xlcof", "_nodedot", "_delmo"):
        assert getattr(t_tle, name) == pytest.approx(g[name], rel=1e-13)
    assert t_tle._isimp == g["_isimp"]
    st = sgp4.propagate(t_tle, torch.tensor(g["_t"], dtype=torch.float64))
    assert st.shape == (2, 3)
    el, batch = sgp4.initialize_tle([t_tle, c_tle])
    assert el.shape == (2, 7) and len(batch) == 2
    rv = sgp4.propagate_batch(batch, torch.tensor([g["_t"], 10.0], dtype=torch.float64))
    assert rv.shape == (2, 2, 3) and np.allclose(rv[0].numpy(), st.numpy(), rtol=0, atol=1e-9)
    c_ref, _ = C.sgp4_init(t_tle.elements()[:, None])
    rv_ref, _ = C.sgp4_prop(c_ref, np.array([g["_t"]]))
    assert np.abs(st.numpy().reshape(6)[:3] - rv_ref[0, 0, :3]).max() < 5e-8
    # deep space: initialize_tle raises, as dsgp4 does (forward() turns it into conj=False)
    d = dict(t_tle._data)
    d["mean_motion"] = 1.76e-4
    with pytest.raises(ValueError):
        sgp4.initialize_tle(TLE(d))
    tumin, mu, re, xke, j2, j3, j4, j3oj2 = sgp4.util.get_gravity_constants("wgs-84")
    assert float(mu) == pytest.approx(398600.5) and float(xke) == pytest.approx(g["_xke"], rel=1e-7)


def test_util_propagate_upsample_api(kb, golden):
    util = kb["util"]
    t_tle, _ = _golden_tles(kb, golden)
    s = golden["sites"]
    times = K.time_grid(s["time0"], 7.0, 6e5)
    out = util.propagate_upsample(t_tle, times, 100)
    assert isinstance(out, np.ndarray) and out.shape == (600000, 2, 3)
    c_ref, _ = C.sgp4_init(t_tle.elements()[:, None])
    rv_ref, _ = C.sgp4_prop(c_ref, (times[::100] - t_tle.date_mjd) * 1440.0)
    want = (C.upsample(rv_ref[0], len(times)) * 1e3).reshape(-1, 2, 3)
    assert np.abs(out[:, 0] - want[:, 0]).max() < 5e-5          # metres
    direct = util.propagate_upsample(t_tle, times[:500], 1)     # util.py:569-571: no interpolation
    rv1, _ = C.sgp4_prop(c_ref, (times[:500] - t_tle.date_mjd) * 1440.0)
    assert np.abs(direct.reshape(-1, 6) - rv1[0] * 1e3)[:, :3].max() < 5e-5
    up = util.upsample(torch.arange(12.0, dtype=torch.float64).reshape(6, 2), 11)
    assert np.array_equal(up.numpy(), K.upsample_torch(torch.arange(12.0, dtype=torch.float64).reshape(6, 2), 11).numpy())


def test_forward_reproduces_the_golden_trace(kb, golden):
    """Replay the golden run's two mean-anomaly draws through forward(): every deterministic site of the
    space segment must equal docs/notebooks/trace.pickle; the first CDM (Newton solve + 100-sample Monte
    Carlo covariance + TEME->ITRF) must agree with the pickled CDM within Monte Carlo noise."""
    M, ppl = kb["M"], kb["ppl"]
    s = golden["sites"]
    t_tle, c_tle = _golden_tles(kb, golden)
    torch.manual_seed(0)
    np.random.seed(0)
    model = M.Conjunction(time0=s["time0"], max_duration_days=s["max_duration_days"], t_tle=t_tle, c_tle=c_tle)
    model._replay = {"t_mean_anomaly": torch.tensor(s["t_mean_anomaly"], dtype=torch.float64),
                     "c_mean_anomaly": torch.tensor(s["c_mean_anomaly"], dtype=torch.float64)}
    tr = ppl.poutine.trace(model.forward).get_trace()
    v = lambda k: tr.nodes[k]["value"]
    assert bool(v("conj")) and not bool(v("prop_error"))
    assert int(v("i_conj")) == s["i_conj"] == 305767
    assert float(v("time_min")) == s["time_min"] and float(v("time_conj")) == s["time_conj"]
    assert abs(float(v("d_min")) - s["d_min"]) < 1e-4 and abs(float(v("d_conj")) - s["d_conj"]) < 1e-4      # metres
    assert int(v("max_ncdm")) == s["max_ncdm"] == 10
    for i in range(10):
        assert float(v(f"time_cdm_{i}")) == s[f"time_cdm_{i}"]
    for k in ("t_mean_motion", "t_eccentricity", "t_inclination", "t_raan", "c_argument_of_perigee", "c_b_star"):
        assert float(v(k)) == pytest.approx(s[k], rel=1e-15)
    cdms = tr.nodes["cdms"]["infer"]["cdms"]
    assert int(v("num_cdms")) == len(cdms) and 1 <= len(cdms) <= 10 and bool(v("cdm_issued_0"))
    g0 = golden["cdms"][0]
    assert cdms[0]["TCA"][:23] == g0["TCA"][:23] and cdms[0]["CREATION_DATE"][:23] == g0["CREATION_DATE"][:23]
    for oid in (0, 1):
        st = cdms[0].get_state(oid).reshape(6)
        # MC mean of 100 samples with sigma ~ 70 m (target) / ~ 2 km (chaser): hold to 5 sigma of the mean
        tol_km = 0.05 if oid == 0 else 1.5
        assert np.abs(st[:3] - np.array(g0["state"][oid][:3])).max() < tol_km
        got = np.diag(cdms[0].get_covariance(oid))
        ratio = got[:3] / np.array(g0["cov_rtn_diag"][oid][:3])
        assert (ratio > 0.4).all() and (ratio < 2.5).all(), ratio
    assert 0.0 <= cdms[0]["COLLISION_PROBABILITY"] <= 1.0
    assert float(v("cdm0_time_to_tca")) == pytest.approx(s["cdm0_time_to_tca"], abs=1e-9)
    assert float(v("cdm0_t_sma")) == pytest.approx(s["cdm0_t_sma"], rel=2e-3)     # from the ITRF state (sic)
    assert "obs_cdm0_t_sma" in tr.nodes and tr.nodes["t_tle"]["infer"]["t_tle"].line2 == golden["tles"]["t_tle"]["line2"]


def test_forward_rejections(kb, golden):
    """Altitude pre-filter and deep-space TLEs end the trace with conj=False (model.py:567-599)."""
    M, ppl, TLE = kb["M"], kb["ppl"], kb["TLE"]
    t_tle, c_tle = _golden_tles(kb, golden)
    s = golden["sites"]
    d = dict(c_tle._data)
    d["mean_motion"] = c_tle.mean_motion * 0.8
    far = TLE(d)
    tr = ppl.poutine.trace(M.Conjunction(time0=s["time0"], t_tle=t_tle, c_tle=far).forward).get_trace()
    assert not bool(tr.nodes["conj"]["value"]) and "time0" not in tr.nodes
    d = dict(c_tle._data)
    d["mean_motion"], d["eccentricity"] = 3.35e-4, 0.6          # period 312 min, perigee inside the target's shell
    deep = TLE(d)
    tr = ppl.poutine.trace(M.Conjunction(time0=s["time0"], t_tle=t_tle, c_tle=deep).forward).get_trace()
    assert not bool(tr.nodes["conj"]["value"]) and "time0" in tr.nodes and "c_prop_error" not in tr.nodes


def test_find_conjunction_function(kb, refnpz):
    M = kb["M"]
    for a, b, res in zip(refnpz["fc_tr0"], refnpz["fc_tr1"], refnpz["fc_res"]):
        i_min, d_min, i_conj, d_conj = M.find_conjunction(torch.from_numpy(a), torch.from_numpy(b), 5e3)
        assert i_min == int(res[0]) and float(d_min) == pytest.approx(res[1], rel=1e-13)
        if res[2] < 0:
            assert i_conj is None and d_conj is None
        else:
            assert i_conj == int(res[2]) and float(d_conj) == pytest.approx(res[3], rel=1e-13)


def test_mc_covariance_against_oracle(kb, golden):
    """propagate_uncertainty_monte_carlo with a fixed seed vs the same samples through the oracle."""
    M, sgp4, util = kb["M"], kb["sgp4"], kb["util"]
    t_tle, _ = _golden_tles(kb, golden)
    s = golden["sites"]
    model = M.Conjunction(time0=s["time0"], t_tle=t_tle, c_tle=t_tle)
    obs_time = s["time_cdm_3"]
    obs_tle, y = sgp4.newton_method(t_tle, obs_time)
    target = sgp4.propagate(t_tle, (obs_time - t_tle.date_mjd) * 1440.0).numpy()
    assert abs(obs_tle.date_mjd - obs_time) < 2e-8
    # the solved elements reproduce the state; the returned TLE carries the TLE-text quantisation (~10 m)
    st_y, _ = sgp4._states_at_epoch(y.numpy().reshape(6, 1), t_tle._bstar)
    assert np.abs(st_y[0] - target.reshape(6))[:3].max() < 1e-6
    assert np.abs(sgp4.propagate(obs_tle, 0.0).numpy() - target)[0].max() < 0.05
    cov_diag = np.array([1e-9, 1.115849341564346, 0.059309835843067, 1e-9, 1e-9, 1e-9]) ** 2
    state_m = target * 1e3
    torch.manual_seed(11)
    mean, cov = model.propagate_uncertainty_monte_carlo(state_m, cov_diag, s["time_min"], obs_tle)
    # same draws through the oracle
    torch.manual_seed(11)
    Cov = model.tle_covariance(state_m, cov_diag)
    mu = torch.tensor([float(obs_tle.semi_major_axis), obs_tle._ecco, obs_tle._inclo, obs_tle._nodeo, obs_tle._argpo, obs_tle._mo],
                      dtype=torch.float64)
    smp = torch.distributions.MultivariateNormal(loc=mu, covariance_matrix=torch.tensor(Cov)).sample((100,)).numpy()
    el = M.quantize_sgp4_inputs((398600.5e9 / smp[:, 0] ** 3) ** 0.5, np.maximum(smp[:, 1], 0.0), smp[:, 2], smp[:, 3], smp[:, 4],
                                smp[:, 5], np.full(100, obs_tle.b_star))
    st_ref, _ = C.tca_states(el, (s["time_min"] - obs_tle.date_mjd) * 1440.0)
    mean_ref, cov_ref = K.mc_covariance(st_ref.reshape(-1, 2, 3))
    assert np.abs(mean - mean_ref).max() < 1e-3
    assert np.allclose(cov, cov_ref, rtol=1e-6, atol=1e-9 * np.abs(cov_ref).max())
    assert 30.0 < np.sqrt(cov[0, 0]) < 200.0      # the 1e-10 regularisation of model.py:476-478 dominates: sigma ~ 70 m


def test_get_conjunction_on_population(kb):
    """ConjunctionSimplified + the 20-TLE sample population (docs notebook use case): batched rejection."""
    M = kb["M"]
    pop = json.load(open(os.path.join(ROOT, "tests", "golden", "population.json")))
    tles = kb["load"]([x for p in pop for x in ("0 " + p["name"], p["line1"], p["line2"])])
    torch.manual_seed(4)
    np.random.seed(4)
    model = M.ConjunctionSimplified(tles=tles, time0=60727.13899462018)
    batch = model.forward_batch(512)
    assert batch["conj"].shape == (512,) and batch["prefiltered"].sum() < 512
    k = np.nonzero(~batch["prefiltered"] & ~batch["prop_error"])[0][:6]
    # the fused screen agrees with the oracle on these draws
    times = model.time_grid()
    ep_t = np.array([tles[i].date_mjd for i in batch["t_raw"]["sampled_index"][k]])
    ep_c = np.array([tles[i].date_mjd for i in batch["c_raw"]["sampled_index"][k]])
    ref = C.screen(batch["elems_t"][:, k].copy(), ep_t, batch["elems_c"][:, k].copy(), ep_c, times[::100].copy(), len(times), 5e3)
    assert np.array_equal(ref["i_min"], batch["i_min"][k]) and np.array_equal(ref["i_conj"], batch["i_conj"][k])
    trace, iterations = model.get_conjunction(batch_size=2048)
    assert bool(trace.nodes["conj"]["value"]) and iterations >= 1
    cdms = trace.nodes["cdms"]["infer"]["cdms"]
    assert len(cdms) >= 1 and int(trace.nodes["i_conj"]["value"]) > 0
    assert float(trace.nodes["d_conj"]["value"]) < 5e3
    assert cdms[0].get_object(0, "OBJECT_NAME") == trace.nodes["t_tle"]["infer"]["t_tle"].name


def test_newton_and_partials_kernels(kb, golden, refnpz):
    """ksl_newton_elements (batched dsgp4.newton_method) and ksl_kepler_partials (util.py:187-302, batched)."""
    from kessler_b200 import engine
    sgp4, util = kb["sgp4"], kb["util"]
    # partials vs the reference's own autograd output (tests/golden/reference_functions.npz) and the host mirror
    st = refnpz["conv_states"].reshape(-1, 6)
    el, jac = engine.kepler_partials(st, 398600.5e9)
    el, jac = el.cpu().numpy(), jac.cpu().numpy()
    assert np.allclose(el, refnpz["conv_kepler"], rtol=1e-11, atol=1e-12)      # acos-based angles: cancellation near 0, pi
    ref = refnpz["conv_partials"]
    assert np.abs(jac - ref).max() / np.abs(ref).max() < 1e-14
    assert (np.abs(jac - ref) / np.maximum(np.abs(ref), 1e-300)).max() < 1e-11
    assert np.allclose(util.keplerian_cartesian_partials_batch(refnpz["conv_states"], 398600.5e9), jac, rtol=0, atol=0)
    # Newton: population TLEs at several observation times -> solved elements reproduce the target state
    pop = json.load(open(os.path.join(ROOT, "tests", "golden", "population.json")))
    tles = kb["load"]([x for p in pop for x in ("0 " + p["name"], p["line1"], p["line2"])])
    targets, bstars = [], []
    for t in tles:
        for dt_min in (-20000.0, -3000.0, 0.0, 1440.0):
            rv = sgp4.propagate(t, dt_min).numpy().reshape(6)
            if np.all(np.isfinite(rv)) and np.linalg.norm(rv[:3]) > 6400.0:
                targets.append(rv)
                bstars.append(t._bstar)
    targets, bstars = np.array(targets), np.array(bstars)
    assert len(targets) > 60
    y, resid, iters = engine.newton_elements(targets, bstars)
    y, resid, iters = y.cpu().numpy(), resid.cpu().numpy(), iters.cpu().numpy()
    assert np.isfinite(y).all() and (iters <= 12).all() and resid.max() < 1e-8
    el7 = np.concatenate([y.T, bstars[None, :]], axis=0)
    back, _ = C.tca_states(el7, 0.0)                       # oracle SGP4 of the solved elements, metres
    assert np.abs(back[:, :3] / 1e3 - targets[:, :3]).max() < 1e-6      # km: the north-star position tolerance
    assert np.abs(back[:, 3:] / 1e3 - targets[:, 3:]).max() < 1e-9
    # dsgp4-style call: new TLE at the observation epoch, text-quantised (~10 m)
    t = tles[0]
    new, yy = sgp4.newton_method(t, t.date_mjd + 0.3)
    assert abs(new.date_mjd - (t.date_mjd + 0.3)) < 2e-8 and yy.shape == (6,)
    want = sgp4.propagate(t, 0.3 * 1440.0).numpy()
    assert np.abs(sgp4.propagate(new, 0.0).numpy() - want)[0].max() < 0.05


def test_device_prior_sampling(kb):
    """ksl_sample_prior: same distributions as the host priors (two-sample KS per field), TLE-text quantisation
    bit-equal to the host formatter, draws keyed by the global index, and forward_batch() on top of it."""
    from scipy import stats
    from kessler_b200 import engine
    M = kb["M"]
    torch.manual_seed(21)
    model = M.Conjunction()
    params = model._device_prior()
    assert params is not None
    n = 200000
    raw, el, flag = engine.sample_prior(params, 5, 0, n)
    raw, el, flag = raw.cpu().numpy(), el.cpu().numpy(), flag.cpu().numpy()
    assert flag.mean() < 1e-4
    ok = flag == 0
    want = M.quantize_sgp4_inputs(raw[0], raw[2], raw[3], raw[5], raw[4], raw[1], raw[7])
    assert np.array_equal(el[:, ok], want[:, ok])                      # device quantisation == text round trip
    pr = model._prior_dict
    for k, name in enumerate(engine.PRIOR_FIELDS):
        host = pr[name + "_prior"].sample((n,)).double().numpy()
        ks = stats.ks_2samp(raw[k], host)
        assert ks.pvalue > 1e-4, (name, ks)
    assert (raw[0] >= 0).all() and (raw[0] <= 0.004).all() and (raw[2] >= 0).all() and (raw[2] <= 0.9).all()
    assert (raw[3] >= 0).all() and (raw[3] <= np.pi).all()
    a, _, _ = engine.sample_prior(params, 5, 0, 1000)
    b, _, _ = engine.sample_prior(params, 5, 1000, 1000)
    assert np.array_equal(np.concatenate([a.cpu().numpy(), b.cpu().numpy()], axis=1), raw[:, :2000])
    c, _, _ = engine.sample_prior(params, 6, 0, 1000)
    assert not np.array_equal(c.cpu().numpy(), raw[:, :1000])
    # forward_batch on device draws: reproducible under torch.manual_seed, consistent with the oracle
    torch.manual_seed(33)
    b1 = model.forward_batch(20000)
    torch.manual_seed(33)
    b2 = model.forward_batch(20000)
    assert np.array_equal(b1["elems_c"], b2["elems_c"]) and np.array_equal(b1["i_min"], b2["i_min"])
    assert 0.2 < b1["prefiltered"].mean() < 0.99 and b1["prop_error"].sum() > 0
    k = np.nonzero(~b1["prefiltered"] & ~b1["prop_error"])[0][:5]
    times = model.time_grid()
    ref = C.screen(b1["elems_t"][:, k].copy(), b1["epoch_t"][k], b1["elems_c"][:, k].copy(), b1["epoch_c"][k], times[::100].copy(),
                   len(times), 5e3)
    assert np.array_equal(ref["i_min"], b1["i_min"][k]) and np.array_equal(ref["i_conj"], b1["i_conj"][k])
    # replay of a device draw through make_target reproduces the same SGP4 inputs
    model._replay = model._replay_values(b1, int(k[0]))
    t_tle, c_tle = model.make_target(), model.make_chaser()
    model._replay = {}
    assert np.array_equal(t_tle.elements(), b1["elems_t"][:, k[0]]) and np.array_equal(c_tle.elements(), b1["elems_c"][:, k[0]])
    host_batch = model.forward_batch(2000, device_sampling=False)
    assert host_batch["elems_t"].shape == (7, 2000)


def test_default_prior_forward_and_batch(kb):
    """The full-prior model (no TLE given): traced forward() runs through TLE(dict) objects built at time0, and the
    batched front end classifies draws the same way (pre-filter / deep-space / screened)."""
    M, ppl = kb["M"], kb["ppl"]
    torch.manual_seed(8)
    np.random.seed(8)
    model = M.Conjunction()
    kinds = {"prefiltered": 0, "prop_error": 0, "screened": 0}
    for _ in range(12):
        tr = ppl.poutine.trace(model.forward).get_trace()
        assert "conj" in tr.nodes and "t_mean_motion" in tr.nodes and "c_b_star" in tr.nodes
        if "time0" not in tr.nodes:
            kinds["prefiltered"] += 1
        elif "c_prop_error" not in tr.nodes:
            kinds["prop_error"] += 1
        else:
            kinds["screened"] += 1
            assert "mc_prop_error" in tr.nodes and "num_cdms" in tr.nodes
    assert sum(kinds.values()) == 12 and kinds["prefiltered"] > 0
    b = model.forward_batch(4096)
    # replaying batched draws through forward() gives the same classification and, when screened, the same result
    picks = (list(np.nonzero(b["prefiltered"])[0][:2]) + list(np.nonzero(b["prop_error"])[0][:2]) +
             list(np.nonzero(~b["prefiltered"] & ~b["prop_error"])[0][:3]))
    for k in picks:
        model._replay = model._replay_values(b, int(k))
        tr = ppl.poutine.trace(model.forward).get_trace()
        model._replay = {}
        if b["prefiltered"][k]:
            assert "time0" not in tr.nodes and not bool(tr.nodes["conj"]["value"])
        elif b["prop_error"][k]:
            assert "time0" in tr.nodes and "c_prop_error" not in tr.nodes
        else:
            assert "c_prop_error" in tr.nodes and bool(tr.nodes["conj"]["value"]) == bool(b["conj"][k])
            if b["conj"][k]:
                assert int(tr.nodes["i_conj"]["value"]) == b["i_conj"][k]


def test_config0_thousand_prior_traces_of_the_default_pair(kb):
    """BASELINE configs[0]: Conjunction prior sampling, 1 000 traces of the default LEO pair (tests/test_model.py:20-26
    of the reference) over the 7-day window.  With both TLEs given the prior draws only the mean anomalies
    (model.py:269,308), quantised by the TLE text (1e-4 deg); every trace's space segment is checked against the C
    oracle's exhaustive enumeration of the 600 000 fine points on the very same quantised elements."""
    from kessler_b200 import workloads as W
    M = kb["M"]
    t_tle, c_tle = W.pair_a()
    torch.manual_seed(0)
    np.random.seed(0)
    model = M.Conjunction(t_tle=t_tle, c_tle=c_tle, time0=t_tle.date_mjd - 3.5)
    n = 1000
    b = model.forward_batch(n)
    assert not b["prefiltered"].any() and not b["prop_error"].any()
    # the draws: uniform mean anomalies on the text grid, everything else the given TLE's
    for raw, el, tle in ((b["t_raw"], b["elems_t"], t_tle), (b["c_raw"], b["elems_c"], c_tle)):
        m = np.asarray(raw["mean_anomaly"])
        assert m.min() >= 0.0 and m.max() < 2 * np.pi and 2.8 < m.mean() < 3.5
        deg = np.rad2deg(el[5])
        assert np.abs(deg * 1e4 - np.round(deg * 1e4)).max() < 1e-6
        assert np.all(el[1] == tle._ecco) and np.all(el[2] == tle._inclo) and np.all(el[6] == tle._bstar)
    times = model.time_grid()
    tc = np.ascontiguousarray(times[::100])
    ref = C.screen(np.ascontiguousarray(b["elems_t"]), b["epoch_t"], np.ascontiguousarray(b["elems_c"]), b["epoch_c"], tc, len(times),
                   5000.0)
    assert np.array_equal(b["i_min"], ref["i_min"]) and np.array_equal(b["i_conj"], ref["i_conj"])
    assert np.abs(b["d_min"] - ref["d_min"]).max() < 1e-3                         # metres (tolerance of record: 1 m)
    assert np.array_equal(b["conj"], ref["i_conj"] > 0)
    assert np.array_equal(b["time_min"], times[ref["i_min"]])                     # TCA: same fine-grid epoch, bit for bit


def test_torch_quantisation_on_the_device_is_bit_identical_to_the_text_path():
    """quantize_sgp4_inputs_torch (the GPU-resident ground segment) == quantize_sgp4_inputs (== TLE(dict), checked on the
    CPU) bit for bit on 2e5 random element sets -- including the divisions, which torch's CUDA kernels would otherwise
    turn into multiplications by a rounded reciprocal."""
    import math
    import kessler_b200.model as M
    rng = np.random.default_rng(12)
    n = 200000
    mm = rng.uniform(5e-4, 1.2e-3, n)
    ecc = np.abs(rng.normal(0.0, 0.01, n))
    ang = rng.uniform(0, 2 * math.pi, (4, n))
    bstar = rng.normal(0, 1e-3, n)
    want = M.quantize_sgp4_inputs(mm, ecc, ang[0], ang[1], ang[2], ang[3], bstar)
    dev = torch.device("cuda", 0)
    d = lambda a: torch.from_numpy(a).to(dev)
    got = M.quantize_sgp4_inputs_torch(d(mm), d(ecc), d(ang[0]), d(ang[1]), d(ang[2]), d(ang[3]), d(M._quantize_exp_field(bstar))).cpu().numpy()
    assert np.array_equal(got, want)
