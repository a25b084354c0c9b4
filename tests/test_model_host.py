"""Host-side logic of the drop-in layer (no GPU): TLE stand-in vs the golden pickle and the
reference's own unit tests, the util conversions vs outputs of the reference code, priors,
the trace recorder, constructor error behaviour."""
import json
import os

import numpy as np
import pytest
import torch

import common
from kessler_b200 import model as M
from kessler_b200 import ppl, util
from kessler_b200.cdm import ConjunctionDataMessage
from kessler_b200.observation_model import GNSS, Radar
from kessler_b200.tle import TLE, load_from_lines

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- dsgp4.TLE stand-in -----------------------------------------------------------------------
@pytest.mark.parametrize("who", ["t_tle", "c_tle"])
def test_tle_parses_golden_bit_exactly(golden, who):
    g = golden["tles"][who]
    t = TLE([g["line1"], g["line2"]])
    for k in ("_no_kozai", "_ecco", "_inclo", "_nodeo", "_argpo", "_mo", "_bstar", "_ndot", "_nddot"):
        assert getattr(t, k) == g[k], k
    assert t.date_mjd == g["date_mjd"] and t.mean_motion == g["mean_motion"]
    assert float(t.semi_major_axis) == g["semi_major_axis"]
    assert t.mean_motion_first_derivative == g["mean_motion_first_derivative"]
    for k in ("satellite_catalog_number", "classification", "international_designator", "ephemeris_type",
              "element_number", "revolution_number_at_epoch", "epoch_year", "epoch_days"):
        assert getattr(t, k) == g[k], k


def test_tle_update_is_the_text_quantisation_step(golden):
    """update({'mean_anomaly': sample}) keeps the raw value in the public field and derives _mo from
    the re-encoded text: golden sample 0.4932476473698829 -> _mo 0.49324749990611744 (28.2610 deg)."""
    for who, site in (("t_tle", "t_mean_anomaly"), ("c_tle", "c_mean_anomaly")):
        g = golden["tles"][who]
        t = TLE([g["line1"], g["line2"]]).copy()
        t.update({"mean_anomaly": torch.tensor(golden["sites"][site], dtype=torch.float64)})
        assert t._mo == g["_mo"] and t.line2 == g["line2"]
        assert t.mean_anomaly == golden["sites"][site]


def test_reference_test_make_chaser_make_target_roundtrip():
    """/root/reference/tests/test_model.py:19-39 (without the GPU part): copy()/update() reproduce the text."""
    for l1, l2 in common.PAIR_A:
        t = TLE([l1, l2])
        c = t.copy()
        c.update({"mean_anomaly": 1.2345})
        assert c.line2 != l2
        c.update({"mean_anomaly": float(t.mean_anomaly)})
        assert c.line1 == l1 and c.line2 == l2
    pop = json.load(open(os.path.join(ROOT, "tests", "golden", "population.json")))
    tles = load_from_lines([x for p in pop for x in ("0 " + p["name"], p["line1"], p["line2"])])
    assert len(tles) == 20 and tles[3].name == pop[3]["name"]
    for p, t in zip(pop, tles):
        c = t.copy()
        c.update({"mean_anomaly": t.mean_anomaly})
        assert c.line1 == p["line1"] and c.line2 == p["line2"]
        assert np.array_equal(t.elements(), [p[k] for k in ("no_kozai", "ecco", "inclo", "nodeo", "argpo", "mo", "bstar")])


def test_tle_dict_constructor_matches_vectorised_quantisation():
    torch.manual_seed(3)
    m = M.Conjunction()
    raw, el, ep = m.sample_elements("t", 300)
    for i in range(0, 300, 7):
        d = dict(m._STATIC_TLE_FIELDS, mean_motion_second_derivative=0.0, epoch_year=2020, epoch_days=143.40384230,
                 **{k: raw[k][i] for k in raw["_sites"]})
        t = TLE(d)
        assert np.array_equal(t.elements(), el[:, i])
        assert t.mean_anomaly == raw["mean_anomaly"][i]
    assert abs(ep[0] - m._time0) < 1e-7          # epoch = time0 through the 8-decimal day-of-year field


def test_apogee_perigee_and_prefilter():
    g = common.golden_trace()["tles"]["t_tle"]
    t = TLE([g["line1"], g["line2"]])
    assert t.apogee_alt() == pytest.approx(g["_alta"] * 6378.137e3, rel=1e-2)    # Kozai a (TLE method) vs un-Kozai a (sgp4init): 0.4 % apart
    assert t.apogee_alt() > t.perigee_alt() > 4.0e5
    m = M.Conjunction()
    lo = TLE(list(common.PAIR_A[0]))
    hi = lo.copy()
    hi.update({"mean_motion": lo.mean_motion * 0.8})
    assert m._prefilter_rejects(lo, hi) and m._prefilter_rejects(hi, lo) and not m._prefilter_rejects(lo, lo)


# ---- kessler.util mirror ---------------------------------------------------------------------------
def test_util_conversions_match_reference_outputs(refnpz):
    for s, R, rtn, back, kep, jac in zip(refnpz["conv_states"], refnpz["conv_rotation"], refnpz["conv_rtn"],
                                         refnpz["conv_back"], refnpz["conv_kepler"], refnpz["conv_partials"]):
        assert np.array_equal(util.rotation_matrix(s), R)
        got, Rm = util.from_cartesian_to_rtn(s)
        assert np.array_equal(got, rtn)
        assert np.array_equal(util.from_rtn_to_cartesian(got, Rm.T), back)
        ts = torch.tensor(s)
        k = np.array([float(x) for x in util.from_cartesian_to_keplerian(ts[0], ts[1], 398600.5e9)])
        assert np.allclose(k, kep, rtol=1e-15, atol=0)
        assert np.allclose(util.keplerian_cartesian_partials(s, 398600.5e9), jac, rtol=1e-12, atol=1e-300)


def test_reference_util_unit_tests():
    """/root/reference/tests/test_util.py:33-60 (TEME->ITRF known vector, RTN round trip) and :133-149."""
    state_TEME = np.array([[5094.18016210e3, 6127.64465950e3, 6380.34453270e3], [-4.746131487e3, 0.785818041e3, 5.531931288e3]])
    got = util.from_TEME_to_ITRF(state_TEME, 2433282.4235)
    want = np.array([[7377976.66089733, -3010674.50719708, 6380344.5327], [-900.58420598, 4224.28457883, 5531.931288]])
    assert np.allclose(want, got)
    assert abs(np.linalg.norm(state_TEME[0]) - np.linalg.norm(got[0])) < 0.05
    s = np.array([[1.0, 2.0, 3.4], [4.5, 6.2, 7.4]])
    rtn, R = util.from_cartesian_to_rtn(s)
    assert np.allclose(R.T @ rtn[0], s[0]) and np.allclose(R.T @ rtn[1], s[1])
    locs = torch.tensor([6.391167644720491e-05, 0.021593032643530245, 0.00089714840255561, 0.004279950440413096])
    scales = torch.tensor([9.155414891172675e-05, 0.08825398369676822, 0.0006834100202961423, 0.003464680595037937])
    probs = torch.tensor([0.36699734, 0.02809785, 0.39047779, 0.21442703])
    mix = M.MixtureSameFamily(M.Categorical(probs=probs), util.TruncatedNormal(loc=locs, scale=scales, low=-0.73577, high=0.68639))
    for x, lp in ((0.0001, 7.382209300994873), (0.001, 5.485926151275635), (0.01, 1.863307237625122), (0.1, -2.458112955093384)):
        assert round(mix.log_prob(torch.tensor(x)).item() - lp, 6) == 0
    assert util.from_jd_to_cdm_datetime_str(2469808.0229167)[:16] == "2050-01-01T12:33"
    idx, val = util.find_closest(np.array([1.0, 2.0, 4.0]), 2.9)
    assert idx == 1 and val == 2.0


def test_truncated_normal_matches_reference_logprob(refnpz):
    tn = util.TruncatedNormal(loc=torch.tensor([0.0010028142482042313, 0.00017592836171388626]),
                              scale=torch.tensor([0.00004670943133533001, 0.00003172305878251791]), low=0.0, high=0.004)
    assert np.allclose(tn.log_prob(torch.tensor(refnpz["tn_x"])).numpy(), refnpz["tn_logp"], rtol=1e-6)
    torch.manual_seed(0)
    s = tn.sample((2000,))
    assert s.shape == (2000, 2) and (s >= 0).all() and (s <= 0.004).all()
    assert abs(float(s[:, 0].mean()) - 0.0010028) < 1e-5


def test_observation_models():
    """/root/reference/tests/test_observation_model.py: observe = state + bias, covariance passes through."""
    st = np.arange(6.0).reshape(2, 3)
    for cls in (GNSS, Radar):
        ic = {"bias_xyz": np.ones((2, 3)) * 0.5, "covariance_rtn": np.arange(6.0) ** 2}
        s, c = cls(ic).observe(0.0, {"state_xyz": st})
        assert np.array_equal(s, st + 0.5) and np.array_equal(c, ic["covariance_rtn"])
        s, c = cls().observe(0.0, {"state_xyz": st})
        assert np.array_equal(s, st) and c.shape == (6,)


# ---- model: priors, recorder, constructor -----------------------------------------------------------
def test_constructor_error_behaviour_and_defaults(capsys):
    with pytest.raises(ValueError):
        M.Conjunction(pc_method="CHAN")
    with pytest.raises(ValueError):
        M.Conjunction(up_method="UT")
    with pytest.raises(ValueError):
        M.Conjunction(t_observing_instruments=[])
    m = M.Conjunction()
    assert "No observing instruments for target" in capsys.readouterr().out
    assert m._delta_time == 7.0 / 6e5 and len(m.time_grid()) == 600000
    assert set(m._prior_dict) == {"mean_motion_prior", "mean_anomaly_prior", "eccentricity_prior", "inclination_prior",
                                  "argument_of_perigee_prior", "raan_prior", "mean_motion_first_derivative_prior", "b_star_prior"}


def test_make_target_records_the_reference_sites():
    torch.manual_seed(1)
    m = M.Conjunction()
    tr = ppl.poutine.trace(lambda: (m.make_target(), m.make_chaser())).get_trace()
    want = [f"{w}_{s}" for w in "tc" for s in ("mean_motion", "mean_anomaly", "eccentricity", "inclination",
                                                "argument_of_perigee", "raan", "mean_motion_first_derivative",
                                                "mean_motion_second_derivative", "b_star")]
    assert [k for k in tr.nodes if not k.startswith("_")] == want
    t_tle, _ = tr.nodes["_RETURN"]["value"]
    assert t_tle.satellite_catalog_number == 43437 and t_tle.international_designator == "18100A"
    assert abs(t_tle.date_mjd - m._time0) < 1e-7
    # given TLEs: only the mean anomaly is drawn, the rest are deterministic sites (model.py:269-280)
    t0, c0 = TLE(list(common.PAIR_A[0])), TLE(list(common.PAIR_A[1]))
    m2 = M.Conjunction(t_tle=t0, c_tle=c0, t_observing_instruments=[GNSS()], c_observing_instruments=[Radar()])
    tr = ppl.poutine.trace(m2.make_target).get_trace()
    assert not tr.nodes["t_mean_anomaly"]["infer"].get("_deterministic")
    assert tr.nodes["t_raan"]["infer"]["_deterministic"] and float(tr.nodes["t_raan"]["value"]) == t0.raan
    tle = tr.nodes["_RETURN"]["value"]
    tle.update({"mean_anomaly": float(t0.mean_anomaly)})
    assert tle.line1 == common.PAIR_A[0][0] and tle.line2 == common.PAIR_A[0][1]


def test_replay_reproduces_a_batched_draw():
    torch.manual_seed(5)
    m = M.Conjunction()
    raw_t, el_t, _ = m.sample_elements("t", 8)
    raw_c, el_c, _ = m.sample_elements("c", 8)
    m._replay = m._replay_values(dict(t_raw=raw_t, c_raw=raw_c), 3)
    t_tle, c_tle = m.make_target(), m.make_chaser()
    assert np.array_equal(t_tle.elements(), el_t[:, 3]) and np.array_equal(c_tle.elements(), el_c[:, 3])


def test_simplified_batched_sampling_matches_sequential_semantics():
    pop = json.load(open(os.path.join(ROOT, "tests", "golden", "population.json")))
    tles = load_from_lines([x for p in pop for x in ("0 " + p["name"], p["line1"], p["line2"])])
    torch.manual_seed(2)
    m = M.ConjunctionSimplified(tles=tles)
    raw, el, ep = m.sample_elements("t", 500)
    assert raw["sampled_index"].min() >= 1          # sic: Categorical(tensor(range(n))) gives index 0 weight 0 (model.py:786)
    k = 17
    m._replay = {"t_mean_anomaly": torch.tensor(raw["mean_anomaly"][k]), "t_sampled_index": torch.tensor(raw["sampled_index"][k])}
    tle = m.make_target()
    assert np.array_equal(tle.elements(), el[:, k]) and tle.date_mjd == ep[k]
    assert tle.name == tles[raw["sampled_index"][k]].name


def test_cdm_record_roundtrip():
    c = ConjunctionDataMessage()
    c.set_state(0, np.array([[7000., 0, 0], [0, 7.5, 0]]))
    assert np.isnan(c["MISS_DISTANCE"])
    c.set_state(1, np.array([[7000., 1, 0.5], [0, -7.5, 0]]))
    assert c["MISS_DISTANCE"] == pytest.approx(np.hypot(1000.0, 500.0))
    assert np.allclose(c.get_state_relative(), [[0, 1000, 500], [0, -15000, 0]])
    cov = np.arange(36.).reshape(6, 6)
    cov = cov + cov.T
    c.set_covariance(1, cov)
    assert np.array_equal(c.get_covariance(1), cov) and c.get_object(1, "CT_R") == cov[1, 0] and c["OBJECT2_CNDOT_T"] == cov[5, 1]
    d = c.copy()
    d.set_object(0, "X", 1.0)
    assert c.get_object(0, "X") == 7000.0
    with pytest.raises(ValueError):
        c.set_object(2, "X", 0.0)


def test_cdm_kvn_wire_format_roundtrip(tmp_path):
    """The reference's own CDM test (tests/test_cdm.py: load -> save -> load gives an equal message) on a message in
    the reference's KVN layout, plus what that layout is: 37-column keys, obligatory keys always written, unknown keys
    and comments ignored, exponents as %.3E."""
    text = """
        COMMENT an example in the layout the reference writes
        CCSDS_CDM_VERS                        = 1.0
        CREATION_DATE                         = 2021-03-04T10:20:30.000
        ORIGINATOR                            = KESSLER_SOFTWARE
        MESSAGE_ID                            = KESSLER_SOFTWARE_0001
        TCA                                   = 2021-03-07T01:02:03.500
        MISS_DISTANCE                         = 715.5
        RELATIVE_SPEED                        = 14123.25
        COLLISION_PROBABILITY                 = 3.25E-05
        COLLISION_PROBABILITY_METHOD          = MC
        SOME_FUTURE_KEY                       = ignored
        OBJECT                                = OBJECT1
        OBJECT_DESIGNATOR                     = 12345.0
        CATALOG_NAME                          = SATCAT
        OBJECT_NAME                           = TARGET
        INTERNATIONAL_DESIGNATOR              = 1998-067A
        EPHEMERIS_NAME                        = NONE
        COVARIANCE_METHOD                     = CALCULATED
        MANEUVERABLE                          = N/A
        REF_FRAME                             = ITRF
        SEDR                                  = 7.305E-05
        X                                     = -731.97
        Y                                     = -1871.2
        Z                                     = -6876.4
        X_DOT                                 = -1.1177
        Y_DOT                                 = -7.044
        Z_DOT                                 = 2.0366
        CR_R                                  = 38.33
        CT_R                                  = 93.6
        CT_T                                  = 3410.0
        CDRG_R                                = 1.5
        OBJECT                                = OBJECT2
        OBJECT_DESIGNATOR                     = 54321.0
        OBJECT_NAME                           = CHASER
        X                                     = -731.5
        Y                                     = -1871.0
        Z                                     = -6876.0
        X_DOT                                 = 1.0
        Y_DOT                                 = 7.0
        Z_DOT                                 = -2.0
        CR_R                                  = 1.0e-9
    """
    p1, p2 = str(tmp_path / "a.cdm.kvn.txt"), str(tmp_path / "b.cdm.kvn.txt")
    with open(p1, "w") as f:
        f.write(text)
    a = ConjunctionDataMessage.load(p1)
    a.save(p2)
    b = ConjunctionDataMessage(p2)                      # constructor form of load (cdm.py:61-62)
    assert a == b and hash(a) == hash(b) and a.kvn() == b.kvn()
    assert a["MISS_DISTANCE"] == 715.5 and a["COLLISION_PROBABILITY"] == 3.25e-5 and a["OBJECT1_OBJECT_NAME"] == "TARGET"
    assert a["OBJECT2_X"] == -731.5 and a.get_object(0, "SEDR") == 7.305e-5 and a.get_object(0, "CDRG_R") == 1.5
    assert np.isnan(a.get_object(1, "CT_T")) and a.to_dict()["OBJECT2_CT_T"] is None
    lines = a.kvn().splitlines()
    assert all(line[37:39] == " =" for line in lines)
    assert "COLLISION_PROBABILITY".ljust(37) + " = 3.250E-05" in lines and "MISS_DISTANCE".ljust(37) + " = 715.5" in lines
    assert "CT_T".ljust(37) + " =" in lines                      # obligatory but unset for object 2: written empty
    assert not any(line.startswith("SOME_FUTURE_KEY") or line.startswith("COMMENT") for line in lines)
    assert len(a.kvn(show_all=True).splitlines()) == 5 + 20 + 2 * (21 + 18 + 6 + 45)
    missing = a.validate()
    assert ("Object2 metadata", "CATALOG_NAME") in missing and ("header", "ORIGINATOR") not in missing
    c = b.copy()
    c["OBJECT2_Z"] = -6875.0
    assert c != a
    # a message produced by set_state / set_covariance survives the wire format with its derived fields
    m = ConjunctionDataMessage()
    m.set_header("ORIGINATOR", "KESSLER_SOFTWARE")
    m.set_header("MESSAGE_ID", "X1")
    m.set_relative_metadata("TCA", "2021-03-07T01:02:03.500000")
    m.set_state(0, np.array([[7000., 0, 0], [0, 7.5, 0]]))
    m.set_state(1, np.array([[7000., 1, 0.5], [0, -7.5, 0]]))
    cov = np.diag([1.0, 2.0, 3.0, 1e-6, 2e-6, 3e-6])
    m.set_covariance(0, cov)
    m.set_covariance(1, 2 * cov)
    m.save(p1)
    back = ConjunctionDataMessage.load(p1)
    assert back["MISS_DISTANCE"] == pytest.approx(m["MISS_DISTANCE"]) and np.allclose(back.get_covariance(1), 2 * cov)
    assert np.allclose(back.get_state_relative(), m.get_state_relative())
    frame = m.to_dataframe()
    assert frame.shape[0] == 1 and frame["OBJECT1_X"][0] == 7000.0 and "OBJECT2_CNDOT_NDOT" in frame.columns


def test_package_level_names_of_the_reference():
    """`import kessler` exposes seed, ConjunctionDataMessage / CDM, Event, GNSS, Radar, model, util
    (/root/reference/kessler/__init__.py): the same names resolve on this package."""
    import kessler_b200 as kb
    from kessler_b200 import cdm, observation_model
    assert kb.ConjunctionDataMessage is cdm.ConjunctionDataMessage is kb.CDM and kb.Event is cdm.Event
    assert kb.GNSS is observation_model.GNSS and kb.Radar is observation_model.Radar
    assert callable(kb.seed) and hasattr(kb.model, "Conjunction") and hasattr(kb.util, "from_cartesian_to_rtn")
    with pytest.raises(AttributeError):
        kb.no_such_name


def test_tle_copy_keeps_the_text_quantised_sgp4_inputs():
    """dsgp4's copy rebuilds through TLE(dict); a copied / deep-copied TLE must propagate like the original."""
    import copy as _copy
    from kessler_b200.tle import TLE
    t = TLE(list(common.PAIR_A[0]))
    t.update({"mean_anomaly": 0.4932476473698829, "inclination": 0.69851234567, "eccentricity": 0.00080961234})
    assert t.mean_anomaly == 0.4932476473698829 and t._mo != t.mean_anomaly          # raw field kept, SGP4 input quantised
    for c in (t.copy(), _copy.deepcopy(t), TLE(t)):
        assert np.array_equal(c.elements(), t.elements())
        assert c.line1 == t.line1 and c.line2 == t.line2 and c.date_mjd == t.date_mjd
    c = t.copy()
    assert c.mean_anomaly == t.mean_anomaly and c._data is not t._data
    c.update({"mean_anomaly": 1.0})
    assert t._mo != c._mo and t.mean_anomaly == 0.4932476473698829                  # independent objects


def test_conjunction_simplified_exclude_object_name_batched_sampling():
    """model.py:813-816 per draw: a target carrying the excluded name draws its chaser index from
    Categorical(tensor(include_idxs)); the batched sampler no longer refuses the option."""
    import torch
    from kessler_b200.model import ConjunctionSimplified
    from kessler_b200.tle import TLE
    el, ep = common.population_elems()
    tles = []
    for k in range(6):
        t = TLE(list(common.PAIR_A[k % 2]))
        t.name = ("STARLINK-%d" % k) if k in (1, 4) else ("OBJ-%d" % k)
        tles.append(t)
    m = ConjunctionSimplified(tles, exclude_object_name="STARLINK")
    assert m._include_idxs == [0, 2, 3, 5]
    torch.manual_seed(0)
    raw_t, el_t, _ = m.sample_elements('t', 4000)
    raw_c, el_c, _ = m.sample_elements('c', 4000)
    ti, ci = raw_t['sampled_index'], raw_c['sampled_index']
    assert ti.min() >= 1 and ti.max() == 5                 # Categorical(range(6)): index 0 has probability 0 (reference quirk)
    excl = np.isin(ti, (1, 4))
    assert excl.any() and (~excl).any()
    # include_idxs = [0, 2, 3, 5] as probabilities: values in {1, 2, 3}; the unconditional prior reaches 4 and 5
    assert set(np.unique(ci[excl])) <= {1, 2, 3} and ci[~excl].max() == 5
    assert el_c.shape == (7, 4000)


def test_cdm_header_dates_follow_the_reference():
    """cdm.py:193-203: day-of-year CREATION_DATE converted, malformed refused."""
    from kessler_b200.cdm import ConjunctionDataMessage
    c = ConjunctionDataMessage()
    c.set_header("CREATION_DATE", "2022-068T19:10:13.500")
    assert c._header["CREATION_DATE"] == "2022-03-09T19:10:13.500"
    c.set_header("CREATION_DATE", "2022-03-09T19:10:13.5")
    with pytest.raises(RuntimeError):
        c.set_header("CREATION_DATE", "09/03/2022 19:10")
    with pytest.raises(RuntimeError):
        c.set_header("CREATION_DATE", "2022-03-09T19:10:13")
    with pytest.raises(ValueError):
        c.set_header("NOT_A_KEY", 1)
