"""Batched ground segment (kessler_b200.events) against the per-event path and the oracle."""
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
def setup():
    import kessler_b200.model as M
    from kessler_b200 import engine, events, sgp4, util
    from kessler_b200.tle import load_from_lines
    torch.cuda.set_device(0)
    pop = json.load(open(os.path.join(ROOT, "tests", "golden", "population.json")))
    tles = load_from_lines([x for p in pop for x in ("0 " + p["name"], p["line1"], p["line2"])])
    torch.manual_seed(3)
    np.random.seed(3)
    model = M.ConjunctionSimplified(tles=tles, time0=60727.13899462018)
    batch = model.forward_batch(16384)
    hits = np.nonzero(batch["conj"])[0][:6]
    assert len(hits) >= 3
    return dict(M=M, engine=engine, events=events, sgp4=sgp4, util=util, model=model, batch=batch, hits=hits, tles=tles)


def test_states_at_fine_equals_propagate_upsample_rows(setup):
    eng, model, batch, hits, tles = setup["engine"], setup["model"], setup["batch"], setup["hits"], setup["tles"]
    times = model.time_grid()
    h = hits[0]
    tle = tles[int(batch["t_raw"]["sampled_index"][h])].copy()
    tle.update({"mean_anomaly": batch["t_raw"]["mean_anomaly"][h]})
    full = setup["sgp4"].propagate_upsample_device(tle, times, 100).cpu().numpy()
    idx = np.array([0, 1, 99, 100, 28571, 305767, 599998, 599999])
    c, _ = eng.sgp4_init(batch["elems_t"][:, [h]])
    got, err = eng.states_at_fine(c, np.zeros(len(idx), dtype=np.int64), np.full(len(idx), batch["epoch_t"][h]), idx,
                                  times[::100].copy(), len(times))
    # same algorithm in two kernels: equal up to fma-contraction / last-bit effects (<< 1e-6 km)
    assert np.abs(got.cpu().numpy() - full[idx])[:, :3].max() < 2e-5 and np.abs(got.cpu().numpy() - full[idx])[:, 3:].max() < 1e-7
    assert int(err.max()) == 0


def test_pc_mvn_batched_counts_and_draws(setup):
    eng = setup["engine"]
    rng = np.random.default_rng(0)
    n_cdm, n_pts = 5, 4000
    mean = rng.standard_normal((n_cdm, 2, 3)) * 30.0
    A = rng.standard_normal((n_cdm, 2, 3, 3)) * 25.0
    cov = A @ np.transpose(A, (0, 1, 3, 2)) + 10.0 * np.eye(3)
    L = np.linalg.cholesky(cov)
    counts, pts = eng.pc_mvn_batched(mean, L, n_pts, 70.0, seed=5, return_points=True)
    counts, pts = counts.cpu().numpy(), pts.cpu().numpy()
    for k in range(n_cdm):
        assert counts[k] == C.pc_count(pts[k, 0], pts[k, 1], 70.0, 0)            # identical samples: bit-exact
        for o in range(2):
            assert np.abs(pts[k, o].mean(axis=1) - mean[k, o]).max() < 5 * np.sqrt(np.diag(cov[k, o]).max() / n_pts)
            assert np.abs(np.cov(pts[k, o]) - cov[k, o]).max() < 0.12 * np.abs(cov[k, o]).max()
    assert 0 < counts.sum() < n_cdm * n_pts * n_pts
    again = eng.pc_mvn_batched(mean, L, n_pts, 70.0, seed=5).cpu().numpy()
    assert np.array_equal(again, counts)                                          # reproducible
    shifted = eng.pc_mvn_batched(mean[2:], L[2:], n_pts, 70.0, seed=5, cdm_id0=2).cpu().numpy()
    assert np.array_equal(shifted, counts[2:])                                    # keyed by the CDM id, not the launch


def test_batched_ground_segment_matches_per_event_path(setup):
    """Deterministic stages of generate_cdm_batch vs Conjunction.forward() replayed on the same draws, and the
    Monte Carlo stage vs the oracle on the same normal deviates."""
    M, ev_mod, model, batch, hits, tles = setup["M"], setup["events"], setup["model"], setup["batch"], setup["hits"], setup["tles"]
    from kessler_b200 import ppl
    E = len(hits)
    rng = np.random.default_rng(1)
    max_i = 21
    t_new = rng.uniform(size=(E, max_i)) < 0.96
    c_new = rng.uniform(size=(E, max_i)) < 0.4
    res0 = ev_mod.generate_cdm_batch(model, batch, hits, seed=11, t_new_obs=t_new, c_new_obs=c_new)
    I = res0["issued"].shape[1]
    n_items = len(res0["items"]["event"])
    z = rng.standard_normal((n_items, 100, 6))
    res = ev_mod.generate_cdm_batch(model, batch, hits, seed=11, t_new_obs=t_new[:, :I], c_new_obs=c_new[:, :I], z_elements=z)
    it = res["items"]
    assert np.array_equal(it["event"], res0["items"]["event"]) and it["newton_resid"].max() < 1e-8
    times = model.time_grid()
    for e, h in enumerate(hits):
        # replay the same draw through the traced forward(); inject the same observation flags
        model._replay = model._replay_values(batch, int(h))
        for i in range(1, int(res["max_ncdm"][e])):
            model._replay[f"t_new_obs_{i}"] = torch.tensor(float(t_new[e, i]))
            model._replay[f"c_new_obs_{i}"] = torch.tensor(float(c_new[e, i]))
        torch.manual_seed(100 + e)
        tr = ppl.poutine.trace(model.forward).get_trace()
        model._replay = {}
        assert bool(tr.nodes["conj"]["value"])
        assert int(tr.nodes["max_ncdm"]["value"]) == res["max_ncdm"][e]
        assert float(tr.nodes["time_min"]["value"]) == res["tca_mjd"][e]
        cdms = tr.nodes["cdms"]["infer"]["cdms"]
        issued = [i for i in range(int(res["max_ncdm"][e])) if res["issued"][e, i]]
        assert len(cdms) == len(issued) == int(tr.nodes["num_cdms"]["value"])
        for i in range(int(res["max_ncdm"][e])):
            assert float(tr.nodes[f"time_cdm_{i}"]["value"]) == res["time_cdm"][i]
            assert bool(tr.nodes[f"cdm_issued_{i}"]["value"]) == bool(res["issued"][e, i])
        batched = ev_mod.to_cdms(model, res, e)
        for cdm_seq, cdm_bat in zip(cdms, batched):
            assert cdm_seq["TCA"] == cdm_bat["TCA"] and cdm_seq["CREATION_DATE"] == cdm_bat["CREATION_DATE"]
            for o in (0, 1):
                # same Newton solve / partials / element covariance; different normal deviates: equal within the
                # 100-sample Monte Carlo noise of the mean (sigma/10) and of the variances
                sig = np.sqrt(np.diag(cdm_seq.get_covariance(o))[:3].max())
                assert np.abs(cdm_seq.get_state(o)[0] - cdm_bat.get_state(o)[0]).max() * 1e3 < 0.8 * sig + 1.0
                ratio = np.diag(cdm_bat.get_covariance(o))[:3] / np.diag(cdm_seq.get_covariance(o))[:3]
                assert (ratio > 0.3).all() and (ratio < 3.0).all()
            assert 0.0 <= cdm_bat["COLLISION_PROBABILITY"] <= 1.0
    # Monte Carlo stage on identical deviates vs the oracle (a few items)
    for k in rng.choice(n_items, 5, replace=False):
        L = np.linalg.cholesky(it["cov_tle"][k])
        raw_mm = it["newton_y"][k, 0] / 60.0
        mean_els = np.array([(398600.5e9 / raw_mm ** 2) ** (1 / 3), *it["obs_elements"][1:6, k]])
        smp = mean_els + z[k] @ L.T
        el = M.quantize_sgp4_inputs((398600.5e9 / smp[:, 0] ** 3) ** 0.5, np.maximum(smp[:, 1], 0.0), smp[:, 2], smp[:, 3], smp[:, 4],
                                    smp[:, 5], np.full(100, it["obs_elements"][6, k]))
        tsince = (res["tca_mjd"][it["event"][k]] - it["obs_epoch_mjd"][k]) * 1440.0
        st_ref, _ = C.tca_states(el, tsince)
        mean_ref, cov_ref = K.mc_covariance(st_ref.reshape(-1, 2, 3))
        assert np.abs(it["mean_state_tca_m"][k] - mean_ref.reshape(6))[:3].max() < 1e-3
        assert np.allclose(it["cov_rtn"][k], cov_ref, rtol=1e-6, atol=1e-9 * np.abs(cov_ref).max())


def test_generate_events_driver(setup):
    ev_mod, model = setup["events"], setup["model"]
    out, screened = ev_mod.generate_events(model, 12, batch_size=8192, seed=1)
    n = sum(len(r["hits"]) for r in out)
    assert n == 12 and screened >= 8192
    for r in out:
        assert r["issued"][:, 0].all()                                   # first CDM always issued (model.py:643-646)
        assert np.isfinite(r["state_itrf_km"][r["issued"]]).all() and np.isfinite(r["cov_rtn"][r["issued"]]).all()
        assert ((r["pc"][r["issued"]] >= 0) & (r["pc"][r["issued"]] <= 1)).all()
        assert (np.einsum("eoii->eo", r["cov_rtn"][:, 0, :, :3, :3]) > 0).all()      # positive position variances


def test_feature_tensor_matches_cdm_objects(setup):
    """events.feature_tensor (SoA, no CDM objects) == the same 64 LSTM input features read from the CDM objects
    with the reference's Event semantics for __CREATION_DATE / __TCA (event.py:42-47)."""
    ev_mod, model, batch, hits, util = setup["events"], setup["model"], setup["batch"], setup["hits"], setup["util"]
    res = ev_mod.generate_cdm_batch(model, batch, hits, seed=3)
    feats, ptr = ev_mod.feature_tensor(res)
    assert feats.shape == (int(res["issued"].sum()), 64) and len(ev_mod.LSTM_FEATURES) == 64 and ptr[-1] == len(feats)
    for e in range(len(hits)):
        cdms = ev_mod.to_cdms(model, res, e)
        assert len(cdms) == ptr[e + 1] - ptr[e]
        date0 = cdms[0]["CREATION_DATE"]
        for row, cdm in zip(feats[ptr[e]:ptr[e + 1]], cdms):
            want = [util.from_date_str_to_days(cdm["CREATION_DATE"], date0=date0), util.from_date_str_to_days(cdm["TCA"], date0=date0)]
            want += [cdm[k] for k in ev_mod.LSTM_FEATURES[2:]]
            assert np.allclose(row, np.array(want, dtype=np.float64), rtol=1e-12, atol=1e-9)


def test_generated_cdms_are_complete_and_survive_the_wire_format(setup, tmp_path):
    """Every CDM of the batched ground segment carries all obligatory CCSDS keys and reads back from its KVN file with
    the same numbers (to the 16 significant digits `str(float)` writes; exponents are written with 4 digits, as the
    reference does)."""
    from kessler_b200.cdm import ConjunctionDataMessage, Event
    ev_mod, model, batch, hits = setup["events"], setup["model"], setup["batch"], setup["hits"]
    res = ev_mod.generate_cdm_batch(model, batch, hits, seed=3)
    cdms = ev_mod.to_cdms(model, res, 0)
    assert len(cdms) >= 1
    paths = Event(cdms).save(str(tmp_path), prefix="event_7")
    assert len(paths) == len(cdms) and paths[0].endswith("event_7_0.kvn")
    for cdm, path in zip(cdms, paths):
        assert cdm.validate() == []
        back = ConjunctionDataMessage.load(path)
        assert back.validate() == []
        for key in ("TCA", "CREATION_DATE", "OBJECT1_OBJECT_NAME", "OBJECT2_OBJECT_DESIGNATOR", "COLLISION_PROBABILITY_METHOD"):
            assert str(back[key]) == str(cdm[key])
        for key in ("MISS_DISTANCE", "RELATIVE_SPEED", "OBJECT1_X", "OBJECT2_Z_DOT", "OBJECT1_CT_T", "OBJECT2_CNDOT_NDOT"):
            assert back[key] == pytest.approx(cdm[key], rel=1e-3)       # 1e-3: the %.3E form of tiny covariance terms
        assert back["OBJECT1_X"] == cdm["OBJECT1_X"] and back["MISS_DISTANCE"] == cdm["MISS_DISTANCE"]
    frame = Event(cdms).to_dataframe()
    assert frame.shape[0] == len(cdms) and "OBJECT2_CR_R" in frame.columns
    # ... and the frame built straight from the struct-of-arrays result holds the same rows for this event
    big = ev_mod.to_dataframe(res)
    assert len(big) == int(res["issued"].sum()) and set(big["event_id"]) == set(np.nonzero(res["issued"].any(axis=1))[0])
    mine = big[big["event_id"] == 0].reset_index(drop=True)
    assert len(mine) == len(cdms) and list(mine["TCA"]) == list(frame["TCA"]) and list(mine["CREATION_DATE"]) == list(frame["CREATION_DATE"])
    for key in ("MISS_DISTANCE", "RELATIVE_VELOCITY_T", "OBJECT1_Y", "OBJECT2_CTDOT_N", "COLLISION_PROBABILITY"):
        assert np.allclose(mine[key].to_numpy(dtype=np.float64), frame[key].to_numpy(dtype=np.float64), rtol=1e-12, atol=1e-9)


def test_forward_batch_device_returns_what_the_host_path_would_accept(setup):
    """The device-resident rejection step (draw, pre-filter, screen, compact on the GPU) hands back accepted draws only:
    re-screening the returned element sets through the host entry point gives the same indices and distances, every one is
    a conjunction by the reference's rule (i_conj > 0, no deep-space refusal), passes the altitude pre-filter, and its raw
    draws are the ones its SGP4 inputs were quantised from."""
    import math
    M, eng, model, tles = setup["M"], setup["engine"], setup["model"], setup["tles"]
    from kessler_b200.tle import MU_EARTH, R_EARTH
    torch.manual_seed(5)
    got = None
    for _ in range(6):
        c, nh = model.forward_batch_device(65536)
        if nh >= 3:
            got = c
            break
    assert got is not None and got["elems_t"].shape == (7, nh) and got["t_raw"]["_sites"] == ["mean_anomaly", "sampled_index"]
    times = model.time_grid()
    tc = np.ascontiguousarray(times[::100])
    r = eng.screen_host(got["elems_t"], got["epoch_t"], got["elems_c"], got["epoch_c"], tc, len(times), model._miss_dist_threshold)
    assert np.array_equal(r["i_min"], got["i_min"]) and np.array_equal(r["i_conj"], got["i_conj"])
    assert np.array_equal(r["d_min"], got["d_min"]) and np.array_equal(r["d_conj"], got["d_conj"])
    assert (got["i_conj"] > 0).all() and np.array_equal(got["time_min"], times[got["i_min"]])
    deg = math.pi / 180.0
    for who in ("t", "c"):
        raw, el = got[who + "_raw"], got["elems_" + who]
        idx = raw["sampled_index"]
        assert idx.dtype == np.int64 and idx.min() >= 1 and idx.max() < len(tles)
        base = np.stack([tles[int(i)].elements() for i in idx], axis=1)
        assert np.array_equal(el[[0, 1, 2, 3, 4, 6]], base[[0, 1, 2, 3, 4, 6]])            # the population's own elements ...
        assert np.array_equal(el[5], M._round_decimals(raw["mean_anomaly"] / deg, 4) * deg)  # ... and the quantised mean anomaly
        assert np.array_equal(got["epoch_" + who], np.array([tles[int(i)].date_mjd for i in idx]))
    a = lambda raw: (MU_EARTH / raw["mean_motion"] ** 2) ** (1.0 / 3.0)
    apo = lambda raw: a(raw) * (1 + raw["eccentricity"]) - R_EARTH
    per = lambda raw: a(raw) * (1 - raw["eccentricity"]) - R_EARTH
    margin = model._miss_dist_threshold + 1000
    pre = ((apo(got["t_raw"]) + margin) < per(got["c_raw"])) | ((apo(got["c_raw"]) + margin) < per(got["t_raw"]))
    assert not pre.any()
    # max_hits caps the number handed back; exclude_object_name is honoured per draw on the device too
    c2, n2 = model.forward_batch_device(65536, max_hits=1)
    assert n2 <= 1 and c2["elems_c"].shape[1] == n2
    for t, name in zip(tles, ["STARLINK-%d" % k if k in (3, 7, 11) else "OBJ-%d" % k for k in range(len(tles))]):
        t.name = name
    m2 = M.ConjunctionSimplified(tles=tles, exclude_object_name="STARLINK", time0=60727.13899462018)
    dev = torch.device("cuda", 0)
    dt = m2._draw_device("t", 20000, dev)
    dc = m2._draw_device("c", 20000, dev)
    ti, ci = dt["raw"]["sampled_index"].cpu().numpy(), dc["raw"]["sampled_index"].cpu().numpy()
    excl = np.isin(ti, (3, 7, 11))
    k_incl = len(m2._include_idxs)
    assert excl.any() and ci[excl].max() < k_incl and ci[~excl].max() == len(tles) - 1     # Categorical(tensor(include_idxs)): values < len(include_idxs)
