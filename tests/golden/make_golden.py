"""Regenerates the golden fixtures in this directory.  Runs ONLY in the build
container, where the reference checkout is mounted at /root/reference (it does
not exist on the GPU box; tests read the committed fixtures instead).

Fixture 1  trace_golden.json
    Numbers read out of /root/reference/docs/notebooks/trace.pickle -- a pickled
    Pyro trace of one ConjunctionSimplified.forward() run made by the reference
    authors.  pyro / dsgp4 / kessler are not importable
    here, so the pickle is opened with a stub unpickler that maps their classes
    to attribute bags.  Holds, per TLE: the 7 SGP4 inputs, every sgp4init
    coefficient, the mean elements of the last propagate() call and its tsince;
    plus time0 / delta_time / i_conj / time_conj / d_conj / time_min / d_min /
    max_ncdm / time_cdm_i.

Fixture 3  population.json
    The 20 TLEs of docs/notebooks/tles_sample_population.txt as text + parsed elements.

Fixture 2  reference_functions.npz
    Outputs of the reference's OWN functions that do not need dsgp4 arithmetic
    (kessler/util.py: upsample, find_closest, rotation_matrix,
    from_cartesian_to_rtn, from_rtn_to_cartesian, from_cartesian_to_keplerian,
    keplerian_cartesian_partials, TruncatedNormal.log_prob; kessler/model.py:
    find_conjunction) executed from /root/reference with dsgp4 / pyro /
    matplotlib / skyfield replaced by inert module stubs, on seeded inputs that
    are stored next to the outputs.
"""
import importlib
import json
import os
import pickle
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


class _Bag:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, st):
        if isinstance(st, dict):
            self.__dict__.update(st)
        else:
            self.__dict__["_state"] = st


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] in ("pyro", "dsgp4", "kessler"):
            return type(name, (_Bag,), {"__module__": module})
        if module.startswith("numpy.core"):
            module = module.replace("numpy.core", "numpy._core")
        return super().find_class(module, name)


def _f(x):
    if torch.is_tensor(x):
        x = x.item()
    if isinstance(x, (np.floating, np.integer)):
        x = x.item()
    return x


TLE_KEYS = ["_no_kozai", "_ecco", "_inclo", "_nodeo", "_argpo", "_mo", "_bstar", "_ndot", "_nddot",
            "_isimp", "_aycof", "_con41", "_cc1", "_cc4", "_cc5", "_d2", "_d3", "_d4", "_delmo", "_eta",
            "_argpdot", "_omgcof", "_sinmao", "_t", "_t2cof", "_t3cof", "_t4cof", "_t5cof", "_x1mth2",
            "_x7thm1", "_mdot", "_nodedot", "_xlcof", "_xmcof", "_nodecf", "_no_unkozai", "_a", "_alta",
            "_altp", "_gsto", "_am", "_em", "_im", "_Om", "_om", "_mm", "_nm", "_error",
            "_tumin", "_mu", "_radiusearthkm", "_xke", "_j2", "_j3", "_j4", "_j3oj2"]


def make_trace_golden():
    with open(os.path.join(REF, "docs/notebooks/trace.pickle"), "rb") as fh:
        tr = _Unpickler(fh).load()
    nodes = dict(tr.__dict__["nodes"])
    out = {"source": "docs/notebooks/trace.pickle", "sites": {}, "tles": {}}
    for k, v in nodes.items():
        val = v.get("value")
        if torch.is_tensor(val) and val.numel() == 1:
            out["sites"][k] = _f(val)
        elif isinstance(val, (float, int, np.floating)):
            out["sites"][k] = _f(val)
    for who in ("t_tle", "c_tle"):
        d = nodes[who]["infer"][who].__dict__
        rec = {k: _f(d[k]) for k in TLE_KEYS if k in d}
        data = d["_data"]
        rec["line1"], rec["line2"] = d["_lines"]
        rec["name"] = data.get("name")
        rec["epoch_year"] = int(data["epoch_year"])
        rec["epoch_days"] = float(data["epoch_days"])
        ep = data["_epoch"]
        rec["epoch_datetime"] = [ep.year, ep.month, ep.day, ep.hour, ep.minute, ep.second, ep.microsecond]
        rec["date_mjd"] = float(data["date_mjd"])
        rec["mean_motion"] = _f(data["mean_motion"])
        rec["semi_major_axis"] = _f(data["semi_major_axis"])
        rec["mean_motion_first_derivative"] = _f(data["mean_motion_first_derivative"])
        rec["mean_motion_second_derivative"] = _f(data["mean_motion_second_derivative"])
        for k in ("satellite_catalog_number", "classification", "international_designator",
                  "ephemeris_type", "element_number", "revolution_number_at_epoch"):
            rec[k] = data[k]
        out["tles"][who] = rec
    cdms = nodes["cdms"]["infer"]["cdms"]
    out["cdms"] = []
    for cdm in cdms:
        d = cdm.__dict__
        out["cdms"].append({
            "TCA": d["_values_relative_metadata"]["TCA"],
            "CREATION_DATE": d["_values_header"]["CREATION_DATE"],
            "MISS_DISTANCE": _f(d["_values_relative_metadata"]["MISS_DISTANCE"]),
            "COLLISION_PROBABILITY": _f(d["_values_relative_metadata"]["COLLISION_PROBABILITY"]),
            "state": [[_f(o[k]) for k in ("X", "Y", "Z", "X_DOT", "Y_DOT", "Z_DOT")]
                      for o in d["_values_object_data_state"]],
            "cov_rtn_diag": [[_f(o[k]) for k in ("CR_R", "CT_T", "CN_N", "CRDOT_RDOT", "CTDOT_TDOT", "CNDOT_NDOT")]
                             for o in d["_values_object_data_covariance"]],
        })
    with open(os.path.join(HERE, "trace_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    return out


# ---------------------------------------------------------------------------
def _install_stubs():
    """Inert stand-ins so `import kessler.util` / `kessler.model` run from the
    reference tree without its absent third-party dependencies."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Any:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Any()

        def __getattr__(self, k):
            return _Any()

    dutil = mod("dsgp4.util", clone_w_grad=lambda y: y.clone().detach().requires_grad_(True))
    dtle = mod("dsgp4.tle", TLE=_Any)
    mod("dsgp4", util=dutil, tle=dtle, TLE=_Any)
    mod("matplotlib", pyplot=None, use=lambda *a, **k: None)
    for sub in ("pyplot", "dates", "ticker", "colors", "cm", "patches", "lines", "gridspec"):
        sys.modules["matplotlib." + sub] = mod("matplotlib." + sub)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    import torch.distributions as td

    class TorchDistribution(td.Distribution):
        pass

    pdist = mod("pyro.distributions", MixtureSameFamily=td.MixtureSameFamily, Categorical=td.Categorical,
                Uniform=td.Uniform, Normal=td.Normal, Bernoulli=td.Bernoulli, Delta=_Any)
    ptd = mod("pyro.distributions.torch_distribution", TorchDistribution=TorchDistribution)
    pdist.torch_distribution = ptd
    mod("pyro", distributions=pdist, sample=_Any(), deterministic=_Any(), poutine=_Any())
    mod("skyfield")
    mod("skyfield.sgp4lib")


def make_reference_functions():
    _install_stubs()
    # import the two modules straight from their files (kessler/__init__ pulls
    # in cdm/event/nn/plot, which are out of scope and need more stubs)
    pkg = types.ModuleType("kessler")
    pkg.__path__ = [os.path.join(REF, "kessler")]
    sys.modules["kessler"] = pkg
    om = importlib.import_module("kessler.observation_model")
    pkg.GNSS, pkg.Radar = om.GNSS, om.Radar
    pkg.ConjunctionDataMessage = object
    util = importlib.import_module("kessler.util")
    pkg.util = util
    model = importlib.import_module("kessler.model")

    rng = np.random.default_rng(20261019)
    out = {}

    # upsample (util.py:541-555): ragged sizes incl. the production 6000 -> 600000
    for tag, (n_c, n_f) in {"a": (6000, 600000), "b": (7, 61), "c": (2, 9), "d": (100, 9901), "e": (61, 6001)}.items():
        x = rng.standard_normal((n_c, 6)) * 7000.0
        y = util.upsample(torch.from_numpy(x), n_f).numpy()
        out[f"upsample_{tag}_in"] = x
        if n_f > 20000:  # keep the fixture small: store a strided subset + checksums
            idx = np.unique(np.concatenate([np.arange(0, n_f, 997), np.arange(0, 300), np.arange(n_f - 300, n_f),
                                            rng.integers(0, n_f, 2000)]))
            out[f"upsample_{tag}_idx"] = idx
            out[f"upsample_{tag}_out"] = y[idx]
            out[f"upsample_{tag}_sum"] = y.sum(axis=0)
        else:
            out[f"upsample_{tag}_out"] = y
        out[f"upsample_{tag}_nf"] = np.array(n_f)

    # find_conjunction (model.py:23-52)
    fc_in0, fc_in1, fc_res = [], [], []
    for case in range(6):
        n = 1500
        tt = np.linspace(0, 1, n)[:, None]
        a = rng.standard_normal((1, 3)) * 7e6 + rng.standard_normal((1, 3)) * 1e5 * tt
        b = a + (rng.standard_normal((1, 3)) * 4e4) * (tt - rng.uniform(0.1, 0.9)) + rng.standard_normal((1, 3)) * (2e3 if case % 2 else 2e4)
        if case == 4:  # minimum / first-below at index 0
            b[0] = a[0] + 1.0
        if case == 5:  # exact tie -> first index wins
            b[100] = a[100] + np.array([3.0, 4.0, 0.0])
            b[200] = a[200] + np.array([3.0, 4.0, 0.0])
            b[[i for i in range(n) if i not in (100, 200)]] += 1e6
        i_min, d_min, i_conj, d_conj = model.find_conjunction(torch.from_numpy(a), torch.from_numpy(b), 5e3)
        fc_in0.append(a)
        fc_in1.append(b)
        fc_res.append([i_min, float(d_min), -1 if i_conj is None else i_conj, np.nan if d_conj is None else float(d_conj)])
    out["fc_tr0"], out["fc_tr1"], out["fc_res"] = np.stack(fc_in0), np.stack(fc_in1), np.array(fc_res)

    # find_closest (util.py:527-539)
    times = np.arange(58991.90384230018, 58991.90384230018 + 7.0, 7.0 / 6e5)
    q = 58991.90384230018 + np.arange(21) * 8.0 / 24.0
    out["find_closest_q"] = q
    out["find_closest_res"] = np.array([[util.find_closest(times, v)[0], util.find_closest(times, v)[1]] for v in q])
    out["times_len"] = np.array(len(times))
    out["times_probe_idx"] = np.array([0, 1, 100, 305767, 305768, 599900, 599999])
    out["times_probe"] = times[out["times_probe_idx"]]

    # RTN / Kepler conversions (util.py:187-358)
    states = []
    for _ in range(16):
        r = rng.standard_normal(3)
        r = r / np.linalg.norm(r) * rng.uniform(6.6e6, 8.5e6)
        v = rng.standard_normal(3)
        v -= v.dot(r) / r.dot(r) * r * rng.uniform(0.9, 1.0)
        v = v / np.linalg.norm(v) * np.sqrt(398600.5e9 / np.linalg.norm(r)) * rng.uniform(0.97, 1.08)
        states.append(np.stack([r, v]))
    states = np.stack(states)
    out["conv_states"] = states
    out["conv_rotation"] = np.stack([util.rotation_matrix(s) for s in states])
    out["conv_rtn"] = np.stack([util.from_cartesian_to_rtn(s)[0] for s in states])
    out["conv_back"] = np.stack([util.from_rtn_to_cartesian(util.from_cartesian_to_rtn(s)[0], util.rotation_matrix(s).T) for s in states])
    # util.py:215 builds K = torch.tensor([0.,0.,1.]) in the default dtype; with FP64
    # states torch 2.11 refuses the mixed-dtype cross product the reference's older
    # torch promoted, so run these two calls with the default dtype set to FP64.
    torch.set_default_dtype(torch.float64)
    kep, jac = [], []
    for s in states:
        ts = torch.tensor(s)
        kep.append([float(x) for x in util.from_cartesian_to_keplerian(ts[0], ts[1], 398600.5e9)])
        jac.append(util.keplerian_cartesian_partials(torch.tensor(s).requires_grad_(), 398600.5e9))
    torch.set_default_dtype(torch.float32)
    out["conv_kepler"] = np.array(kep)
    out["conv_partials"] = np.stack([np.asarray(j, dtype=np.float64) for j in jac])

    # TruncatedNormal.log_prob (util.py:46-160) -- the reference's own test values too
    tn = util.TruncatedNormal(loc=torch.tensor([0.0010028142482042313, 0.00017592836171388626]),
                              scale=torch.tensor([0.00004670943133533001, 0.00003172305878251791]), low=0.0, high=0.004)
    xs = torch.tensor([[0.0009, 0.0002], [0.00105, 0.00015], [0.002, 0.0001]])
    out["tn_x"] = xs.numpy()
    out["tn_logp"] = tn.log_prob(xs).numpy()

    np.savez_compressed(os.path.join(HERE, "reference_functions.npz"), **out)
    return out


def make_population():
    """docs/notebooks/tles_sample_population.txt -> population.json (parsed numbers only)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import sgp4_ref as S
    lines = open(os.path.join(REF, "docs/notebooks/tles_sample_population.txt")).read().splitlines()
    pop = []
    for i in range(0, len(lines), 3):
        p = S.parse_tle_lines(lines[i + 1], lines[i + 2])
        p.pop("epoch_dt")
        p["name"] = lines[i][2:].strip()
        p["line1"], p["line2"] = lines[i + 1], lines[i + 2]
        pop.append(p)
    with open(os.path.join(HERE, "population.json"), "w") as fh:
        json.dump(pop, fh, indent=1, sort_keys=True)
    return pop


if __name__ == "__main__":
    print("population.json:", len(make_population()), "TLEs")
    g = make_trace_golden()
    print("trace_golden.json:", len(g["sites"]), "sites,", len(g["tles"]), "tles,", len(g["cdms"]), "cdms")
    o = make_reference_functions()
    print("reference_functions.npz:", len(o), "arrays")
