"""Generates tests/golden/oracle_golden.json from the CPU oracle (the reference itself is Julia and
cannot run in this environment; the reference-held constants -- HMM log-ML, Unfold identities -- are
asserted directly in tests/test_oracle.py). Run: python -m tests.golden.make_golden"""
import json
import os

import numpy as np

from oracle import closed_forms as cf
from oracle import oracle as O

SPECS = [
    {"family": O.LGSSM, "params": [0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 1.0], "N": 64, "T": 12, "seed": 1, "prop": 0, "scheme": 0, "sim": "lgssm"},
    {"family": O.LGSSM, "params": [0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 1.0], "N": 50, "T": 8, "seed": 2, "prop": 1, "scheme": 1, "sim": "lgssm"},
    {"family": O.SV, "params": [-1.0, 0.97, 0.2], "N": 48, "T": 10, "seed": 3, "prop": 0, "scheme": 1, "sim": "sv"},
    {"family": O.BEARINGS, "params": list(cf.BEARINGS_PARAMS), "N": 40, "T": 6, "seed": 4, "prop": 1, "scheme": 0, "sim": "bearings"},
    {"family": O.HMM, "params": list(cf.hmm_params()), "N": 32, "T": 4, "seed": 5, "prop": 1, "scheme": 0, "sim": "hmm"},
]


def observations(spec):
    if spec["sim"] == "lgssm":
        return cf.simulate_lgssm(spec["T"], spec["params"], 0)
    if spec["sim"] == "sv":
        return cf.simulate_sv(spec["T"], spec["params"], 0)
    if spec["sim"] == "bearings":
        return cf.simulate_bearings(spec["T"])
    return np.array(cf.HMM_OBS, dtype=float)


def run_case(orc, spec):
    ys = observations(spec)
    N = spec["N"]
    pf = orc.particle_filter(spec["family"], spec["params"], N, seed=spec["seed"])
    pf.init([ys[0]], proposal=spec["prop"])
    anc_all, n_res = [], 0
    for t in range(1, spec["T"]):
        if pf.maybe_resample(N * 0.75, scheme=spec["scheme"]):
            n_res += 1
            anc_all = [int(a) for a in pf.parents()]
        pf.step([ys[t]], proposal=spec["prop"])
    lw = pf.log_weights()
    return {"log_ml": pf.log_ml_estimate(), "n_resamples": n_res, "last_ancestors": anc_all,
            "lw_first": float(lw[0]), "lw_sum": float(np.sum(lw)), "state_sum": float(np.sum(pf.state()))}


def main():
    orc = O.Oracle()
    cases = [{"spec": s, "expect": run_case(orc, s)} for s in SPECS]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.json")
    with open(path, "w") as fh:
        json.dump({"generator": "tests/golden/make_golden.py", "cases": cases}, fh, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
