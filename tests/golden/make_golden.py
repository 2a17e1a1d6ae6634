"""Generates tests/golden/hmm_golden.json from the CPU oracle (oracle/hmm_oracle.c).

The reference (Jstacs, Java) cannot run in the build container (no JVM) and ships no golden vectors for the HMM path
(SURVEY §4, §8c), so these vectors freeze the ORACLE's behaviour ("parity unpinned"); java/Harness.java recomputes the
same cases with Jstacs itself on any box with a JDK and diffs against this file.

Doubles are stored as C99 hex strings (bit-exact).  Usage:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import ergodic_hmm, order2_hmm, silent_hmm, sparse_hmm  # noqa: E402
from jstacs_b200.workloads import make_data, ragged_lengths  # noqa: E402
from oracle import binding  # noqa: E402

CASES = {
    "erg_s2_k1": lambda: ergodic_hmm(2, 1),                 # BASELINE cfg1 model
    "erg_s8_k2": lambda: ergodic_hmm(8, 2),                 # BASELINE cfg2 model
    "erg_s16_k1_map": lambda: ergodic_hmm(16, 1, ess=4.0),  # BASELINE cfg4 model, MAP prior
    "erg_s32_k3_map": lambda: ergodic_hmm(32, 3, ess=4.0),  # BASELINE cfg3 model
    "sparse": lambda: sparse_hmm(),
    "order2_s3": lambda: order2_hmm(3, 1),
    "silent_profile": lambda: silent_hmm(),
}


def hx(a):
    """Short arrays verbatim; long ones as {n, sum, head, stride sample, sha256} to keep the fixture small."""
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    if a.size <= 300:
        return [float(v).hex() for v in a]
    import hashlib

    return {"n": int(a.size), "sum": float(a[np.isfinite(a)].sum()).hex(), "head": [float(v).hex() for v in a[:24]],
            "every97": [float(v).hex() for v in a[::97]], "sha256": hashlib.sha256(a.tobytes()).hexdigest()}


def main():
    out = {"about": "oracle-generated golden vectors; data = java.util.Random(0x5EED0000+n).nextInt(4), lengths = ragged_lengths(12, seed=42, lo=8, decades=1)",
           "cases": {}}
    lens = ragged_lengths(12, seed=42, lo=8.0, decades=1.0)
    data = make_data(12, lens)
    out["lengths"] = [int(v) for v in lens]
    out["codes"] = "".join("ACGT"[c] for c in data.codes)
    for name, make in CASES.items():
        hmm = make()
        om = binding.OracleModel(hmm.ir())
        c = {}
        c["loglik"] = hx(om.loglik(data.codes, data.offsets))
        paths, scores, plen = om.viterbi(data.codes, data.offsets)
        c["viterbi_paths"] = [[int(s) for s in p] for p in paths]
        c["viterbi_scores"] = hx(scores)
        ts, es, sc = om.estep(data.codes, data.offsets)
        c["trans_stat"], c["emis_stat"], c["score"] = hx(ts), hx(es), float(sc).hex()
        obj = om.train(data.codes, data.offsets, max_iter=4)
        c["objective_4_iterations"] = hx(obj)
        tp, tl, ep, el = om.get_params()
        c["trans_param_after"], c["emis_param_after"] = hx(tp), hx(ep)
        out["cases"][name] = c
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hmm_golden.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=0, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
