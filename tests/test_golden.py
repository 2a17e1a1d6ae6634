"""Golden vectors (tests/golden/hmm_golden.json, made by tests/golden/make_golden.py).
CPU: the oracle still reproduces them bit for bit (freeze).  GPU: the CUDA path matches them within the gates."""
import hashlib
import json
import os

import numpy as np
import pytest

from helpers import ergodic_hmm, order2_hmm, silent_hmm, sparse_hmm
from jstacs_b200 import DNA, DataSet, DNAAlphabet

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "hmm_golden.json")))
CASES = {
    "erg_s2_k1": lambda: ergodic_hmm(2, 1),
    "erg_s8_k2": lambda: ergodic_hmm(8, 2),
    "erg_s16_k1_map": lambda: ergodic_hmm(16, 1, ess=4.0),
    "erg_s32_k3_map": lambda: ergodic_hmm(32, 3, ess=4.0),
    "sparse": lambda: sparse_hmm(),
    "order2_s3": lambda: order2_hmm(3, 1),
    "silent_profile": lambda: silent_hmm(),
}


def golden_data():
    lens = np.array(GOLD["lengths"], np.int64)
    off = np.zeros(lens.size + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    return DataSet.from_packed(DNAAlphabet.encode(GOLD["codes"]), off, DNA)


def unhex(v):
    return np.array([float.fromhex(x) for x in v])


def check(got, gold, exact, rtol=1e-9, atol=1e-9):
    got = np.ascontiguousarray(got, np.float64).reshape(-1)
    if isinstance(gold, dict):
        assert got.size == gold["n"]
        if exact:
            assert hashlib.sha256(got.tobytes()).hexdigest() == gold["sha256"]
        else:
            ref_h, ref_s = unhex(gold["head"]), unhex(gold["every97"])
            for g, r in ((got[:24], ref_h), (got[::97], ref_s)):
                fin = np.isfinite(r)
                assert np.array_equal(np.isfinite(g), fin)
                np.testing.assert_allclose(g[fin], r[fin], rtol=rtol, atol=atol)
            assert got[np.isfinite(got)].sum() == pytest.approx(float.fromhex(gold["sum"]), rel=1e-9)
    else:
        ref = unhex(gold)
        if exact:
            assert np.array_equal(got, ref) or (np.array_equal(np.isnan(got), np.isnan(ref)) and np.array_equal(got[~np.isnan(got)], ref[~np.isnan(ref)]))
        else:
            fin = np.isfinite(ref)
            assert np.array_equal(np.isfinite(got), fin)
            np.testing.assert_allclose(got[fin], ref[fin], rtol=rtol, atol=atol)


def test_java_random_stream_matches_fixture():
    from jstacs_b200.workloads import make_data, ragged_lengths

    lens = ragged_lengths(12, seed=42, lo=8.0, decades=1.0)
    assert [int(v) for v in lens] == GOLD["lengths"]
    data = make_data(12, lens)
    assert "".join("ACGT"[c] for c in data.codes) == GOLD["codes"]


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_golden(oracle_lib, name):
    g = GOLD["cases"][name]
    data = golden_data()
    om = oracle_lib.OracleModel(CASES[name]().ir())
    check(om.loglik(data.codes, data.offsets), g["loglik"], True)
    paths, scores, _ = om.viterbi(data.codes, data.offsets)
    assert [[int(s) for s in p] for p in paths] == g["viterbi_paths"]
    check(scores, g["viterbi_scores"], True)
    ts, es, sc = om.estep(data.codes, data.offsets)
    check(ts, g["trans_stat"], True)
    check(es, g["emis_stat"], True)
    assert sc == float.fromhex(g["score"])
    check(om.train(data.codes, data.offsets, max_iter=4), g["objective_4_iterations"], True)
    tp, tl, ep, el = om.get_params()
    check(tp, g["trans_param_after"], True)
    check(ep, g["emis_param_after"], True)


@pytest.mark.gpu
@pytest.mark.parametrize("fast", [0, 1])
@pytest.mark.parametrize("name", [n for n in CASES if n != "silent_profile"])
def test_cuda_matches_golden(name, fast):
    from jstacs_b200 import _native
    from jstacs_b200.hmm import BaumWelchParameterSet, IterationCondition

    _native.check(_native.lib().jh_set_option(b"fast", fast))
    try:
        g = GOLD["cases"][name]
        data = golden_data()
        hmm = CASES[name]()
        check(hmm.getLogScoreFor(data), g["loglik"], False, rtol=1e-9, atol=0)
        res = hmm.getViterbiPathsFor(data)
        assert [[int(s) for s in p] for p, _ in res] == g["viterbi_paths"]  # bit-exact paths
        check([s for _, s in res], g["viterbi_scores"], True)               # and scores
        ts, es, sc = hmm._engine().estep(hmm._data(data))
        check(ts, g["trans_stat"], False)
        check(es, g["emis_stat"], False)
        assert sc == pytest.approx(float.fromhex(g["score"]), rel=1e-9)
        hmm.trainingParameter = BaumWelchParameterSet(1, IterationCondition(4), 1)
        hmm.train(data)
        check(hmm.last_objectives[0], g["objective_4_iterations"], False, rtol=1e-9, atol=0)
        tp, tl, ep, el = hmm._engine().get_params()
        check(tp, g["trans_param_after"], False, rtol=0, atol=1e-6)
        check(ep, g["emis_param_after"], False, rtol=0, atol=1e-6)
    finally:
        _native.check(_native.lib().jh_set_option(b"fast", 1))
