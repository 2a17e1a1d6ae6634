"""The route a Jstacs user takes — the Java drop-in's call sequences — replayed by tests/harness/shim_replay.c against the
C ABI and checked against the oracle.  (No JDK in the image: the C program is the executable stand-in of
java/de/jstacs/.../models/GpuHmmEngine.java and follows its class comment call by call.)"""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

from helpers import dataset_of, ergodic_hmm, order0_hmm, order2_hmm, random_dna, silent_hmm, sparse_hmm
from jstacs_b200.workloads import banded_hmm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "harness", "shim_replay.c")
DT = {np.dtype(np.int32): 0, np.dtype(np.int64): 1, np.dtype(np.float64): 2, np.dtype(np.uint8): 3}
RDT = {0: np.int32, 1: np.int64, 2: np.float64, 3: np.uint8}


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("harness") / "shim_replay")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-o", exe, SRC, "-ldl"])
    return exe


def write_records(path, recs):
    with open(path, "wb") as f:
        for name, arr in recs.items():
            arr = np.ascontiguousarray(arr)
            b = name.encode()
            f.write(struct.pack("<I", len(b)) + b + struct.pack("<BQ", DT[arr.dtype], arr.size) + arr.tobytes())


def read_records(path):
    out = {}
    raw = open(path, "rb").read()
    p = 0
    while p < len(raw):
        (nl,) = struct.unpack_from("<I", raw, p)
        name = raw[p + 4:p + 4 + nl].decode()
        dt, cnt = struct.unpack_from("<BQ", raw, p + 4 + nl)
        p += 4 + nl + 9
        dtype = np.dtype(RDT[dt])
        out[name] = np.frombuffer(raw, dtype=dtype, count=cnt, offset=p).copy()
        p += cnt * dtype.itemsize
    return out


def model_records(ir):
    r = {k: np.array([getattr(ir, k)], np.int32) for k in ("n_states", "alphabet", "max_order", "n_elem", "n_emis")}
    for k in ir.FIELDS_I32:
        r[k] = np.asarray(getattr(ir, k), np.int32)
    for k in ir.FIELDS_F64:
        r[k] = np.asarray(getattr(ir, k), np.float64)
    for k in ir.FIELDS_U8:
        r[k] = np.asarray(getattr(ir, k), np.uint8)
    return r


def run(harness, tmp_path, recs, ops, expect_rc=0):
    from jstacs_b200 import _native

    inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    write_records(inp, recs)
    p = subprocess.run([harness, _native.LIB_PATH, inp, outp] + list(ops), capture_output=True, text=True, timeout=600)
    assert p.returncode == expect_rc, f"rc {p.returncode}\n{p.stderr}"
    return read_records(outp), p.stderr


def ragged(seed, n, lo=1, hi=90):
    rng = np.random.default_rng(seed)
    return dataset_of([random_dna(rng, int(L)) for L in rng.integers(lo, hi, n)])


def test_harness_builds_and_fails_loudly_without_a_device(harness, tmp_path):
    """Through the shim's route, too, there is no CPU fallback: without a device the first call reports JH_ERR_CUDA."""
    import __graft_entry__ as g

    g.build()
    from jstacs_b200 import _native

    if _native.lib().jh_device_count() > 0:
        pytest.skip("a CUDA device is present")
    hmm = ergodic_hmm(3, 1)
    data = ragged(1, 4)
    recs = model_records(hmm.ir())
    recs.update(codes=data.codes, offsets=data.offsets)
    out, err = run(harness, tmp_path, recs, ["score"], expect_rc=3)
    assert out["error_status"][0] == -3 and "no CUDA device" in err


SHIM_MODELS = {
    "erg_s2_k1": lambda: ergodic_hmm(2, 1),
    "erg_s8_k2": lambda: ergodic_hmm(8, 2, ess=2.0),
    "erg_s32_k3": lambda: ergodic_hmm(32, 3, ess=4.0),
    "sparse_unsorted": lambda: sparse_hmm(),
    "order2_s3": lambda: order2_hmm(3, 1),
    "order0": lambda: order0_hmm(ess=2.0),
    "silent_profile": lambda: silent_hmm(),
    "banded_s40": lambda: banded_hmm(40, 4, 1),
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(SHIM_MODELS))
def test_shim_call_sequence_matches_oracle(harness, tmp_path, oracle_lib, name):
    """score -> viterbi -> decode -> posterior -> windows -> train (weights) -> score again with the trained parameters ->
    a second DataSet: everything the shim's public methods do, on cached handles, against the oracle."""
    hmm = SHIM_MODELS[name]()
    ir = hmm.ir()
    om = oracle_lib.OracleModel(ir)
    data = ragged(11, 24, lo=2)
    rng = np.random.default_rng(5)
    w = rng.random(data.getNumberOfElements()) + 0.5
    lens = np.diff(data.offsets)
    win_seq = np.array([i for i in range(len(lens)) if lens[i] >= 6][:5], np.int64)
    win_start = np.array([1, 0, 2, 3, 1][:win_seq.size], np.int64)
    win_end = np.array([int(lens[s]) - 2 for s in win_seq], np.int64)
    recs = model_records(ir)
    recs.update(codes=data.codes, offsets=data.offsets, weights=w, win_seq=win_seq, win_start=win_start, win_end=win_end)
    out, _ = run(harness, tmp_path, recs, ["score", "viterbi", "decode", "posterior", "windows", "train:1:0:5:0:1", "score", "viterbi"])
    assert out["done"][0] == 1
    # --- before training (records are overwritten by the later ops of the same name: read_records keeps the LAST one) ---
    first, _ = run(harness, tmp_path, recs, ["score", "viterbi", "decode", "posterior", "windows"])
    np.testing.assert_allclose(first["loglik"], om.loglik(data.codes, data.offsets), rtol=1e-9)
    rp, rs, _ = om.viterbi(data.codes, data.offsets)
    ptr = first["viterbi_ptr"]
    for i in range(len(lens)):
        assert np.array_equal(first["viterbi_paths"][ptr[i]:ptr[i + 1]], rp[i]), f"Viterbi path {i}"
    assert np.array_equal(first["viterbi_scores"], rs)
    for i in range(len(lens)):
        ref = om.posterior(data.getElementAt(i).codes)
        got = first[f"posterior_{i}"].reshape(ref.shape)
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(got), fin)
        np.testing.assert_allclose(got[fin], ref[fin], rtol=0, atol=1e-10)
        srt = np.sort(ref, axis=0)
        clear = srt[-1] - srt[-2] > 1e-9
        dec = first["decoded_paths"][data.offsets[i]:data.offsets[i + 1]]
        assert np.array_equal(dec[clear], om.decode_posterior(ref)[clear])
    for k in range(win_seq.size):
        x = data.getElementAt(int(win_seq[k])).codes
        assert first["window_loglik"][k] == pytest.approx(om.loglik_window(x, int(win_start[k]), int(win_end[k])), rel=1e-9)
    # --- training with weights, then scoring with the parameters written back ---
    robj = om.train(data.codes, data.offsets, w, max_iter=5)
    np.testing.assert_allclose(out["train_objective_start0"], robj, rtol=1e-9)
    assert out["train_last_start0"][0] == out["train_objective_start0"][-1] == out["train_best"][0]
    rtp, rtl, rep, rel_ = om.get_params()
    for got, ref in ((out["trained_trans_param"], rtp), (out["trained_trans_lognorm"], rtl), (out["trained_emis_param"], rep),
                     (out["trained_emis_lognorm"], rel_)):
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(got), fin)
        np.testing.assert_allclose(got[fin], ref[fin], rtol=0, atol=1e-6)
    np.testing.assert_allclose(out["loglik"], om.loglik(data.codes, data.offsets), rtol=1e-9)  # trained parameters, both sides
    rp2, rs2, _ = om.viterbi(data.codes, data.offsets)
    ptr = out["viterbi_ptr"]
    # trained parameters agree to ~1e-12, not bit for bit: paths are compared where the oracle's own decision is stable,
    # i.e. through the scores
    np.testing.assert_allclose(out["viterbi_scores"], rs2, rtol=1e-9)
    same = sum(np.array_equal(out["viterbi_paths"][ptr[i]:ptr[i + 1]], rp2[i]) for i in range(len(lens)))
    assert same >= 0.8 * len(lens)  # near-ties can flip with parameters that differ in the last bits; the scores above are the gate


@pytest.mark.gpu
def test_shim_multi_start_keeps_the_best_start(harness, tmp_path, oracle_lib):
    """numberOfStarts > 1 (HigherOrderHMM.java:741-780): per start the initial parameters are pushed, the best start's
    parameters end up in the model; the last objective is compared, not the capped history."""
    hmm = ergodic_hmm(4, 1, ess=1.0)
    ir = hmm.ir()
    data = ragged(3, 40, lo=10, hi=60)
    recs = model_records(ir)
    recs.update(codes=data.codes, offsets=data.offsets)
    rng = np.random.default_rng(0)
    refs = []
    for k in range(3):
        h2 = ergodic_hmm(4, 1, ess=1.0)
        for t in h2.transition.transitions:
            t.setLogProbabilities(np.log(rng.dirichlet(np.ones(len(t.states)))))
        for e in h2.emission:
            e.setLogProbabilities(np.log(rng.dirichlet(np.ones(4), size=e.params.shape[0])))
        ir2 = h2.ir()
        recs[f"start{k}_trans_param"], recs[f"start{k}_trans_lognorm"] = ir2.trans_param, ir2.trans_lognorm
        recs[f"start{k}_emis_param"], recs[f"start{k}_emis_lognorm"] = ir2.emis_param, ir2.emis_lognorm
        om = oracle_lib.OracleModel(ir2)
        refs.append((om.train(data.codes, data.offsets, max_iter=4), om.get_params()))
    out, _ = run(harness, tmp_path, recs, ["train:1:0:4:0:3", "score"])
    finals = [out[f"train_last_start{k}"][0] for k in range(3)]
    for k in range(3):
        np.testing.assert_allclose(out[f"train_objective_start{k}"], refs[k][0], rtol=1e-9)
    best = int(np.argmax(finals))
    assert best == int(np.argmax([r[0][-1] for r in refs])) and out["train_best"][0] == finals[best]
    np.testing.assert_allclose(out["trained_trans_param"], refs[best][1][0], rtol=0, atol=1e-6)
    om = oracle_lib.OracleModel(ir)
    om.set_params(*refs[best][1])
    np.testing.assert_allclose(out["loglik"], om.loglik(data.codes, data.offsets), rtol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["erg_s8_k2", "sparse_unsorted", "order2_s3"])
def test_shim_semi_supervised_training_with_states_groups(harness, tmp_path, oracle_lib, name):
    """A model built with statesGroups and a sequence annotation type (HigherOrderHMM.java:183-185): model() passes the
    groups, train() attaches the per-position AllowedStatesGroups of the annotations (doOneStep :810-820), the scoring call
    afterwards is unconstrained again — objectives, trained parameters and the final scores against the oracle."""
    hmm = SHIM_MODELS[name]()
    S = hmm.getNumberOfStates()
    groups = [list(range(0, S, 2)), list(range(1, S, 2)), [0, S - 1]]
    ir = hmm.ir()
    data = ragged(21, 20, lo=5, hi=70)
    rng = np.random.default_rng(2)
    grp = rng.integers(0, len(groups), data.codes.size).astype(np.int32)
    grp[rng.random(grp.size) < 0.7] = -1
    lens = np.diff(data.offsets)
    for i in range(0, len(lens), 3):  # every third sequence carries no annotation at all
        grp[data.offsets[i]:data.offsets[i + 1]] = -1
    om = oracle_lib.OracleModel(ir)
    om.set_groups(groups)
    om.attach_groups(data.codes, grp)
    fin = np.isfinite(om.loglik(data.codes, data.offsets))
    for i in np.flatnonzero(~fin):  # sequences that are impossible under their labels would stop the reference, too
        grp[data.offsets[i]:data.offsets[i + 1]] = -1
    om.attach_groups(data.codes, grp)
    w = rng.random(len(lens)) + 0.5
    recs = model_records(ir)
    ptr = np.zeros(len(groups) + 1, np.int32)
    ptr[1:] = np.cumsum([len(g) for g in groups])
    recs.update(codes=data.codes, offsets=data.offsets, weights=w, group_ptr=ptr,
                group_states=np.array([s for g in groups for s in g], np.int32), seq_groups=grp)
    out, _ = run(harness, tmp_path, recs, ["train:1:0:4:0:1", "score"])
    robj = om.train(data.codes, data.offsets, w, max_iter=4)
    np.testing.assert_allclose(out["train_objective_start0"], robj, rtol=1e-9)
    rtp, rtl, rep, rel_ = om.get_params()
    for got, ref in ((out["trained_trans_param"], rtp), (out["trained_emis_param"], rep)):
        f = np.isfinite(ref)
        assert np.array_equal(np.isfinite(got), f)
        np.testing.assert_allclose(got[f], ref[f], rtol=0, atol=1e-6)
    om.attach_groups(None, None)  # the score call is unconstrained: the shim detached the groups
    np.testing.assert_allclose(out["loglik"], om.loglik(data.codes, data.offsets), rtol=1e-9)
    # the constraint was really evaluated: training without it gives another objective
    om2 = oracle_lib.OracleModel(ir)
    assert abs(om2.train(data.codes, data.offsets, w, max_iter=4)[-1] - robj[-1]) > 1e-6 * abs(robj[-1])


@pytest.mark.gpu
def test_shim_impossible_path_is_an_error(harness, tmp_path):
    """A sequence without a state path of probability > 0: the reference throws IllegalArgumentException("Impossible path")
    (HigherOrderHMM.java:659); through the shim's route the library reports JH_ERR_IMPOSSIBLE."""
    hmm = sparse_hmm()
    e = hmm.emission[0]
    lp = e.params.copy()
    lp[:, 0] = -np.inf
    lp = lp - np.log(np.exp(lp).sum(1, keepdims=True))
    e.setLogProbabilities(lp)
    for em in hmm.emission[1:]:
        q = em.params.copy()
        q[:, 0] = -np.inf
        em.setLogProbabilities(q - np.log(np.exp(q).sum(1, keepdims=True)))
    data = dataset_of([[1, 2, 3, 1], [0, 0, 0], [2, 2]])
    recs = model_records(hmm.ir())
    recs.update(codes=data.codes, offsets=data.offsets)
    out, err = run(harness, tmp_path, recs, ["viterbi"], expect_rc=3)
    assert out["error_status"][0] == -6 and "Impossible path" in err


if __name__ == "__main__":
    sys.exit(pytest.main([__file__, "-q"]))
