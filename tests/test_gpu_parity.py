"""Parity of the CUDA path (through the C ABI) against the oracle on identical inputs.

Gates (BASELINE.json north_star): Viterbi paths bit-exact; per-sequence log-likelihood <= 1e-9 relative (fp64);
trained parameters after N EM iterations <= 1e-6; expected counts checked at 1e-9.
Every test runs the reference-order kernels (fast=0) and, where the model allows, the scaled fast kernels (fast=1).
"""
import numpy as np
import pytest

from helpers import sunflower_hmm, dataset_of, ergodic_hmm, order0_hmm, order2_hmm, random_dna, silent_hmm, sparse_hmm
from jstacs_b200 import (DNA, BaumWelchParameterSet, DataSet, DiscreteEmission, HigherOrderHMM, IterationCondition,
                         Sequence, SmallDifferenceOfFunctionEvaluationsCondition, TransitionElement, ViterbiParameterSet,
                         WrongAlphabetException)
from jstacs_b200.workloads import banded_hmm, make_data, planted_dna, ragged_lengths

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-9      # north_star: per-sequence log-likelihood within 1e-9 relative
COUNT_TOL = 1e-9    # expected counts (absolute on O(1..L) values, relative otherwise)
PARAM_TOL = 1e-6    # north_star: trained parameters within 1e-6


@pytest.fixture(scope="module")
def nat():
    import __graft_entry__ as g

    g.build()
    from jstacs_b200 import _native

    if _native.lib().jh_device_count() <= 0:
        pytest.fail("no CUDA device: the product has no CPU fallback")
    return _native


@pytest.fixture(params=[0, 1], ids=["exact", "fast"])
def fast(request, nat):
    nat.check(nat.lib().jh_set_option(b"fast", request.param))
    yield request.param
    nat.check(nat.lib().jh_set_option(b"fast", 1))


def ragged_data(seed, n, lo=1, hi=120):
    rng = np.random.default_rng(seed)
    return dataset_of([random_dna(rng, int(L)) for L in rng.integers(lo, hi, n)])


MODELS = {
    "erg_s2_k1": lambda: ergodic_hmm(2, 1),
    "erg_s3_k0_map": lambda: ergodic_hmm(3, 0, ess=4.0),
    "erg_s5_k2": lambda: ergodic_hmm(5, 2),
    "erg_s8_k2": lambda: ergodic_hmm(8, 2),
    "erg_s16_k1": lambda: ergodic_hmm(16, 1),
    "erg_s32_k3": lambda: ergodic_hmm(32, 3, ess=4.0),
    "sparse": lambda: sparse_hmm(),
    "order2_s3": lambda: order2_hmm(3, 1),
    "banded_s40": lambda: banded_hmm(40, 4, 1),
    "sunflower_12": lambda: sunflower_hmm((5, 6), ess=2.0),                       # HMMFactory.createSunflowerHMM: 12 states (octets)
    "sunflower_41": lambda: sunflower_hmm((12, 9, 19), ess=0.0, start_central=False),  # 41 states (sparse tables)
}


@pytest.mark.parametrize("name", list(MODELS))
def test_loglik_matches_oracle(nat, oracle_lib, fast, name):
    hmm = MODELS[name]()
    om = oracle_lib.OracleModel(hmm.ir())
    data = ragged_data(1, 67)
    got = hmm.getLogScoreFor(data)
    ref = om.loglik(data.codes, data.offsets)
    np.testing.assert_allclose(got, ref, rtol=LL_RTOL, atol=0)
    assert np.abs(got - ref).max() <= 1e-11 * np.abs(ref).max() + 1e-13  # in practice ~1 ulp of the DP


@pytest.mark.parametrize("name", list(MODELS))
def test_viterbi_paths_bit_exact(nat, oracle_lib, name):
    hmm = MODELS[name]()
    om = oracle_lib.OracleModel(hmm.ir())
    data = ragged_data(2, 61)
    res = hmm.getViterbiPathsFor(data)
    rp, rs, _ = om.viterbi(data.codes, data.offsets)
    for (p, s), p_ref, s_ref in zip(res, rp, rs):
        assert np.array_equal(p, p_ref)
        assert s == s_ref  # bit-exact score: only adds and compares
    # one byte per residue (jh_viterbi_u8 / jh_posterior_decode_u8): same paths, a quarter of the device->host traffic
    for (p8, s8), (p, s) in zip(hmm.getViterbiPathsFor(data, narrow=True), res):
        assert p8.dtype == np.uint8 and np.array_equal(p8, p) and s8 == s
    for (p8, l8), (p, l) in zip(hmm.decodePosteriorPathsFor(data, narrow=True), hmm.decodePosteriorPathsFor(data)):
        assert p8.dtype == np.uint8 and np.array_equal(p8, p) and l8 == l


def test_viterbi_tie_rule_first_child_in_user_order(nat, oracle_lib):
    pars = BaumWelchParameterSet(1, IterationCondition(1), 1)
    for order in ([0, 1, 2], [2, 0, 1], [1, 2, 0]):
        em = [DiscreteEmission(DNA, 0.0) for _ in range(3)]
        te = [TransitionElement(None, order, None)] + [TransitionElement([s], order, None) for s in range(3)]
        hmm = HigherOrderHMM(pars, ["a", "b", "c"], em, *te)
        seq = Sequence.create(DNA, "ACGTTGCAAC")
        path, score = hmm.getViterbiPathFor(seq)
        assert list(path) == [order[0]] * 10
        om = oracle_lib.OracleModel(hmm.ir())
        rp, rs, _ = om.viterbi(seq.codes, np.array([0, 10]))
        assert score == rs[0] and np.array_equal(path, rp[0])


@pytest.mark.parametrize("name", list(MODELS))
def test_estep_counts_match_oracle(nat, oracle_lib, fast, name):
    hmm = MODELS[name]()
    om = oracle_lib.OracleModel(hmm.ir())
    data = ragged_data(3, 53)
    rng = np.random.default_rng(0)
    w = rng.random(data.getNumberOfElements()) + 0.5
    eng, nd = hmm._engine(), hmm._data(data)
    for weights in (None, w):
        nd.set_weights(weights)
        ts, es, sc = eng.estep(nd)
        rts, res, rsc = om.estep(data.codes, data.offsets, weights)
        np.testing.assert_allclose(ts, rts, rtol=COUNT_TOL, atol=COUNT_TOL)
        np.testing.assert_allclose(es, res, rtol=COUNT_TOL, atol=COUNT_TOL)
        assert sc == pytest.approx(rsc, rel=LL_RTOL)


@pytest.mark.parametrize("name,ess", [("erg_s2_k1", 0.0), ("erg_s3_k0_map", 4.0), ("erg_s8_k2", 0.0), ("sparse", 0.0),
                                      ("order2_s3", 0.0), ("erg_s32_k3", 4.0)])
def test_baum_welch_training_matches_oracle(nat, oracle_lib, fast, name, ess):
    hmm = MODELS[name]()
    om = oracle_lib.OracleModel(hmm.ir())
    data = planted_dna(60, 80, n_states=2, seed=5)
    n_iter = 6
    hmm.trainingParameter = BaumWelchParameterSet(1, IterationCondition(n_iter), 1)
    best = hmm.train(data)
    robj = om.train(data.codes, data.offsets, max_iter=n_iter)
    obj = np.array(hmm.last_objectives[0])
    assert len(obj) == n_iter  # N E-steps, N-1 M-steps (HigherOrderHMM.java:760-770)
    np.testing.assert_allclose(obj, robj, rtol=LL_RTOL)
    assert best == obj[-1]
    got = hmm._engine().get_params()
    ref = om.get_params()
    for g, r in zip(got, ref):
        fin = np.isfinite(r)
        assert np.array_equal(np.isfinite(g), fin)
        np.testing.assert_allclose(g[fin], r[fin], rtol=0, atol=PARAM_TOL)
    # the host mirror received the trained parameters (toXML / getCurrentParameterValues keep working)
    ir = hmm.ir()
    np.testing.assert_array_equal(ir.trans_param, got[0])
    np.testing.assert_array_equal(ir.emis_param, got[2])


def test_multi_start_training_keeps_the_best_start(nat, oracle_lib):
    """HigherOrderHMM.train with numberOfStarts > 1 (HigherOrderHMM.java:741-786): every start is initialised randomly, the
    parameters of the start with the largest final objective are kept."""
    hmm = ergodic_hmm(3, 1, n_iter=4)
    hmm.setSkiptInit(False)
    hmm.trainingParameter = BaumWelchParameterSet(4, IterationCondition(4), 1)
    data = planted_dna(80, 120, n_states=3, seed=9)
    best = hmm.train(data)
    finals = [obj[-1] for obj in hmm.last_objectives]
    assert len(finals) == 4 and len(set(finals)) > 1  # the starts really differ
    assert best == max(finals)
    # the model holds the parameters of the best start: the oracle, started from them, reports the same likelihood
    om = oracle_lib.OracleModel(hmm.ir())
    ll = om.loglik(data.codes, data.offsets).sum()
    np.testing.assert_allclose(hmm.getLogScoreFor(data).sum(), ll, rtol=LL_RTOL)
    # N E-steps and N-1 M-steps: the last objective was evaluated with exactly the parameters that are kept
    assert ll + hmm.getLogPriorTerm() == pytest.approx(best, rel=1e-9)


def test_small_difference_termination(nat, oracle_lib, fast):
    hmm = ergodic_hmm(3, 1)
    om = oracle_lib.OracleModel(hmm.ir())
    data = planted_dna(50, 60, n_states=3, seed=9)
    hmm.trainingParameter = BaumWelchParameterSet(1, SmallDifferenceOfFunctionEvaluationsCondition(1e-2), 1)
    hmm.train(data)
    robj = om.train(data.codes, data.offsets, term_kind=1, eps=1e-2, max_iter=0)
    obj = np.array(hmm.last_objectives[0])
    assert len(obj) == len(robj)
    np.testing.assert_allclose(obj, robj, rtol=LL_RTOL)


def test_viterbi_training_matches_oracle(nat, oracle_lib):
    hmm = ergodic_hmm(3, 1)
    om = oracle_lib.OracleModel(hmm.ir())
    data = planted_dna(40, 50, n_states=3, seed=2)
    eng, nd = hmm._engine(), hmm._data(data)
    ts, es, sc = eng.estep(nd, mode=nat.MODE_VITERBI)
    rts, res, rsc = om.estep(data.codes, data.offsets, mode=0)
    np.testing.assert_array_equal(ts, rts)  # integer counts
    np.testing.assert_array_equal(es, res)
    assert sc == pytest.approx(rsc, rel=1e-13)
    hmm.trainingParameter = ViterbiParameterSet(1, IterationCondition(4), 1)
    hmm.train(data)
    robj = om.train(data.codes, data.offsets, mode=0, max_iter=4)
    np.testing.assert_allclose(hmm.last_objectives[0], robj, rtol=1e-12)


@pytest.mark.parametrize("route", [1, 0], ids=["decode_and_count", "reference_order_kernel"])
@pytest.mark.parametrize("name", ["erg_s2_k1", "erg_s3_k0_map", "erg_s8_k2", "erg_s32_k3", "sparse", "banded_s40", "sunflower_41", "order2_s3"])
def test_viterbi_training_estep_both_routes(nat, oracle_lib, name, route):
    """E-step of Viterbi training (weights added along the best paths): the fast decoders + k_viterbi_path_counts
    (vit_train_fast = 1: every model the scaled tables cover) and the reference-order kernel with counts, weighted and
    unweighted, ragged data incl. sequences shorter than the emission order — counts equal the oracle's exactly."""
    hmm = MODELS[name]()
    om = oracle_lib.OracleModel(hmm.ir())
    rng = np.random.default_rng(31)
    lens = np.concatenate([rng.integers(1, 120, 60), [1, 2, 3, 257, 300]])
    data = dataset_of([random_dna(rng, int(L)) for L in lens])
    w = np.round(rng.random(lens.size) * 4 + 0.5, 3)
    eng, nd = hmm._engine(), hmm._data(data)
    eng.set_option("vit_train_fast", route)
    for weights in (None, w):
        nd.set_weights(weights)
        ts, es, sc = eng.estep(nd, mode=nat.MODE_VITERBI)
        rts, res, rsc = om.estep(data.codes, data.offsets, weights, mode=0)
        np.testing.assert_allclose(ts, rts, rtol=1e-13, atol=1e-12)  # sums of the same weights in another order
        np.testing.assert_allclose(es, res, rtol=1e-13, atol=1e-12)
        assert sc == pytest.approx(rsc, rel=1e-12)
    nd.set_weights(None)
    obj = eng.baum_welch(nd, nat.MODE_VITERBI, nat.TERM_ITERATIONS, 4, 0.0)
    np.testing.assert_allclose(obj, om.train(data.codes, data.offsets, mode=0, max_iter=4), rtol=1e-12)


@pytest.mark.parametrize("name", ["erg_s2_k1", "erg_s5_k2", "sparse", "order2_s3", "erg_s32_k3"])
def test_posterior_matrix_and_decoding(nat, oracle_lib, fast, name):
    hmm = MODELS[name]()
    om = oracle_lib.OracleModel(hmm.ir())
    data = ragged_data(4, 12, lo=3, hi=90)
    mats = hmm.getLogStatePosteriorMatrixFor(data)
    dec = hmm.decodePosteriorPathsFor(data)
    ll_ref = om.loglik(data.codes, data.offsets)
    for i in range(data.getNumberOfElements()):
        x = data.getElementAt(i).codes
        ref = om.posterior(x)
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(mats[i]), fin)
        np.testing.assert_allclose(mats[i][fin], ref[fin], rtol=0, atol=1e-10)
        ref_path = om.decode_posterior(ref)
        path, ll = dec[i]
        assert ll == pytest.approx(ll_ref[i], rel=LL_RTOL)
        if fast == 0:
            # reference-order log-space kernels: the arithmetic is the reference's, the decoded indices equal the oracle's outright
            assert np.array_equal(path, ref_path)
            assert np.array_equal(HigherOrderHMM.decodeStatePosterior(mats[i])[0], ref_path)
        else:
            # scaled kernels: index work is bit-exact wherever the oracle's own top-2 gap is above rounding noise
            srt = np.sort(ref, axis=0)
            clear = (srt[-1] - srt[-2] > 1e-9) if ref.shape[0] > 1 else np.ones(ref.shape[1], bool)
            assert np.array_equal(path[clear], ref_path[clear])
            assert np.array_equal(HigherOrderHMM.decodeStatePosterior(mats[i])[0], ref_path)  # jh_posterior is always log-space


def test_edge_cases(nat, oracle_lib):
    hmm = ergodic_hmm(3, 2)
    om = oracle_lib.OracleModel(hmm.ir())
    # empty data set
    empty = DataSet.from_packed(np.zeros(0, np.uint8), np.zeros(1, np.int64))
    assert hmm.getLogScoreFor(empty).size == 0
    assert hmm.getViterbiPathsFor(empty) == []
    # single residues and lengths below the emission order
    data = dataset_of([[0], [3], [1, 2], [2, 2, 2], [0, 1, 2, 3]])
    np.testing.assert_allclose(hmm.getLogScoreFor(data), om.loglik(data.codes, data.offsets), rtol=LL_RTOL)
    rp, rs, _ = om.viterbi(data.codes, data.offsets)
    for (p, s), pr, sr in zip(hmm.getViterbiPathsFor(data), rp, rs):
        assert np.array_equal(p, pr) and s == sr
    # an empty sequence is rejected like in the reference
    with pytest.raises(nat.NativeError) as ei:
        hmm.getLogScoreFor(DataSet.from_packed(np.array([1, 2], np.uint8), np.array([0, 0, 2], np.int64)))
    assert ei.value.status == -1
    # wrong alphabet
    bad = DataSet.from_packed(np.array([0, 1, 2, 3], np.uint8), np.array([0, 4], np.int64))
    bad.codes = np.array([0, 1, 7, 3], np.uint8)
    with pytest.raises(nat.NativeError) as ei:
        hmm.getLogScoreFor(bad)
    assert ei.value.status == -4
    assert isinstance(WrongAlphabetException("x"), Exception)


def test_zero_probability_parameters(nat, oracle_lib, fast):
    """-inf parameters (an M-step with a zero count and no prior yields log(0)): impossible paths are handled and an
    impossible sequence scores -inf on both sides."""
    hmm = sparse_hmm()
    e = hmm.emission[0]
    lp = e.params.copy()
    lp[:, 0] = -np.inf  # state a never emits A
    lp = lp - np.log(np.exp(lp).sum(1, keepdims=True))
    e.setLogProbabilities(lp)
    hmm._push_params()
    om = oracle_lib.OracleModel(hmm.ir())
    data = ragged_data(6, 40, lo=2, hi=50)
    ref = om.loglik(data.codes, data.offsets)
    got = hmm.getLogScoreFor(data)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin) and fin.any()
    np.testing.assert_allclose(got[fin], ref[fin], rtol=LL_RTOL)
    ts, es, sc = hmm._engine().estep(hmm._data(dataset_of([data.getElementAt(i).codes for i in np.nonzero(fin)[0]])))
    keep = dataset_of([data.getElementAt(i).codes for i in np.nonzero(fin)[0]])
    rts, res, rsc = om.estep(keep.codes, keep.offsets)
    np.testing.assert_allclose(ts, rts, rtol=COUNT_TOL, atol=COUNT_TOL)
    np.testing.assert_allclose(es, res, rtol=COUNT_TOL, atol=COUNT_TOL)


@pytest.mark.parametrize("variant", [0, 1, 3], ids=["octets_dmma", "warp_per_sequence", "two_sequences_per_warp"])
@pytest.mark.parametrize("name", ["erg_s2_k1", "erg_s3_k0_map", "sparse", "erg_s5_k2", "erg_s8_k2", "erg_s16_k1", "erg_s32_k3"])
def test_fast_estep_variants_agree_with_oracle(nat, oracle_lib, name, variant):
    """Every mapping of the scaled E-step (octets on the FP64 MMA path, lane-per-state SIMT) gives the oracle's counts:
    ragged octets, a last octet with empty rows, weights, sequences shorter than the emission order."""
    nat.check(nat.lib().jh_set_option(b"fast", 1))
    nat.check(nat.lib().jh_set_option(b"fast_variant", variant))
    try:
        hmm = MODELS[name]()
        om = oracle_lib.OracleModel(hmm.ir())
        rng = np.random.default_rng(11)
        lens = np.concatenate([rng.integers(1, 6, 9), rng.integers(30, 400, 50), [1, 2, 3, 700]])
        data = dataset_of([random_dna(rng, int(L)) for L in lens])
        w = rng.random(lens.size) + 0.25
        eng, nd = hmm._engine(), hmm._data(data)
        for weights in (None, w):
            nd.set_weights(weights)
            ts, es, sc = eng.estep(nd)
            rts, res, rsc = om.estep(data.codes, data.offsets, weights)
            np.testing.assert_allclose(ts, rts, rtol=COUNT_TOL, atol=COUNT_TOL)
            np.testing.assert_allclose(es, res, rtol=COUNT_TOL, atol=COUNT_TOL)
            assert sc == pytest.approx(rsc, rel=LL_RTOL)
        nd.set_weights(None)
        np.testing.assert_allclose(eng.loglik(nd), om.loglik(data.codes, data.offsets), rtol=LL_RTOL)
    finally:
        nat.check(nat.lib().jh_set_option(b"fast_variant", 0))


@pytest.mark.parametrize("chunk", [8, 13, 64])
@pytest.mark.parametrize("name", ["erg_s5_k2", "erg_s8_k2", "erg_s16_k1", "erg_s32_k3"])
def test_chunked_octet_estep_matches_oracle(nat, oracle_lib, name, chunk):
    """Octet E-step with checkpointed forward rows (long sequences): chunk boundaries inside the ragged tail of an
    octet, chunks shorter than the emission order + rescale cadence, sequences ending exactly at a boundary."""
    nat.check(nat.lib().jh_set_option(b"oct_chunk", chunk))
    nat.check(nat.lib().jh_set_option(b"oct_chunk_force", 1))
    try:
        hmm = MODELS[name]()
        om = oracle_lib.OracleModel(hmm.ir())
        rng = np.random.default_rng(chunk)
        lens = np.concatenate([rng.integers(2 * chunk + 1, 6 * chunk, 21), [4 * chunk, 4 * chunk + 1, 4 * chunk - 1, 3 * chunk, 700, 650, 1, 5]])
        data = dataset_of([random_dna(rng, int(L)) for L in lens])
        w = rng.random(lens.size) + 0.25
        eng, nd = hmm._engine(), hmm._data(data)
        for weights in (None, w):
            nd.set_weights(weights)
            ts, es, sc = eng.estep(nd)
            rts, res, rsc = om.estep(data.codes, data.offsets, weights)
            np.testing.assert_allclose(ts, rts, rtol=COUNT_TOL, atol=COUNT_TOL)
            np.testing.assert_allclose(es, res, rtol=COUNT_TOL, atol=COUNT_TOL)
            assert sc == pytest.approx(rsc, rel=LL_RTOL)
        nd.set_weights(None)
        # posterior decoding through the same chunking (k_mma_decode<NT, true>)
        dec = hmm.decodePosteriorPathsFor(data)
        ll_ref = om.loglik(data.codes, data.offsets)
        for i in range(data.getNumberOfElements()):
            ref = om.posterior(data.getElementAt(i).codes)
            ref_path = om.decode_posterior(ref)
            path, ll = dec[i]
            assert ll == pytest.approx(ll_ref[i], rel=LL_RTOL)
            srt = np.sort(ref, axis=0)
            clear = srt[-1] - srt[-2] > 1e-9
            assert path.shape == ref_path.shape and np.array_equal(path[clear], ref_path[clear])
    finally:
        nat.check(nat.lib().jh_set_option(b"oct_chunk", -1))
        nat.check(nat.lib().jh_set_option(b"oct_chunk_force", 0))


@pytest.mark.parametrize("chunked", [0, 1], ids=["full_rows", "chunked"])
@pytest.mark.parametrize("name", ["erg_s5_k2", "erg_s8_k2", "erg_s16_k1", "erg_s32_k3"])
def test_deterministic_estep_is_bit_identical_run_to_run(nat, oracle_lib, name, chunked):
    """Option "deterministic" (a fixed-order E-step): static octet assignment, fixed-point emission counts,
    per-slot partial sums folded in slot order.  Repeated E-steps and repeated EM runs give bit-identical statistics,
    objectives and parameters, and still match the oracle; without the option the same calls are only equal to ~1e-13."""
    hmm = MODELS[name]()
    om = oracle_lib.OracleModel(hmm.ir())
    rng = np.random.default_rng(23)
    lens = rng.integers(20, 200, 800)
    data = dataset_of([random_dna(rng, int(L)) for L in lens])
    w = rng.random(lens.size) + 0.25
    eng, nd = hmm._engine(), hmm._data(data)
    eng.set_option("deterministic", 1)
    if chunked:
        eng.set_option("oct_chunk", 24)
        eng.set_option("oct_chunk_force", 1)
    for weights in (None, w):
        nd.set_weights(weights)
        runs = [eng.estep(nd) for _ in range(4)]
        for ts, es, sc in runs[1:]:
            assert np.array_equal(ts, runs[0][0]) and np.array_equal(es, runs[0][1]) and sc == runs[0][2]
        rts, res, rsc = om.estep(data.codes, data.offsets, weights)
        np.testing.assert_allclose(runs[0][0], rts, rtol=COUNT_TOL, atol=COUNT_TOL)
        np.testing.assert_allclose(runs[0][1], res, rtol=COUNT_TOL, atol=COUNT_TOL)
        assert runs[0][2] == pytest.approx(rsc, rel=LL_RTOL)
    nd.set_weights(None)
    p0 = eng.get_params()
    objs, pars = [], []
    for _ in range(3):
        eng.set_params(*p0)
        objs.append(eng.baum_welch(nd, nat.MODE_BAUM_WELCH, nat.TERM_ITERATIONS, 5, 0.0))
        pars.append(eng.get_params())
    for o, p in zip(objs[1:], pars[1:]):
        assert np.array_equal(o, objs[0]) and all(np.array_equal(x, y) for x, y in zip(p, pars[0]))
    np.testing.assert_allclose(objs[0], om.train(data.codes, data.offsets, max_iter=5), rtol=1e-9)


def test_deterministic_mode_refuses_what_it_does_not_cover(nat):
    """Small models (lane kernels) and the sparse / long-sequence path have no fixed-order variant: the option is an error there."""
    hmm = ergodic_hmm(2, 1)
    rng = np.random.default_rng(1)
    data = dataset_of([random_dna(rng, 50) for _ in range(40)])
    eng, nd = hmm._engine(), hmm._data(data)
    eng.set_option("deterministic", 1)
    with pytest.raises(nat.NativeError) as ei:
        eng.estep(nd)
    assert ei.value.status == -2 and "deterministic" in str(ei.value)


@pytest.mark.parametrize("name", ["erg_s5_k2", "erg_s8_k2", "erg_s16_k1", "erg_s32_k3"])
def test_posterior_decoding_octets_ragged(nat, oracle_lib, fast, name):
    """decodeStatePosterior(getStatePosteriorMatrixFor(seq)) fused on the device for a ragged data set (octets with
    unequal rows, a last octet with empty rows, sequences shorter than the emission order)."""
    hmm = MODELS[name]()
    om = oracle_lib.OracleModel(hmm.ir())
    rng = np.random.default_rng(21)
    lens = np.concatenate([rng.integers(1, 5, 6), rng.integers(20, 300, 37), [1, 2, 450]])
    data = dataset_of([random_dna(rng, int(L)) for L in lens])
    dec = hmm.decodePosteriorPathsFor(data)
    ll_ref = om.loglik(data.codes, data.offsets)
    for i in range(data.getNumberOfElements()):
        ref = om.posterior(data.getElementAt(i).codes)
        ref_path = om.decode_posterior(ref)
        path, ll = dec[i]
        assert ll == pytest.approx(ll_ref[i], rel=LL_RTOL)
        srt = np.sort(ref, axis=0)
        clear = srt[-1] - srt[-2] > 1e-9
        assert path.shape == ref_path.shape and np.array_equal(path[clear], ref_path[clear])
        assert clear.mean() > 0.95


@pytest.mark.parametrize("states,order", [(32, 4), (8, 5), (16, 4), (4, 6), (2, 7)])
def test_large_emission_tables_stay_on_the_fast_path(nat, oracle_lib, states, order):
    """High emission orders: the [context hash][state] table exceeds 64 KB and is read from global memory (EG kernels)
    instead of shared memory: scoring, bit-exact Viterbi, posterior decoding, E-step counts (ragged, weighted), training."""
    hmm = ergodic_hmm(states, order, n_iter=3, ess=2.0)
    om = oracle_lib.OracleModel(hmm.ir())
    data = ragged_data(order, 45, lo=1, hi=150)
    np.testing.assert_allclose(hmm.getLogScoreFor(data), om.loglik(data.codes, data.offsets), rtol=LL_RTOL)
    rp, rs, _ = om.viterbi(data.codes, data.offsets)
    for (p, s), p_ref, s_ref in zip(hmm.getViterbiPathsFor(data), rp, rs):
        assert np.array_equal(p, p_ref) and s == s_ref
    eng, nd = hmm._engine(), hmm._data(data)
    w = np.random.default_rng(order).random(data.getNumberOfElements()) + 0.5
    for weights in (None, w):
        nd.set_weights(weights)
        ts, es, sc = eng.estep(nd)
        rts, res, rsc = om.estep(data.codes, data.offsets, weights)
        np.testing.assert_allclose(ts, rts, rtol=COUNT_TOL, atol=COUNT_TOL)
        np.testing.assert_allclose(es, res, rtol=COUNT_TOL, atol=COUNT_TOL)
        assert sc == pytest.approx(rsc, rel=LL_RTOL)
    nd.set_weights(None)
    dec = hmm.decodePosteriorPathsFor(data)
    for i in range(0, data.getNumberOfElements(), 5):
        ref = om.posterior(data.getElementAt(i).codes)
        path, _ = dec[i]
        srt = np.sort(ref, axis=0)
        clear = srt[-1] - srt[-2] > 1e-9
        assert np.array_equal(path[clear], om.decode_posterior(ref)[clear])
    hmm.train(data)
    np.testing.assert_allclose(hmm.last_objectives[0], om.train(data.codes, data.offsets, max_iter=3), rtol=LL_RTOL)


def test_transition_order_zero(nat, oracle_lib):
    """getMaximalMarkovOrder() == 0: one context for every layer, every context final (HigherOrderHMM.java:498, 1046-1058);
    runs on the general context-graph kernels.  Scoring, windows, bit-exact Viterbi, counts, Baum-Welch and Viterbi training."""
    hmm = order0_hmm(ess=2.0)
    om = oracle_lib.OracleModel(hmm.ir())
    data = ragged_data(4, 40, lo=1, hi=60)
    np.testing.assert_allclose(hmm.getLogScoreFor(data), om.loglik(data.codes, data.offsets), rtol=LL_RTOL)
    res = hmm.getViterbiPathsFor(data)
    rp, rs, _ = om.viterbi(data.codes, data.offsets)
    for (p, s), p_ref, s_ref in zip(res, rp, rs):
        assert np.array_equal(p, p_ref) and s == s_ref
    eng, nd = hmm._engine(), hmm._data(data)
    ts, es, sc = eng.estep(nd)
    rts, res_, rsc = om.estep(data.codes, data.offsets)
    np.testing.assert_allclose(ts, rts, rtol=COUNT_TOL, atol=COUNT_TOL)
    np.testing.assert_allclose(es, res_, rtol=COUNT_TOL, atol=COUNT_TOL)
    assert sc == pytest.approx(rsc, rel=LL_RTOL)
    # state posteriors of order-0 transitions are a special case of the reference (HigherOrderHMM.java:296-308)
    mats = hmm.getLogStatePosteriorMatrixFor(data)
    dec = hmm.decodePosteriorPathsFor(data)
    ll_ref = om.loglik(data.codes, data.offsets)
    for i in range(data.getNumberOfElements()):
        ref = om.posterior(data.getElementAt(i).codes)
        np.testing.assert_allclose(mats[i], ref, rtol=0, atol=1e-12)
        assert np.array_equal(dec[i][0], om.decode_posterior(ref))
        assert dec[i][1] == pytest.approx(ll_ref[i], rel=LL_RTOL)
    x = data.getElementAt(0).codes
    if x.size >= 4:
        np.testing.assert_allclose(hmm.getLogStatePosteriorMatrixFor(data.getElementAt(0), 1, x.size - 2), om.posterior_window(x, 1, x.size - 2), rtol=0, atol=1e-12)
    hmm.trainingParameter = BaumWelchParameterSet(1, IterationCondition(4), 1)
    hmm.train(data)
    np.testing.assert_allclose(hmm.last_objectives[0], om.train(data.codes, data.offsets, max_iter=4), rtol=LL_RTOL)
    for g, r in zip(hmm._engine().get_params(), om.get_params()):
        fin = np.isfinite(r)
        np.testing.assert_allclose(g[fin], r[fin], rtol=0, atol=PARAM_TOL)


# ---- general context graph: silent states (profile-like HMM), one sequence per thread --------------------------
def test_silent_states_loglik_viterbi_counts_training(nat, oracle_lib):
    """silent states + trailing silent states on the GPU, reference order (kernels_general.cuh)."""
    hmm = silent_hmm()
    om = oracle_lib.OracleModel(hmm.ir())
    data = ragged_data(8, 45, lo=1, hi=40)
    # scoring
    np.testing.assert_allclose(hmm.getLogScoreFor(data), om.loglik(data.codes, data.offsets), rtol=1e-12)
    # Viterbi: paths contain the silent states, bit-exact scores
    rp, rs, rl = om.viterbi(data.codes, data.offsets)
    res = hmm.getViterbiPathsFor(data)
    for (p, sc), p_ref, s_ref in zip(res, rp, rs):
        assert np.array_equal(p, p_ref) and sc == s_ref
    assert any(len(p) > L for (p, _), L in zip(res, np.diff(data.offsets)))  # silent states really are on the paths
    # E-step, both modes
    eng, nd = hmm._engine(), hmm._data(data)
    w = np.random.default_rng(1).random(data.getNumberOfElements()) + 0.5
    for weights in (None, w):
        nd.set_weights(weights)
        ts, es, sc = eng.estep(nd)
        rts, res_, rsc = om.estep(data.codes, data.offsets, weights)
        np.testing.assert_allclose(ts, rts, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(es, res_, rtol=1e-10, atol=1e-12)
        assert sc == pytest.approx(rsc, rel=1e-12)
        ts, es, sc = eng.estep(nd, mode=nat.MODE_VITERBI)
        rts, res_, rsc = om.estep(data.codes, data.offsets, weights, mode=0)
        np.testing.assert_allclose(ts, rts, rtol=1e-13)
        np.testing.assert_allclose(es, res_, rtol=1e-13)
        assert sc == pytest.approx(rsc, rel=1e-13)
    nd.set_weights(None)
    # state posteriors (silent states score -inf: silentZero, HigherOrderHMM.java:316) and posterior decoding
    mats = hmm.getLogStatePosteriorMatrixFor(data)
    dec = hmm.decodePosteriorPathsFor(data)
    ll_ref = om.loglik(data.codes, data.offsets)
    for i in range(data.getNumberOfElements()):
        ref = om.posterior(data.getElementAt(i).codes)
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(mats[i]), fin)
        np.testing.assert_allclose(mats[i][fin], ref[fin], rtol=0, atol=1e-10)
        path, ll = dec[i]
        assert ll == pytest.approx(ll_ref[i], rel=1e-12)
        srt = np.sort(ref, axis=0)
        clear = srt[-1] - srt[-2] > 1e-9
        assert np.array_equal(path[clear], om.decode_posterior(ref)[clear])
    # EM
    hmm.trainingParameter = BaumWelchParameterSet(1, IterationCondition(5), 1)
    hmm.train(data)
    robj = om.train(data.codes, data.offsets, max_iter=5)
    np.testing.assert_allclose(hmm.last_objectives[0], robj, rtol=LL_RTOL)
    for g, r in zip(hmm._engine().get_params(), om.get_params()):
        fin = np.isfinite(r)
        assert np.array_equal(np.isfinite(g), fin)
        np.testing.assert_allclose(g[fin], r[fin], rtol=0, atol=PARAM_TOL)


@pytest.mark.parametrize("name", ["erg_s3_k0_map", "erg_s5_k2", "order2_s3", "sparse", "silent"])
def test_window_scoring_sees_the_residues_before_the_window(nat, oracle_lib, name):
    """getLogProbFor(seq, start, end) for many windows in one call: layer 0 is the window start, the emission contexts
    use absolute positions (HigherOrderDiscreteEmission.java:138-146)."""
    hmm = silent_hmm() if name == "silent" else MODELS[name]()
    om = oracle_lib.OracleModel(hmm.ir())
    data = ragged_data(9, 7, lo=30, hi=80)
    rng = np.random.default_rng(4)
    si, st, en = [], [], []
    for i in range(data.getNumberOfElements()):
        L = data.getElementAt(i).getLength()
        for _ in range(12):
            a = int(rng.integers(0, L))
            b = int(rng.integers(a, L))
            si.append(i); st.append(a); en.append(b)
        si.append(i); st.append(0); en.append(L - 1)
    got = hmm.getLogProbForWindows(data, si, st, en)
    ref = np.array([om.loglik_window(data.getElementAt(i).codes, a, b) for i, a, b in zip(si, st, en)])
    np.testing.assert_allclose(got, ref, rtol=1e-12)
    seq = data.getElementAt(2)
    assert hmm.getLogProbFor(seq, 5, 20) == pytest.approx(om.loglik_window(seq.codes, 5, 20), rel=1e-12)
    with pytest.raises(nat.NativeError):
        hmm.getLogProbForWindows(data, [0], [5], [10_000])


@pytest.mark.parametrize("name", ["erg_s3_k0_map", "erg_s5_k2", "erg_s32_k3", "order2_s3", "sparse", "silent"])
def test_window_viterbi_and_posterior(nat, oracle_lib, name):
    """getViterbiPathFor(startPos, endPos, seq) / getLogStatePosteriorMatrixFor(startPos, endPos, seq): layer 0 is the window
    start, higher-order emissions see the residues before it (HigherOrderHMM.java:590-598, AbstractHMM.java:776-797)."""
    hmm = silent_hmm() if name == "silent" else MODELS[name]()
    om = oracle_lib.OracleModel(hmm.ir())
    rng = np.random.default_rng(17)
    x = random_dna(rng, 90)
    seq = Sequence(DNA, x)
    for s0, s1 in [(0, 89), (5, 40), (37, 37), (60, 89), (1, 2)]:
        path, score = hmm.getViterbiPathFor(seq, s0, s1)
        rp, rs = om.viterbi_window(x, s0, s1)
        assert np.array_equal(path, rp) and score == rs  # bit-exact, tie rule included
        post = hmm.getLogStatePosteriorMatrixFor(seq, s0, s1)
        ref = om.posterior_window(x, s0, s1)
        fin = np.isfinite(ref)
        assert post.shape == ref.shape and np.array_equal(np.isfinite(post), fin)
        np.testing.assert_allclose(post[fin], ref[fin], rtol=1e-9, atol=1e-9)
    # many windows in one call
    eng, nd = hmm._engine(), hmm._data(dataset_of([x, x[::-1].copy()]))
    si, st, en = [0, 1, 1, 0], [3, 0, 10, 80], [50, 89, 10, 89]
    n_silent = sum(1 for i in hmm.emissionIdx if hmm.emission[i].order < 0)
    paths, scores, plen = eng.viterbi_windows(nd, si, st, en, n_silent)
    for k in range(4):
        rp, rs = om.viterbi_window(x if si[k] == 0 else x[::-1].copy(), st[k], en[k])
        assert np.array_equal(paths[k], rp) and scores[k] == rs
    with pytest.raises(nat.NativeError):
        eng.viterbi_windows(nd, [0], [10], [90], n_silent)  # window end outside the sequence


# ---- sparse path: 33..256 states, checkpointed chunk-parallel decoding (BASELINE cfg5 shape) ------------------------
@pytest.mark.parametrize("states,degree,order", [(40, 4, 1), (128, 4, 0), (128, 4, 1), (64, 8, 0), (200, 3, 0),
                                                 (128, 4, 3), (40, 4, 5)])  # the last two: emission table in global memory
@pytest.mark.parametrize("chunk,cta,overlap", [(16, 0, 0), (64, 0, 0), (16, 1, 0), (16, 1, 1), (64, 1, 1), (2048, 1, 1)],
                         ids=["warp_16", "warp_64", "cta_16", "cta_streams_16", "cta_streams_64", "cta_streams_2048"])
def test_sparse_models_scoring_and_chunked_decoding(nat, oracle_lib, states, degree, order, chunk, cta, overlap):
    """cta: the checkpoint pass as one warp per (sequence, direction) (many sequences) or one CTA with a state per thread
    and the lazy rescale (few long sequences: latency mode); overlap: one stream per long sequence, its chunk pass starts
    when its own checkpoint pass is done."""
    nat.check(nat.lib().jh_set_option(b"pass1_cta", cta))
    nat.check(nat.lib().jh_set_option(b"overlap", overlap))
    hmm = banded_hmm(states, degree, order)
    om = oracle_lib.OracleModel(hmm.ir())
    rng = np.random.default_rng(states + chunk)
    lens = np.concatenate([rng.integers(1, 6, 4), rng.integers(40, 400, 9), [chunk, chunk + 1, 2 * chunk - 1, 1]])
    data = dataset_of([random_dna(rng, int(L)) for L in lens])
    nat.check(nat.lib().jh_set_option(b"chunk", chunk))
    try:
        ll_ref = om.loglik(data.codes, data.offsets)
        np.testing.assert_allclose(hmm.getLogScoreFor(data), ll_ref, rtol=LL_RTOL)
        dec = hmm.decodePosteriorPathsFor(data)
        n_clear = n_all = 0
        for i in range(data.getNumberOfElements()):
            ref = om.posterior(data.getElementAt(i).codes)
            ref_path = om.decode_posterior(ref)
            path, ll = dec[i]
            assert ll == pytest.approx(ll_ref[i], rel=LL_RTOL)
            srt = np.sort(ref, axis=0)
            clear = srt[-1] - srt[-2] > 1e-9
            assert path.shape == ref_path.shape and np.array_equal(path[clear], ref_path[clear])
            n_clear += clear.sum(); n_all += clear.size
        assert n_clear > 0.9 * n_all
        # Baum-Welch E-step of the same path: checkpoints + chunk-parallel counts (k_sparse_pass1 + k_sparse_estep_chunks)
        w = rng.uniform(0.5, 2.0, data.getNumberOfElements())
        eng, nd = hmm._engine(), hmm._data(data)
        for weights in (None, w):
            nd.set_weights(weights)
            ts, es, sc = eng.estep(nd)
            rts, res, rsc = om.estep(data.codes, data.offsets, weights)
            np.testing.assert_allclose(ts, rts, rtol=COUNT_TOL, atol=COUNT_TOL)
            np.testing.assert_allclose(es, res, rtol=COUNT_TOL, atol=COUNT_TOL)
            assert sc == pytest.approx(rsc, rel=LL_RTOL)
        nd.set_weights(None)
    finally:
        nat.check(nat.lib().jh_set_option(b"chunk", 2048))
        nat.check(nat.lib().jh_set_option(b"pass1_cta", 1))
        nat.check(nat.lib().jh_set_option(b"scan", 1))
        nat.check(nat.lib().jh_set_option(b"overlap", 1))


@pytest.mark.parametrize("cta,scan,lane", [(0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 1, 1)],
                         ids=["warp_per_sequence", "cta_per_sequence", "parallel_scan", "parallel_scan_lane_chunks"])
@pytest.mark.parametrize("name", ["erg_s2_k1", "sparse", "erg_s5_k2", "erg_s16_k1", "erg_s32_k3", "erg_s3_k0_map"])
def test_warp_per_sequence_path_for_small_models(nat, oracle_lib, name, cta, scan, lane):
    """The checkpointed warp-per-sequence path (taken for a few long sequences) also covers models with <= 32 states:
    forced here with fast_variant 4 and a tiny chunk length."""
    hmm = MODELS[name]()
    om = oracle_lib.OracleModel(hmm.ir())
    rng = np.random.default_rng(5)
    lens = np.concatenate([rng.integers(1, 6, 3), rng.integers(50, 500, 8), [32, 33, 63]])
    data = dataset_of([random_dna(rng, int(L)) for L in lens])
    nat.check(nat.lib().jh_set_option(b"fast_variant", 4))
    nat.check(nat.lib().jh_set_option(b"chunk", 32))
    nat.check(nat.lib().jh_set_option(b"pass1_cta", cta))
    nat.check(nat.lib().jh_set_option(b"scan", scan))  # checkpoint pass parallel in the position (kernels_scan.cuh)
    nat.check(nat.lib().jh_set_option(b"lane_chunks", lane))  # 2..4 states: chunk pass with one chunk per lane (kernels_lanechunks.cuh)
    if lane and hmm.getNumberOfStates() > 8:
        pytest.skip("one chunk per lane is for 2..8 states")
    try:
        ll_ref = om.loglik(data.codes, data.offsets)
        np.testing.assert_allclose(hmm.getLogScoreFor(data), ll_ref, rtol=LL_RTOL)
        dec = hmm.decodePosteriorPathsFor(data)
        for i in range(data.getNumberOfElements()):
            ref = om.posterior(data.getElementAt(i).codes)
            path, ll = dec[i]
            assert ll == pytest.approx(ll_ref[i], rel=LL_RTOL)
            srt = np.sort(ref, axis=0)
            clear = srt[-1] - srt[-2] > 1e-9
            assert np.array_equal(path[clear], om.decode_posterior(ref)[clear])
        ts, es, sc = hmm._engine().estep(hmm._data(data))
        rts, res, rsc = om.estep(data.codes, data.offsets)
        np.testing.assert_allclose(ts, rts, rtol=COUNT_TOL, atol=COUNT_TOL)
        np.testing.assert_allclose(es, res, rtol=COUNT_TOL, atol=COUNT_TOL)
        assert sc == pytest.approx(rsc, rel=LL_RTOL)
        hmm.trainingParameter = BaumWelchParameterSet(1, IterationCondition(3), 1)
        hmm.train(data)
        robj = om.train(data.codes, data.offsets, max_iter=3)
        np.testing.assert_allclose(hmm.last_objectives[0], robj, rtol=LL_RTOL)
    finally:
        nat.check(nat.lib().jh_set_option(b"fast_variant", 0))
        nat.check(nat.lib().jh_set_option(b"chunk", 2048))
        nat.check(nat.lib().jh_set_option(b"pass1_cta", 1))
        nat.check(nat.lib().jh_set_option(b"scan", 1))
        nat.check(nat.lib().jh_set_option(b"lane_chunks", 1))


@pytest.mark.parametrize("seed", range(8))
def test_parallel_checkpoint_pass_matches_the_sequential_one_randomized(nat, seed):
    """Random models / chunk lengths / ragged long sequences: the position-parallel checkpoint pass (+ lane-per-chunk pass)
    and the sequential pass are two evaluation orders of the same sums: scores, counts and decoded paths agree to 1e-11."""
    rng = np.random.default_rng(100 + seed)
    states = int(rng.choice([2, 3, 4, 6, 8, 13, 16, 27, 32]))
    order = int(rng.integers(0, 3))
    chunk = int(rng.choice([16, 24, 40, 128]))
    hmm = ergodic_hmm(states, order, n_iter=2, ess=float(rng.choice([0.0, 3.0])))
    lens = np.concatenate([rng.integers(1, 2 * chunk, 5), rng.integers(3 * chunk, 40 * chunk, 6), [chunk, chunk + 1, 7 * chunk]])
    data = dataset_of([random_dna(rng, int(L)) for L in lens])
    w = rng.random(lens.size) + 0.5
    out = {}
    try:
        nat.check(nat.lib().jh_set_option(b"fast_variant", 4))
        nat.check(nat.lib().jh_set_option(b"chunk", chunk))
        for mode in (0, 1):
            nat.check(nat.lib().jh_set_option(b"scan", mode))
            nat.check(nat.lib().jh_set_option(b"lane_chunks", mode))
            eng, nd = hmm._engine(), hmm._data(data)
            nd.set_weights(w)
            ts, es, sc = eng.estep(nd)
            nd.set_weights(None)
            paths, ll = eng.posterior_decode(nd)
            out[mode] = (ts, es, sc, eng.loglik(nd), ll, paths)
    finally:
        for k, v in ((b"fast_variant", 0), (b"chunk", 2048), (b"scan", 1), (b"lane_chunks", 1)):
            nat.check(nat.lib().jh_set_option(k, v))
    a, b = out[0], out[1]
    np.testing.assert_allclose(b[0], a[0], rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(b[1], a[1], rtol=1e-11, atol=1e-11)
    assert b[2] == pytest.approx(a[2], rel=1e-12)
    np.testing.assert_allclose(b[3], a[3], rtol=1e-12)
    np.testing.assert_allclose(b[4], a[4], rtol=1e-12)
    assert (a[5] != b[5]).mean() < 1e-3  # decisions differ only where two posteriors tie to the last bits


@pytest.mark.parametrize("states,order", [(2, 1), (3, 0), (4, 2)])
def test_short_sequences_of_tiny_models_are_split_into_chunks(nat, oracle_lib, states, order):
    """2..4 states, a data set too small to fill the GPU with a sequence per lane (cfg1's shape): the library cuts the
    sequences into chunks and runs the position-parallel path by itself (split_short).  Same results as the oracle, and as
    the sequence-per-lane kernels with the rule switched off."""
    hmm = ergodic_hmm(states, order, n_iter=3, ess=1.0)
    om = oracle_lib.OracleModel(hmm.ir())
    rng = np.random.default_rng(states)
    lens = rng.integers(300, 900, 1800)
    lens[:3] = [1, 2, 64]
    data = make_data(lens.size, lens)
    assert data.getNumberOfResidues() >= 1 << 20
    w = rng.random(lens.size) + 0.5
    eng, nd = hmm._engine(), hmm._data(data)
    res = {}
    for split in (1, 0):
        nat.check(nat.lib().jh_set_option(b"split_short", split))
        try:
            nd.set_weights(w)
            ts, es, sc = eng.estep(nd)
            nd.set_weights(None)
            paths, ll = eng.posterior_decode(nd)
            res[split] = (ts, es, sc, eng.loglik(nd), ll, paths)
        finally:
            nat.check(nat.lib().jh_set_option(b"split_short", 1))
    rts, res_, rsc = om.estep(data.codes, data.offsets, w)
    ll_ref = om.loglik(data.codes, data.offsets)
    for split in (1, 0):
        ts, es, sc, ll1, ll2, paths = res[split]
        np.testing.assert_allclose(ts, rts, rtol=COUNT_TOL, atol=COUNT_TOL)
        np.testing.assert_allclose(es, res_, rtol=COUNT_TOL, atol=COUNT_TOL)
        assert sc == pytest.approx(rsc, rel=LL_RTOL)
        np.testing.assert_allclose(ll1, ll_ref, rtol=LL_RTOL)
        np.testing.assert_allclose(ll2, ll_ref, rtol=LL_RTOL)
    assert (res[0][5] != res[1][5]).mean() < 1e-4  # decisions differ only where two posteriors tie to the last bits
    off = data.offsets
    for i in range(0, lens.size, 97):
        ref = om.posterior(data.getElementAt(i).codes)
        srt = np.sort(ref, axis=0)
        clear = srt[-1] - srt[-2] > 1e-9
        assert np.array_equal(res[1][5][off[i]:off[i + 1]][clear], om.decode_posterior(ref)[clear])
    hmm.train(data)
    np.testing.assert_allclose(hmm.last_objectives[0], om.train(data.codes, data.offsets, max_iter=3), rtol=LL_RTOL)


def test_full_size_properties_cfg3_sample(nat):
    """BASELINE cfg3 model (32 states, order-3 emissions) on 16k x 1 kb sequences: size-independent properties.
    Transition counts add up to the residues, emission counts to the residues with a full context, the score is the
    sum of the per-sequence log-likelihoods, and the octet kernel agrees with the lane-per-state kernel."""
    hmm = ergodic_hmm(32, 3, n_iter=3, ess=4.0)
    data = make_data(16_000, 1000)
    eng, nd = hmm._engine(), hmm._data(data)
    ts, es, sc = eng.estep(nd)
    assert ts.sum() == pytest.approx(1.6e7, rel=1e-10)
    assert es.sum() == pytest.approx(1.6e7 - 3 * 1.6e4, rel=1e-10)
    ll = eng.loglik(nd)
    assert sc == pytest.approx(ll.sum(), rel=1e-11)
    nat.check(nat.lib().jh_set_option(b"fast_variant", 3))
    try:
        ts2, es2, sc2 = eng.estep(nd)
    finally:
        nat.check(nat.lib().jh_set_option(b"fast_variant", 0))
    np.testing.assert_allclose(ts, ts2, rtol=1e-10)
    np.testing.assert_allclose(es, es2, rtol=1e-10, atol=1e-12)
    assert sc == pytest.approx(sc2, rel=1e-13)
    hmm.train(data)
    assert (np.diff(hmm.last_objectives[0]) >= -1e-6).all()


def test_java_random_workload_and_ragged_buckets(nat, oracle_lib, fast):
    """cfg4-style ragged lengths on the java.util.Random stream; length buckets must not change results or order."""
    lens = ragged_lengths(48, seed=42, lo=20.0, decades=2.0)
    data = make_data(48, lens)
    hmm = ergodic_hmm(16, 1)
    om = oracle_lib.OracleModel(hmm.ir())
    np.testing.assert_allclose(hmm.getLogScoreFor(data), om.loglik(data.codes, data.offsets), rtol=LL_RTOL)
    ts, es, sc = hmm._engine().estep(hmm._data(data))
    rts, res, rsc = om.estep(data.codes, data.offsets)
    np.testing.assert_allclose(ts, rts, rtol=COUNT_TOL, atol=COUNT_TOL)
    np.testing.assert_allclose(es, res, rtol=COUNT_TOL, atol=COUNT_TOL)
    assert sc == pytest.approx(rsc, rel=LL_RTOL)


def test_full_size_properties_cfg1(nat, fast):
    """BASELINE cfg1 at full size (10k x 1 kb): size-independent properties instead of an oracle run:
    counts add up (transitions = sum L_n, emissions = sum (L_n - k)), score == sum of per-sequence log-likelihoods,
    Viterbi score <= log-likelihood, EM objective is monotone."""
    hmm = ergodic_hmm(2, 1, n_iter=4)
    data = make_data(10_000, 1000)
    eng, nd = hmm._engine(), hmm._data(data)
    ts, es, sc = eng.estep(nd)
    assert ts.sum() == pytest.approx(1e7, rel=1e-10)
    assert es.sum() == pytest.approx(1e7 - 1e4, rel=1e-10)
    ll = eng.loglik(nd)
    assert sc == pytest.approx(ll.sum(), rel=1e-11)
    paths, vs = eng.viterbi(nd)
    assert (vs <= ll + 1e-9).all() and paths.min() >= 0 and paths.max() <= 1
    hmm.train(data)
    assert (np.diff(hmm.last_objectives[0]) >= -1e-6).all()


# ---- chromosome-scale Viterbi: checkpointed sweep + per-chunk recomputation of the decisions (kernels_vitlong.cuh) ----
@pytest.mark.parametrize("name,chunk", [("banded_s128", 64), ("banded_s128", 1024), ("sparse", 16), ("erg_s8_k2", 32), ("banded_s40", 48),
                                        ("erg_s2_k1", 16), ("erg_s32_k3", 128)])
def test_long_sequence_viterbi_is_bit_exact(nat, oracle_lib, name, chunk):
    """HigherOrderHMM.viterbi (:617-708) without the (L+1) x C matrix: rows at chunk boundaries from one sequential sweep,
    decisions recomputed per chunk, chunk entry states by composing per-chunk maps.  Paths and scores equal the oracle's bit
    for bit, for chunk lengths that do and do not divide the sequence lengths, sequences of one chunk, of length 1, unsorted
    child lists (tie rule in child order) and an impossible sequence."""
    hmm = banded_hmm(128, 4, 1) if name == "banded_s128" else MODELS[name]()
    om = oracle_lib.OracleModel(hmm.ir())
    rng = np.random.default_rng(7)
    lens = [5 * chunk + 3, 1, chunk, chunk + 1, 2 * chunk, 7 * chunk - 1, 3, 12 * chunk + chunk // 2]
    data = dataset_of([random_dna(rng, L) for L in lens])
    eng, nd = hmm._engine(), hmm._data(data)
    eng.set_option("vit_chunk", chunk)
    eng.set_option("vit_chunk_force", 1)
    rp, rs, _ = om.viterbi(data.codes, data.offsets)
    for narrow in (False, True):
        paths, scores = eng.viterbi(nd, narrow=narrow)
        assert np.array_equal(scores, rs)
        for i in range(len(lens)):
            assert np.array_equal(paths[data.offsets[i]:data.offsets[i + 1]], rp[i]), f"sequence {i} (L={lens[i]})"
    # the variable-length entry point of the Java shim takes the same route
    p2, s2, plen = eng.viterbi_paths(nd, np.asarray(lens) + 2)
    assert np.array_equal(s2, rs) and np.array_equal(plen, lens)
    for i in range(len(lens)):
        assert np.array_equal(p2[i], rp[i])
    # ... and agrees with the ordinary kernels
    eng.set_option("vit_chunk_force", 0)
    paths0, scores0 = eng.viterbi(nd)
    assert np.array_equal(scores0, rs) and np.array_equal(paths0, paths)


def test_long_sequence_viterbi_impossible_path(nat, oracle_lib):
    hmm = sparse_hmm()
    for em in hmm.emission:
        q = em.params.copy()
        q[:, 0] = -np.inf
        em.setLogProbabilities(q - np.log(np.exp(q).sum(1, keepdims=True)))
    hmm._push_params()
    rng = np.random.default_rng(1)
    ok = rng.integers(1, 4, 200).astype(np.uint8)
    bad = ok.copy()
    bad[137] = 0
    data = dataset_of([ok, bad, ok[:5]])
    eng, nd = hmm._engine(), hmm._data(data)
    eng.set_option("vit_chunk", 16)
    eng.set_option("vit_chunk_force", 1)
    paths, scores = eng.viterbi(nd, allow_impossible=True)
    om = oracle_lib.OracleModel(hmm.ir())
    rp, rs, _ = om.viterbi(np.concatenate([ok, ok[:5]]), np.array([0, 200, 205]))
    assert scores[1] == -np.inf and scores[0] == rs[0] and scores[2] == rs[1]
    assert np.array_equal(paths[:200], rp[0]) and np.array_equal(paths[400:], rp[1]) and (paths[200:400] == -1).all()
    with pytest.raises(nat.NativeError) as ei:
        eng.viterbi(nd)
    assert ei.value.status == -6


@pytest.mark.parametrize("name", ["ring128_0to3", "ring64_0to2", "ring128_pm1_unsorted", "ring256_0to3", "ring128_pm4_gaps"])
def test_band_scoring_meets_in_the_middle(nat, oracle_lib, name):
    """Scoring of long sequences of banded models: forward chain over the first half, backward chain over the second half,
    P(x) = sum_i alpha_mid[i] beta_mid[i] (k_band_pass1 with meet buffers + k_band_meet) — against the oracle and against
    the one-directional pass; sequences below 4096 positions in the same data set keep the forward-only pass."""
    hmm = BAND_MODELS[name]()
    om = oracle_lib.OracleModel(hmm.ir())
    rng = np.random.default_rng(29)
    lens = [4096, 4097, 5000, 12345, 100, 8191, 20000, 1, 4095, 6002]
    data = dataset_of([random_dna(rng, L) for L in lens])
    ll_ref = om.loglik(data.codes, data.offsets)
    got, launches = {}, {}
    for meet in (1, 0):
        eng, nd = hmm._engine(), hmm._data(data)
        eng.set_option("band_meet", meet)
        n0 = nat.lib().jh_kernel_launches()
        got[meet] = eng.loglik(nd)
        launches[meet] = nat.lib().jh_kernel_launches() - n0
        np.testing.assert_allclose(got[meet], ll_ref, rtol=LL_RTOL)
    # two-directional pass + k_band_meet + forward-only pass for the short sequences, against one forward-only pass
    assert launches == {1: 3, 0: 1}
    # log P ~ -5e3 has an ulp of 1e-12: the different association of the sums (1e-15 relative in P) usually vanishes in it
    np.testing.assert_allclose(got[1], got[0], rtol=1e-12)
    assert np.array_equal(got[1][[4, 7, 8]], got[0][[4, 7, 8]])  # the short sequences took the same pass


# ---- banded models: warp-resident whole-sequence passes (kernels_band.cuh) ---------------------------------------------
BAND_MODELS = {
    "ring128_0to3": lambda: banded_hmm(128, 4, 1),                         # cfg5: SPL 4, band [0, 3]
    "ring64_0to2": lambda: banded_hmm(64, 3, 0, ess=2.0),                  # SPL 2, band [0, 2]
    "ring256_0to3": lambda: banded_hmm(256, 4, 1),                         # SPL 8
    "ring128_pm1_unsorted": lambda: banded_hmm(128, 3, 1, offsets=[1, -1, 0]),  # band [-1, 1], children not in state order
    "ring128_pm4_gaps": lambda: banded_hmm(128, 3, 0, offsets=[-4, 2, 4]),      # full window with gaps
    "ring64_skip": lambda: banded_hmm(64, 2, 1, offsets=[2, 0]),            # no edge to the direct neighbour
}


@pytest.mark.parametrize("band", [1, 0], ids=["band", "general"])
@pytest.mark.parametrize("name", list(BAND_MODELS))
def test_banded_models_whole_sequence_passes(nat, oracle_lib, name, band):
    """Scoring, posterior decoding, E-step and chromosome-style Viterbi of banded models with the shuffle exchange
    (band = 1) and with the general shared-memory kernels (band = 0): both against the oracle."""
    hmm = BAND_MODELS[name]()
    om = oracle_lib.OracleModel(hmm.ir())
    rng = np.random.default_rng(3)
    chunk = 48
    lens = [7 * chunk + 5, 1, 2, chunk, chunk + 1, 3 * chunk, 11 * chunk - 1, 9]
    data = dataset_of([random_dna(rng, L) for L in lens])
    eng, nd = hmm._engine(), hmm._data(data)
    for k, v in (("band", band), ("chunk", chunk), ("vit_chunk", chunk), ("vit_chunk_force", 1)):
        eng.set_option(k, v)
    ll_ref = om.loglik(data.codes, data.offsets)
    np.testing.assert_allclose(eng.loglik(nd), ll_ref, rtol=LL_RTOL)
    paths, ll = eng.posterior_decode(nd)
    np.testing.assert_allclose(ll, ll_ref, rtol=LL_RTOL)
    for i in range(len(lens)):
        ref = om.posterior(data.getElementAt(i).codes)
        srt = np.sort(ref, axis=0)
        clear = srt[-1] - srt[-2] > 1e-9
        assert np.array_equal(paths[data.offsets[i]:data.offsets[i + 1]][clear], om.decode_posterior(ref)[clear])
    w = rng.uniform(0.5, 2.0, len(lens))
    for weights in (None, w):
        nd.set_weights(weights)
        ts, es, sc = eng.estep(nd)
        rts, res, rsc = om.estep(data.codes, data.offsets, weights)
        np.testing.assert_allclose(ts, rts, rtol=COUNT_TOL, atol=COUNT_TOL)
        np.testing.assert_allclose(es, res, rtol=COUNT_TOL, atol=COUNT_TOL)
        assert sc == pytest.approx(rsc, rel=LL_RTOL)
    nd.set_weights(None)
    if hmm.getNumberOfStates() <= 255:
        rp, rs, _ = om.viterbi(data.codes, data.offsets)
        vp, vs = eng.viterbi(nd, narrow=True)
        assert np.array_equal(vs, rs)
        for i in range(len(lens)):
            assert np.array_equal(vp[data.offsets[i]:data.offsets[i + 1]], rp[i])
