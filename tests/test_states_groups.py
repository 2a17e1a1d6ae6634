"""Allowed states groups (semi-supervised training / constrained scoring and decoding): HigherOrderHMM.getAllowedContext
(HigherOrderHMM.java:331-340), fillPreComputedContext (AbstractHMM.java:173-228), doOneStep (:812-835).

The oracle's restatement is pinned by brute-force enumeration of all state paths under the constraint; the CUDA path
(reference-order kernels) is compared with the oracle."""
import itertools

import numpy as np
import pytest

from helpers import dataset_of, emission_logp, logsumexp, order2_hmm, random_dna, sparse_hmm
from jstacs_b200.workloads import ergodic_hmm

NEG_INF = float("-inf")


def brute_force(hmm, x, groups_of_states, grp):
    """log P(x, constraint), best constrained path and its score by enumerating every state path.
    Constraint (getAllowedContext): for layers >= m the context's last state — the state that emitted position p — has to
    belong to group grp[p] (if grp[p] >= 0); contexts shorter than m (the first m-1 positions) are unconstrained."""
    S, L = hmm.getNumberOfStates(), len(x)
    T = hmm.transition.transitions
    m = hmm.transition.getMaximalMarkovOrder()
    by_ctx = {tuple(te.context): te for te in T}
    total, best, best_path = [], NEG_INF, None
    for path in itertools.product(range(S), repeat=L):
        lp, ctx, ok = 0.0, (), True
        for p, s in enumerate(path):
            te = by_ctx.get(ctx)
            if te is None or s not in te.states:
                ok = False
                break
            if p + 1 >= m and grp[p] >= 0 and s not in groups_of_states[grp[p]]:
                ok = False
                break
            lp += float(te.parameters[te.states.index(s)] - te.logNorm) + emission_logp(hmm, s, x, p)
            ctx = (ctx + (s,))[-m:]
        if not ok or not hmm.finalState[path[-1]]:
            continue
        total.append(lp)
        if lp > best:
            best, best_path = lp, path
    return logsumexp(total), best, best_path


MODELS = {"erg_s3_k1": lambda: ergodic_hmm(3, 1, ess=1.0), "sparse": sparse_hmm, "order2_s2": lambda: order2_hmm(2, 1)}
GROUPS = [[0], [1, 2], [0, 2]]


@pytest.mark.parametrize("name", list(MODELS))
def test_oracle_constrained_scores_match_brute_force(oracle_lib, name):
    hmm = MODELS[name]()
    S = hmm.getNumberOfStates()
    groups = [[s for s in g if s < S] for g in GROUPS]
    om = oracle_lib.OracleModel(hmm.ir())
    om.set_groups(groups)
    rng = np.random.default_rng(4)
    for trial in range(6):
        L = int(rng.integers(1, 7))
        x = random_dna(rng, L)
        grp = rng.integers(-1, len(groups), L).astype(np.int32)
        data = dataset_of([x])
        om.attach_groups(data.codes, grp)
        ll = om.loglik(data.codes, data.offsets)[0]
        paths, scores, _ = om.viterbi(data.codes, data.offsets)
        ref_ll, ref_best, ref_path = brute_force(hmm, x, groups, grp)
        if ref_ll == NEG_INF:
            assert ll == NEG_INF
            continue
        assert ll == pytest.approx(ref_ll, rel=1e-12)
        assert scores[0] == pytest.approx(ref_best, rel=1e-12)
        got = tuple(int(v) for v in paths[0])
        # the returned path obeys the constraint and has the best score (ties: any maximiser)
        assert all(grp[p] < 0 or p + 1 < hmm.transition.getMaximalMarkovOrder() or got[p] in groups[grp[p]] for p in range(L))
        # unconstrained positions leave the result of the unconstrained model
        om.attach_groups(None, None)
        assert om.loglik(data.codes, data.offsets)[0] >= ll - 1e-12
    om.attach_groups(None, None)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["erg_s3_k1", "sparse", "order2_s2", "erg_s8_k2", "erg_s32_k3"])
def test_gpu_constrained_scoring_decoding_training(oracle_lib, name):
    import __graft_entry__ as g

    g.build()
    from jstacs_b200 import _native

    hmm = MODELS[name]() if name in MODELS else (ergodic_hmm(8, 2, ess=2.0) if name == "erg_s8_k2" else ergodic_hmm(32, 3, ess=4.0))
    S = hmm.getNumberOfStates()
    groups = [list(range(0, S, 2)), list(range(1, S, 2)), [0, S - 1], list(range(S // 2))]
    rng = np.random.default_rng(8)
    data = dataset_of([random_dna(rng, int(L)) for L in rng.integers(1, 90, 30)])
    grp = rng.integers(-1, len(groups), data.codes.size).astype(np.int32)
    grp[rng.random(grp.size) < 0.6] = -1  # most positions unconstrained, as in partially labelled data
    w = rng.random(data.getNumberOfElements()) + 0.5
    om = oracle_lib.OracleModel(hmm.ir())
    om.set_groups(groups)
    om.attach_groups(data.codes, grp)
    eng = hmm._engine()
    eng.set_states_groups(groups)
    nd = _native.NativeData(data.codes, data.offsets, None, 4)
    nd.set_groups(grp)
    for fast in (1, 0):  # the constraint is evaluated by the reference-order kernels whatever the setting
        eng.set_option("fast", fast)
        ref = om.loglik(data.codes, data.offsets)
        got = eng.loglik(nd)
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(got), fin)
        np.testing.assert_allclose(got[fin], ref[fin], rtol=1e-9)
        rp, rs, _ = om.viterbi(data.codes, data.offsets)
        paths, scores = eng.viterbi(nd, allow_impossible=True)
        assert np.array_equal(scores, rs)
        for i in np.flatnonzero(np.isfinite(rs)):
            assert np.array_equal(paths[data.offsets[i]:data.offsets[i + 1]], rp[i])
    # training: only sequences that are possible under their constraints can be used (the reference fails on the others)
    keep = np.flatnonzero(fin)
    sub = dataset_of([data.getElementAt(int(i)).codes for i in keep])
    gsub = np.concatenate([grp[data.offsets[i]:data.offsets[i + 1]] for i in keep])
    om.attach_groups(sub.codes, gsub)
    nd2 = _native.NativeData(sub.codes, sub.offsets, w[keep], 4)
    nd2.set_groups(gsub)
    ts, es, sc = eng.estep(nd2)
    rts, res, rsc = om.estep(sub.codes, sub.offsets, w[keep])
    np.testing.assert_allclose(ts, rts, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(es, res, rtol=1e-9, atol=1e-9)
    assert sc == pytest.approx(rsc, rel=1e-9)
    obj = eng.baum_welch(nd2, _native.MODE_BAUM_WELCH, _native.TERM_ITERATIONS, 4, 0.0)
    np.testing.assert_allclose(obj, om.train(sub.codes, sub.offsets, w[keep], max_iter=4), rtol=1e-9)
    # detaching the groups gives the unconstrained results again
    nd.set_groups(None)
    om.attach_groups(None, None)
    np.testing.assert_allclose(eng.loglik(nd), om.loglik(data.codes, data.offsets), rtol=1e-9)
