"""Profile HMMs (silent delete states, PLAN7) built by the mirror of HMMFactory.createProfileHMM / parseProfileHMMFromHMMer
from a REAL fixture of the reference (projects/xanthogenomes/data/repeats.hmm -> tests/golden/profile_repeats.json):
the oracle and the GPU's general context-graph kernels against log-likelihoods computed independently by dense linear
algebra (silent states eliminated in closed form), and against each other for Viterbi / posteriors / counts / training."""
import json
import os

import numpy as np
import pytest

from helpers import dataset_of
from jstacs_b200 import BaumWelchParameterSet, IterationCondition, ViterbiParameterSet
from jstacs_b200.factory import createProfileHMM, getInitFromTo, parseProfileHMMFromHMMer, propagateESS, PseudoTransitionElement

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "profile_repeats.json")
SRC = "/root/reference/projects/xanthogenomes/data/repeats.hmm"


def fixture_model():
    fx = json.load(open(FIX))
    hmm = createProfileHMM(ViterbiParameterSet(1, IterationCondition(3), 1), "PLAN7", 1, fx["n_nodes"] + 2, None, 16.0)
    hmm.setParameters(np.array([float.fromhex(v) for v in fx["params"]]), 0)
    hmm.setSkiptInit(True)
    seqs = [np.array(["ACGT".index(c) for c in s], dtype=np.uint8) for s in fx["sequences"]]
    return fx, hmm, seqs, np.array([float.fromhex(v) for v in fx["loglik_independent"]])


def test_profile_structure_is_the_reference_architecture():
    hmm = createProfileHMM(ViterbiParameterSet(1, IterationCondition(3), 1), "PLAN7", 1, 4, None, 16.0)
    # HMMFactory.createProfileHMM: end chain, start chain, D0, I0, [D, I, M] per layer, last layer without I, D(n+1), F
    assert hmm.name == ["E0", "S0", "D0", "I0", "D1", "I1", "M1", "D2", "I2", "M2", "D3", "I3", "M3", "D4", "M4", "D5", "F"]
    silent = [hmm.emission[i].order < 0 for i in hmm.emissionIdx]
    assert [n for n, s in zip(hmm.name, silent) if s] == ["E0", "S0", "D0", "D1", "D2", "D3", "D4", "D5", "F"]
    assert [n for n, f in zip(hmm.name, hmm.finalState) if f] == ["F"]  # the only absorbing state (AbstractHMM.java:1101-1113)
    T = {tuple(hmm.name[c] for c in te.context): [hmm.name[s] for s in te.states] for te in hmm.transition.transitions}
    assert T[()] == ["S0"] and T[("S0",)] == ["D0", "I0"] and T[("D0",)] == ["D1", "M1"] and T[("I0",)] == ["I0", "M1"]
    assert T[("M1",)] == ["I1", "D2", "M2"] and T[("I1",)] == ["I1", "M2"] and T[("D1",)] == ["D2", "M2"]
    assert T[("M3",)] == ["I3", "D4", "M4"] and T[("M4",)] == ["D5"] and T[("D4",)] == ["D5"] and T[("D5",)] == ["E0"] and T[("E0",)] == ["F"]
    # the equivalent sample size is conserved along the chain (propagateESS): what enters a state leaves it
    for te in hmm.transition.transitions:
        if te.states:
            assert te.hyperParameters.sum() > 0
    assert hmm.transition.transitions[0].hyperParameters.tolist() == [16.0]
    ess_in = {n: 0.0 for n in hmm.name}
    for te in hmm.transition.transitions:
        for s, h in zip(te.states, te.hyperParameters):
            ess_in[hmm.name[s]] += h
    assert ess_in["F"] == pytest.approx(16.0, rel=1e-9) and ess_in["E0"] == pytest.approx(16.0, rel=1e-9)
    # higher transition order builds too (context graph with silent states of order 2)
    h2 = createProfileHMM(ViterbiParameterSet(1, IterationCondition(3), 1), "PLAN9", 2, 3, None, 8.0)
    assert h2.transition.getMaximalMarkovOrder() == 2 and h2.name[:4] == ["E0", "E1", "S0", "S1"]


def test_propagate_ess_small_chain():
    lst = [PseudoTransitionElement([], [0], [1.0]), PseudoTransitionElement([0], [0, 1], [3.0, 1.0])]
    hyper, state_ess = propagateESS(8.0, lst)
    # state 0 is visited 1/(1 - 3/4) = 4 times per unit of mass: 8 * 4 = 32 in total, 24 back to itself, 8 on to state 1
    assert hyper[0] == [8.0] and hyper[1] == pytest.approx([24.0, 8.0], rel=1e-9) and state_ess[0] == pytest.approx(32.0, rel=1e-9)
    assert np.isnan(getInitFromTo("PLAN7", 6.0)[0][0]) and getInitFromTo("PLAN7", 6.0)[2][1] == 2.0


@pytest.mark.skipif(not os.path.exists(SRC), reason="reference tree not present")
def test_fixture_is_what_the_hmmer_file_parses_to():
    fx, hmm, _, _ = fixture_model()
    consensus = []
    ref, bg = parseProfileHMMFromHMMer(open(SRC), consensus)
    assert "".join(consensus) == fx["consensus"] and len(ref.name) == fx["n_states"] == 317
    assert np.array_equal(ref.getCurrentParameterValues(), hmm.getCurrentParameterValues())
    # 4 * (1 + 2 * 102 + 3) emission and 6 + 7 * 102 + 7 transition parameters: exactly the model's free parameters
    assert ref.getNumberOfParameters() == 4 * (1 + 2 * 102 + 3) + 6 + 7 * 102 + 7
    assert np.allclose(np.exp(bg).sum(), 1.0, atol=1e-4)


def test_oracle_matches_the_independent_likelihoods(oracle_lib):
    fx, hmm, seqs, ll = fixture_model()
    data = dataset_of(seqs)
    om = oracle_lib.OracleModel(hmm.ir())
    got = om.loglik(data.codes, data.offsets)
    fin = np.isfinite(ll)
    assert np.array_equal(np.isfinite(got), fin) and fin.sum() >= 10
    np.testing.assert_allclose(got[fin], ll[fin], rtol=1e-10)
    # the consensus is the most likely of the sequences derived from it
    assert got[0] == got[:8].max()


@pytest.mark.gpu
def test_profile_hmm_on_the_gpu(oracle_lib):
    import __graft_entry__ as g

    g.build()
    fx, hmm, seqs, ll = fixture_model()
    keep = [s for s, v in zip(seqs, ll) if np.isfinite(v)]
    ll = ll[np.isfinite(ll)]
    data = dataset_of(keep)
    om = oracle_lib.OracleModel(hmm.ir())
    np.testing.assert_allclose(hmm.getLogScoreFor(data), ll, rtol=1e-9)  # against the INDEPENDENT values
    # Viterbi with silent states: variable-length paths, bit-exact against the oracle
    rp, rs, _ = om.viterbi(data.codes, data.offsets)
    for (p, s), p_ref, s_ref in zip(hmm.getViterbiPathsFor(data), rp, rs):
        assert np.array_equal(p, p_ref) and s == s_ref
    names = [hmm.name[i] for i in hmm.getViterbiPathsFor(dataset_of(keep[:1]))[0][0]]
    assert names[0] == "S0" and names[-1] == "F" and sum(n.startswith("M") for n in names) >= 90  # the consensus runs through the match states
    # state posteriors (silent states -inf) and the Baum-Welch E-step
    post = hmm.getLogStatePosteriorMatrixFor(data.getElementAt(0))
    ref = om.posterior(data.getElementAt(0).codes)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(post), fin)
    np.testing.assert_allclose(post[fin], ref[fin], rtol=0, atol=1e-9)
    ts, es, sc = hmm._engine().estep(hmm._data(data))
    rts, res, rsc = om.estep(data.codes, data.offsets)
    np.testing.assert_allclose(ts, rts, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(es, res, rtol=1e-9, atol=1e-9)
    assert sc == pytest.approx(rsc, rel=1e-9) and sc == pytest.approx(ll.sum(), rel=1e-9)
    # Baum-Welch training of the profile
    hmm.trainingParameter = BaumWelchParameterSet(1, IterationCondition(3), 1)
    hmm.train(data)
    np.testing.assert_allclose(hmm.last_objectives[0], om.train(data.codes, data.offsets, max_iter=3), rtol=1e-9)
