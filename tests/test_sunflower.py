"""HMMFactory.createSunflowerHMM mirrored in jstacs_b200/factory.py: structure and the propagation of the equivalent sample
size, checked against properties that do not depend on the restatement (mass conservation, the reference's own layout)."""
import numpy as np
import pytest

from jstacs_b200.factory import createSunflowerHMM
from jstacs_b200.hmm import BaumWelchParameterSet, IterationCondition

PARS = BaumWelchParameterSet(1, IterationCondition(2), 1)


@pytest.mark.parametrize("start_central", [True, False])
def test_structure_matches_the_reference_layout(start_central):
    motifs = [3, 1, 4]
    hmm = createSunflowerHMM(PARS, None, 8.0, 50, start_central, motifs)
    assert hmm.getNumberOfStates() == 1 + sum(motifs)
    names = [hmm.getNameOfState(i) for i in range(hmm.getNumberOfStates())] if hasattr(hmm, "getNameOfState") else hmm.name
    assert names[0] == "bg" and names[1] == "motif 0 position 0" and names[-1] == "motif 2 position 3"
    te = hmm.transition.transitions
    assert len(te) == 2 + sum(motifs)                       # start, bg, one element per motif position (HMMFactory.java:335, 376-393)
    assert list(te[0].states) == ([0] if start_central else list(range(9)))
    assert list(te[1].context) == [0] and list(te[1].states) == [0, 1, 4, 5]   # bg -> bg | first state of every motif
    firsts, idx = [1, 4, 5], 1
    for m, n in enumerate(motifs):
        for p in range(n):
            el = te[1 + idx]
            assert list(el.context) == [idx]
            assert list(el.states) == ([0] if p == n - 1 else [idx + 1])        # chain, the last position returns to bg
            idx += 1
    assert firsts == [1, 1 + motifs[0], 1 + motifs[0] + motifs[1]]
    assert all(e.order == 0 for e in hmm.emission)


@pytest.mark.parametrize("start_central", [True, False])
def test_ess_propagation_conserves_mass(start_central):
    ess, L, motifs = 6.0, 40, [2, 5]
    hmm = createSunflowerHMM(PARS, None, ess, L, start_central, motifs, motifProb=[0.05, 0.2])
    # every position carries ess in total, so the state ESS sum to ess * L
    state_ess = np.array([e.hyperParams.sum() for e in hmm.emission])
    assert state_ess.sum() == pytest.approx(ess * L, rel=1e-12)
    bg = hmm.transition.transitions[1]
    # every visit of bg leaves through exactly one of its edges: hyper-parameters of the bg element sum to the bg state's ESS,
    # split self : motif 0 : motif 1 = 0.75 : 0.05 : 0.2
    assert bg.hyperParameters.sum() == pytest.approx(state_ess[0], rel=1e-12)
    np.testing.assert_allclose(bg.hyperParameters / bg.hyperParameters.sum(), [0.75, 0.05, 0.2], rtol=1e-12)
    # inside a motif the mass only moves along the chain: started in the centre, position p+1 has seen what position p saw,
    # one step later (started everywhere, the chains carry their share of the first position as well)
    idx = 1
    for n in motifs:
        chain = state_ess[idx:idx + n]
        if start_central:
            assert np.all(np.diff(chain) <= 1e-12)
        idx += n
    start = hmm.transition.transitions[0]
    assert start.hyperParameters.sum() == pytest.approx(ess, rel=1e-12)


GOLDEN = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "golden", "promoters_head.fa")


def _promoter_model(n_iter):
    from jstacs_b200.data import DataSet

    data = DataSet.from_fasta(GOLDEN)
    pars = BaumWelchParameterSet(1, IterationCondition(n_iter), 1)
    hmm = createSunflowerHMM(pars, None, 4.0, int(np.mean(np.diff(data.offsets))), True, [8, 12])
    rng = np.random.default_rng(17)  # a fixed "random" initialisation: both sides start from the same parameters
    for t in hmm.transition.transitions:
        t.setLogProbabilities(np.log(rng.dirichlet(np.ones(len(t.states)) * 5)))
    for e in hmm.emission:
        e.setLogProbabilities(np.log(rng.dirichlet(np.ones(4) * 5, size=e.params.shape[0])))
    return hmm, data


def test_oracle_trains_the_sunflower_model_on_the_references_promoter_data(oracle_lib):
    """Real cookbook data (64 of the reference's 500 promoters) through the oracle: MAP-EM objectives never decrease."""
    hmm, data = _promoter_model(6)
    assert data.getNumberOfElements() == 64 and data.getNumberOfResidues() > 10000
    om = oracle_lib.OracleModel(hmm.ir())
    obj = om.train(data.codes, data.offsets, max_iter=6)
    assert np.all(np.diff(obj) > -1e-9 * np.abs(obj[:-1]))


@pytest.mark.gpu
def test_gpu_trains_the_sunflower_model_on_promoter_data_like_the_oracle(oracle_lib):
    """HMMFactory.createSunflowerHMM + HigherOrderHMM.train on the promoter data: objectives, trained parameters, Viterbi
    paths (motif occurrences) and posterior decoding against the oracle; the streamed 2-bit route gives the same model."""
    import tempfile

    from jstacs_b200.data import DataSet

    hmm, data = _promoter_model(6)
    om = oracle_lib.OracleModel(hmm.ir())
    robj = om.train(data.codes, data.offsets, max_iter=6)
    hmm.skipInit = True
    obj = hmm.train(data)
    assert obj == pytest.approx(robj[-1], rel=1e-9)
    rtp, rtl, rep, rel_ = om.get_params()
    ir = hmm.ir()
    for got, ref in ((ir.trans_param, rtp), (ir.emis_param, rep)):
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(got), fin)
        np.testing.assert_allclose(np.asarray(got)[fin], ref[fin], rtol=0, atol=1e-6)
    rp, rs, _ = om.viterbi(data.codes, data.offsets)
    eng, nd = hmm._engine(), hmm._data(data)
    vs = eng.viterbi(nd)[1]
    np.testing.assert_allclose(vs, rs, rtol=1e-9)  # the trained parameters agree to ~1e-12, not bit for bit
    np.testing.assert_allclose(eng.loglik(nd), om.loglik(data.codes, data.offsets), rtol=1e-9)
    # the same data streamed FASTA -> 2-bit shards -> device
    hmm2, _ = _promoter_model(6)
    hmm2.skipInit = True
    with tempfile.TemporaryDirectory() as tmp:
        sharded = DataSet.from_fasta_stream(GOLDEN, tmp)
        assert hmm2.train(sharded) == pytest.approx(obj, rel=1e-12)
