"""BASELINE.json configs at their FULL single-GPU sizes: size-independent properties over the whole data set plus oracle
comparisons on samples drawn from it (the oracle finishes samples in seconds, not 1e9 residues).

cfg2: Viterbi, 8 states, order-2 emissions, 1M x 1 kb.   cfg3: Baum-Welch, 32 states, order-3 emissions, 1M x 1 kb per GPU.
cfg5 shape: 128 states, 4 children per state, sequences of many chunks (checkpointed scoring / decoding / E-step).
"""
import numpy as np
import pytest

from jstacs_b200 import DataSet
from jstacs_b200.workloads import banded_hmm, ergodic_hmm, make_data

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-9
COUNT_TOL = 1e-9


@pytest.fixture(scope="module")
def nat():
    import __graft_entry__ as g

    g.build()
    from jstacs_b200 import _native

    if _native.lib().jh_device_count() <= 0:
        pytest.fail("no CUDA device: the product has no CPU fallback")
    return _native


@pytest.fixture(scope="module")
def million():
    """1M x 1 kb on the java.util.Random stream: the data of cfg2 and of one GPU's share of cfg3."""
    return make_data(1_000_000, 1000)


def subset(data, idx):
    off = data.offsets
    codes = np.concatenate([data.codes[off[i]:off[i + 1]] for i in idx])
    lens = np.array([off[i + 1] - off[i] for i in idx], dtype=np.int64)
    o = np.zeros(len(idx) + 1, dtype=np.int64)
    np.cumsum(lens, out=o[1:])
    return DataSet.from_packed(codes, o, data.getAlphabetContainer())


def test_cfg2_viterbi_full_size(nat, oracle_lib, million):
    hmm = ergodic_hmm(8, 2)
    om = oracle_lib.OracleModel(hmm.ir())
    eng, nd = hmm._engine(), hmm._data(million)
    paths, scores = eng.viterbi(nd, narrow=True)  # one byte per residue: 1 GB instead of 4 GB over PCIe
    t_u8 = eng.last_kernel_ms()
    assert paths.size == 1_000_000_000 and paths.max() <= 7
    ll = eng.loglik(nd)
    assert np.isfinite(scores).all() and (scores <= ll + 1e-9).all()  # the best path cannot beat the sum over all paths
    off = million.offsets
    rng = np.random.default_rng(2)
    # the returned path really has the returned score (AbstractHMM.getLogProbForPath), anywhere in the data set
    for i in rng.integers(0, 1_000_000, 1500):
        x, p = million.codes[off[i]:off[i + 1]], paths[off[i]:off[i + 1]]
        assert om.path_logprob(x, p.astype(np.int32)) == pytest.approx(scores[i], rel=1e-12)
    # and it is the reference's path, bit for bit (tie rule included), on a sample
    idx = np.concatenate([[0, 999_999], rng.integers(0, 1_000_000, 126)])
    sub = subset(million, idx)
    rp, rs, _ = om.viterbi(sub.codes, sub.offsets)
    for k, i in enumerate(idx):
        assert np.array_equal(paths[off[i]:off[i + 1]], rp[k]) and scores[i] == rs[k]
    # The general entry point the Java shim binds for silent-state models, jh_viterbi_paths, runs the SAME kernels for a
    # model without silent states (it used to take the thread-per-sequence kernel): device time within
    # 10 % of the byte route, identical paths.
    paths2, scores2, plen = eng.viterbi_paths(nd, np.diff(off))
    t_paths = eng.last_kernel_ms()
    assert (plen == 1000).all() and np.array_equal(scores2, scores)
    for i in rng.integers(0, 1_000_000, 2000):
        assert np.array_equal(paths2[i], paths[off[i]:off[i + 1]])
    assert t_paths <= 1.10 * t_u8 + 1.0, (t_paths, t_u8)
    nd.close()
    million._native = None


def test_cfg3_baum_welch_full_size(nat, oracle_lib, million):
    hmm = ergodic_hmm(32, 3, n_iter=3, ess=4.0)
    om = oracle_lib.OracleModel(hmm.ir())
    eng, nd = hmm._engine(), hmm._data(million)
    ts, es, sc = eng.estep(nd)
    # every position contributes one transition (the start element included) and, from position 3 on, one emission
    assert ts.sum() == pytest.approx(1e9, rel=1e-10)
    assert es.sum() == pytest.approx(1e9 - 3e6, rel=1e-10)
    ll = eng.loglik(nd)
    assert sc == pytest.approx(ll.sum(), rel=1e-11)
    # the start element sees every sequence once; the counts of a transition element sum to the posterior mass of its state
    ir = hmm.ir()
    assert ts[ir.elem_child_ptr[0]:ir.elem_child_ptr[1]].sum() == pytest.approx(1e6, rel=1e-10)
    # linearity: the counts of a subset with weights 2 are twice the counts of the subset; and they match the oracle
    idx = np.random.default_rng(3).integers(0, 1_000_000, 48)
    sub = subset(million, idx)
    snd = hmm._data(sub)
    ts1, es1, sc1 = eng.estep(snd)
    snd.set_weights(np.full(idx.size, 2.0))
    ts2, es2, sc2 = eng.estep(snd)
    np.testing.assert_allclose(ts2, 2 * ts1, rtol=1e-12)
    np.testing.assert_allclose(es2, 2 * es1, rtol=1e-12)
    rts, res, rsc = om.estep(sub.codes, sub.offsets)
    np.testing.assert_allclose(ts1, rts, rtol=COUNT_TOL, atol=COUNT_TOL)
    np.testing.assert_allclose(es1, res, rtol=COUNT_TOL, atol=COUNT_TOL)
    assert sc1 == pytest.approx(rsc, rel=LL_RTOL)
    off = million.offsets
    np.testing.assert_allclose(ll[idx], om.loglik(sub.codes, sub.offsets), rtol=LL_RTOL)
    assert off[-1] == 1_000_000_000
    # three EM iterations on the full data set: the objective does not decrease
    hmm.train(million)
    assert (np.diff(hmm.last_objectives[0]) >= -1e-3).all()
    nd.close()
    million._native = None


def test_cfg5_shape_many_chunks_vs_oracle(nat, oracle_lib):
    """128 states, 4 children per state; one sequence of 73 chunks (150 kb), one of 3 and two single-chunk ones: the
    checkpoint pass (latency mode), chunk-parallel decoding and the chunk-parallel E-step against the oracle's full matrices."""
    hmm = banded_hmm(128, 4, 0)
    om = oracle_lib.OracleModel(hmm.ir())
    data = make_data(4, [150_000, 5000, 2048, 700])
    ll_ref = om.loglik(data.codes, data.offsets)
    np.testing.assert_allclose(hmm.getLogScoreFor(data), ll_ref, rtol=LL_RTOL)
    dec = hmm.decodePosteriorPathsFor(data, narrow=True)
    for i in range(4):
        ref = om.posterior(data.getElementAt(i).codes)
        path, ll = dec[i]
        assert ll == pytest.approx(ll_ref[i], rel=LL_RTOL)
        srt = np.sort(ref, axis=0)
        clear = srt[-1] - srt[-2] > 1e-9
        assert clear.mean() > 0.9 and np.array_equal(path[clear], om.decode_posterior(ref)[clear])
    ts, es, sc = hmm._engine().estep(hmm._data(data))
    rts, res, rsc = om.estep(data.codes, data.offsets)
    # The reference formulation forms every count as exp(fwd + bwd - logP) with |fwd|, |bwd|, |logP| ~ 2e5 here (1 ulp = 3e-11)
    # and its log-space recursions accumulate rounding over 150 000 positions: its own counts carry a relative error of
    # ~1e-8 at this length (observed difference 6.5e-9), the scaled recursion on the GPU ~1e-13.  1e-9 holds at 1 kb (above).
    np.testing.assert_allclose(ts, rts, rtol=5e-8, atol=COUNT_TOL)
    np.testing.assert_allclose(es, res, rtol=5e-8, atol=COUNT_TOL)
    assert ts.sum() == pytest.approx(data.getNumberOfResidues(), rel=1e-11)  # ... while the scaled counts add up to 1e-11
    assert sc == pytest.approx(rsc, rel=LL_RTOL)


def test_cfg4_full_size_ragged(nat, oracle_lib):
    """BASELINE configs[3] at FULL size: 100 000 sequences with log-uniform lengths 100 bp - 100 kb (1.45e9 residues),
    16 states, order-1 emissions.  Scoring and one E-step over the whole data set; oracle parity on a sample of 64 sequences
    that includes the longest one (chunked octets) and the shortest ones; size-independent properties on everything."""
    from jstacs_b200.workloads import ragged_lengths

    lens = ragged_lengths(100_000, seed=42)
    assert lens.min() >= 100 and lens.max() <= 100_000 and 1.3e9 < lens.sum() < 1.6e9
    data = make_data(lens.size, lens)
    hmm = ergodic_hmm(16, 1)
    om = oracle_lib.OracleModel(hmm.ir())
    eng, nd = hmm._engine(), hmm._data(data)
    ll = eng.loglik(nd)
    assert np.isfinite(ll).all() and (ll < 0).all()
    # longer sequences are less likely: the per-residue log-likelihood stays in a narrow band
    per_res = ll / lens
    assert per_res.min() > -1.5 and per_res.max() < -1.2
    ts, es, sc = eng.estep(nd)
    assert sc == pytest.approx(ll.sum(), rel=1e-11)
    assert ts.sum() == pytest.approx(float(lens.sum()), rel=1e-10)            # one transition per residue (start element included)
    assert es.sum() == pytest.approx(float(lens.sum() - lens.size), rel=1e-10)  # one emission count per residue from position 1 on
    order = np.argsort(lens)
    rng = np.random.default_rng(11)
    idx = np.unique(np.concatenate([order[:8], order[-3:], rng.integers(0, lens.size, 53)]))
    sub = subset(data, idx)
    np.testing.assert_allclose(ll[idx], om.loglik(sub.codes, sub.offsets), rtol=LL_RTOL)
    snd = hmm._data(sub)
    w = rng.random(idx.size) + 0.5
    snd.set_weights(w)
    ts1, es1, sc1 = eng.estep(snd)
    rts, res, rsc = om.estep(sub.codes, sub.offsets, w)
    np.testing.assert_allclose(ts1, rts, rtol=2e-8, atol=COUNT_TOL)  # 100 kb sequences: see the note on the reference's own rounding below
    np.testing.assert_allclose(es1, res, rtol=2e-8, atol=COUNT_TOL)
    assert sc1 == pytest.approx(rsc, rel=LL_RTOL)
    # posterior decoding of the sample, byte paths
    paths, dll = eng.posterior_decode(snd, narrow=True)
    np.testing.assert_allclose(dll, ll[idx], rtol=LL_RTOL)
    k = int(np.argmin(np.diff(sub.offsets)))
    ref = om.posterior(sub.getElementAt(k).codes)
    srt = np.sort(ref, axis=0)
    clear = srt[-1] - srt[-2] > 1e-9
    assert np.array_equal(paths[sub.offsets[k]:sub.offsets[k + 1]][clear], om.decode_posterior(ref)[clear])
    nd.close()
    data._native = None


def test_cfg5_shape_multi_megabase_vs_oracle(nat, oracle_lib):
    """cfg5 shape beyond 5 Mb: a 5.2 Mb and a 1.1 Mb sequence, 128 states, banded (the warp-resident band kernels and the
    chunk passes): log-likelihood and E-step against the oracle (which holds two 5.2M x 128 matrices = 10.6 GB here),
    chromosome-scale Viterbi bit for bit on the 1.1 Mb sequence and through the path score on the long one."""
    hmm = banded_hmm(128, 4, 0)
    om = oracle_lib.OracleModel(hmm.ir())
    data = make_data(2, [5_200_000, 1_100_000])
    eng, nd = hmm._engine(), hmm._data(data)
    ll = eng.loglik(nd)
    ll_ref = om.loglik(data.codes, data.offsets)
    np.testing.assert_allclose(ll, ll_ref, rtol=LL_RTOL)
    ts, es, sc = eng.estep(nd)
    assert sc == pytest.approx(ll_ref.sum(), rel=LL_RTOL)
    assert ts.sum() == pytest.approx(data.getNumberOfResidues(), rel=1e-11) and es.sum() == pytest.approx(data.getNumberOfResidues(), rel=1e-11)
    short = subset(data, [1])
    rts, res, rsc = om.estep(short.codes, short.offsets)
    ts1, es1, sc1 = eng.estep(hmm._data(short))
    # the reference's log-space counts lose ~5e-7 at 1.1 Mb (|fwd|, |bwd| ~ 1.5e6: 1 ulp = 2e-10, accumulated over 1.1e6 positions;
    # 6.5e-9 at 150 kb above); the scaled counts on the GPU add up to 1e-11 (checked above)
    np.testing.assert_allclose(ts1, rts, rtol=3e-6, atol=COUNT_TOL)
    np.testing.assert_allclose(es1, res, rtol=3e-6, atol=COUNT_TOL)
    paths, scores = eng.viterbi(nd, narrow=True)
    rp, rs, _ = om.viterbi(short.codes, short.offsets)
    assert scores[1] == rs[0] and np.array_equal(paths[data.offsets[1]:], rp[0])
    x, p = data.codes[:data.offsets[1]], paths[:data.offsets[1]]
    # the 5.2 Mb path really has its score (1e7 terms summed from the other end: rounding ~1e-11 relative)
    assert om.path_logprob(x, p.astype(np.int32)) == pytest.approx(scores[0], rel=1e-9)
    assert scores[0] <= ll[0]
    nd.close()
