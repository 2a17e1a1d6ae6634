"""A second, independent route to the flattened model.

`flatten()` (jstacs_b200/hmm.py) reads the context tables that BasicHigherOrderTransition._init — a restatement of the
reference's init — produces; the CUDA path, the oracle and the brute-force enumerator all consume them, so a mis-ordered
or mis-linked context table could pass everywhere.  Here the context graph is derived FROM FIRST PRINCIPLES from the
user-level description only (the elements' context / states lists): contexts of a layer = the reachable tuples of the last
min(layer, m) states, found by breadth-first search from the start; element of a context = the element declared for that
tuple; the contexts are numbered in sorted order — nothing of _init's lookup tables, ordering or descendant links is used.
Both descriptions must define the same model: identical Viterbi paths and scores, likelihoods and counts equal up to the
summation order."""
import numpy as np
import pytest

from helpers import dataset_of, order2_hmm, random_dna, sparse_hmm, sunflower_hmm
from jstacs_b200.hmm import ModelIR
from jstacs_b200.workloads import banded_hmm, ergodic_hmm


def independent_ir(hmm) -> ModelIR:
    T = hmm.transition.transitions
    by_ctx = {}
    for e, te in enumerate(T):
        key = tuple(te.context)
        assert key not in by_ctx, "two elements for one context"
        by_ctx[key] = e
    m = max(len(te.context) for te in T)
    assert all(e.order >= 0 for e in hmm.emission), "models without silent states only"
    layers = [[()]]
    for lc in range(1, m + 1):  # contexts reachable after lc emissions (tuples of the last min(lc, m) states)
        nxt = set()
        for c in layers[-1]:
            for s in T[by_ctx[c]].states:
                nxt.add((c + (s,))[-m:])
        layers.append(sorted(nxt))
    # class m must be closed under the transitions
    frontier, seen = list(layers[m]), set(layers[m])
    while frontier:
        c = frontier.pop()
        for s in T[by_ctx[c]].states:
            n = (c + (s,))[-m:]
            if n not in seen:
                seen.add(n)
                frontier.append(n)
    layers[m] = sorted(seen)
    ir = ModelIR()
    ir.n_states, ir.alphabet, ir.max_order = len(hmm.name), hmm.alphabets.getAlphabetLengthAt(0), m
    ir.n_elem, ir.n_emis = len(T), len(hmm.emission)
    ptr, ctx_elem, ctx_last = [0], [], []
    for lc in range(m + 1):
        for c in layers[lc]:
            ctx_elem.append(by_ctx[c])
            ctx_last.append(c[-1] if c else -1)
        ptr.append(len(ctx_elem))
    ir.lc_ctx_ptr, ir.ctx_elem, ir.ctx_last_state = np.array(ptr, np.int32), np.array(ctx_elem, np.int32), np.array(ctx_last, np.int32)
    cptr = [0]
    for te in T:
        cptr.append(cptr[-1] + len(te.states))
    ir.elem_child_ptr = np.array(cptr, np.int32)
    ir.child_state = np.array([s for te in T for s in te.states], np.int32)
    ir.child_desc = np.array([by_ctx[(tuple(te.context) + (s,))[-m:]] for te in T for s in te.states], np.int32)
    elem_ctx = np.full((m + 1, len(T)), -1, np.int32)
    for lc in range(m + 1):
        for i, c in enumerate(layers[lc]):
            elem_ctx[lc, by_ctx[c]] = i
    ir.elem_ctx = elem_ctx.reshape(-1)
    ir.trans_index = np.array(hmm.transition.transIndex, np.int32)
    ir.trans_param = np.array([v for te in T for v in te.parameters], np.float64)
    ir.trans_lognorm = np.array([te.logNorm for te in T], np.float64)
    ir.trans_hyper = np.array([v for te in T for v in te.hyperParameters], np.float64)
    ir.state_emis = np.array(hmm.emissionIdx, np.int32)
    ir.emis_order = np.array([e.order for e in hmm.emission], np.int32)
    rows = [0]
    for e in hmm.emission:
        rows.append(rows[-1] + e.params.shape[0])
    ir.emis_cond_ptr = np.array(rows, np.int32)
    ir.emis_param = np.concatenate([e.params.reshape(-1) for e in hmm.emission])
    ir.emis_lognorm = np.concatenate([e.logNorm for e in hmm.emission])
    ir.emis_hyper = np.concatenate([e.hyperParams.reshape(-1) for e in hmm.emission])
    ir.state_silent = np.zeros(len(hmm.name), np.uint8)
    # final states: the absorbing ones, or every state if there is none (AbstractHMM.java:1101-1113), from the element lists
    has_out = {te.context[-1] for te in T if te.context and te.states}
    absorbing = [s not in has_out for s in range(len(hmm.name))]
    ir.state_final = np.array(absorbing if any(absorbing) else [True] * len(hmm.name), np.uint8)
    return ir


MODELS = {
    "erg_s3_k1": lambda: ergodic_hmm(3, 1, ess=2.0),
    "erg_s8_k2": lambda: ergodic_hmm(8, 2),
    "erg_s32_k3": lambda: ergodic_hmm(32, 3, ess=4.0),
    "sparse_unsorted_children": sparse_hmm,
    "transition_order_2": lambda: order2_hmm(3, 1),
    "transition_order_3": lambda: ergodic_hmm(2, 1, transition_order=3),
    "banded_ring_40": lambda: banded_hmm(40, 4, 1),
    "ring_unsorted_offsets": lambda: banded_hmm(64, 3, 0, offsets=[1, -1, 0]),
    "sunflower": lambda: sunflower_hmm((4, 7), ess=1.0, start_central=False),
}


def data_for(seed=9):
    rng = np.random.default_rng(seed)
    return dataset_of([random_dna(rng, int(L)) for L in [1, 2, 3, 5, 9, 33, 64, 120]])


@pytest.mark.parametrize("name", list(MODELS))
def test_flatten_and_first_principles_define_the_same_model(oracle_lib, name):
    hmm = MODELS[name]()
    a, b = hmm.ir(), independent_ir(hmm)
    # same contexts per layer class (as sets of (element, last state)), possibly in another order
    for lc in range(a.max_order + 1):
        sa = sorted(zip(a.ctx_elem[a.lc_ctx_ptr[lc]:a.lc_ctx_ptr[lc + 1]].tolist(), a.ctx_last_state[a.lc_ctx_ptr[lc]:a.lc_ctx_ptr[lc + 1]].tolist()))
        sb = sorted(zip(b.ctx_elem[b.lc_ctx_ptr[lc]:b.lc_ctx_ptr[lc + 1]].tolist(), b.ctx_last_state[b.lc_ctx_ptr[lc]:b.lc_ctx_ptr[lc + 1]].tolist()))
        assert sa == sb, f"layer class {lc}"
    assert np.array_equal(a.child_desc, b.child_desc) and np.array_equal(a.state_final, b.state_final)
    data = data_for()
    oa, ob = oracle_lib.OracleModel(a), oracle_lib.OracleModel(b)
    np.testing.assert_allclose(oa.loglik(data.codes, data.offsets), ob.loglik(data.codes, data.offsets), rtol=1e-12)
    pa, sa_, _ = oa.viterbi(data.codes, data.offsets)
    pb, sb_, _ = ob.viterbi(data.codes, data.offsets)
    assert np.array_equal(sa_, sb_) and all(np.array_equal(x, y) for x, y in zip(pa, pb))
    ta, ea, ca = oa.estep(data.codes, data.offsets)
    tb, eb, cb = ob.estep(data.codes, data.offsets)
    np.testing.assert_allclose(ta, tb, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(ea, eb, rtol=1e-10, atol=1e-12)
    assert ca == pytest.approx(cb, rel=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(MODELS))
def test_gpu_accepts_the_first_principles_ir(oracle_lib, name):
    import __graft_entry__ as g

    g.build()
    from jstacs_b200 import _native

    hmm = MODELS[name]()
    data = data_for(10)
    ref = oracle_lib.OracleModel(hmm.ir())
    eng = _native.NativeModel(independent_ir(hmm))
    nd = _native.NativeData(data.codes, data.offsets, None, 4)
    np.testing.assert_allclose(eng.loglik(nd), ref.loglik(data.codes, data.offsets), rtol=1e-9)
    paths, scores = eng.viterbi(nd)
    rp, rs, _ = ref.viterbi(data.codes, data.offsets)
    assert np.array_equal(scores, rs) and np.array_equal(paths, np.concatenate(rp))
    ts, es, sc = eng.estep(nd)
    rts, res, rsc = ref.estep(data.codes, data.offsets)
    np.testing.assert_allclose(ts, rts, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(es, res, rtol=1e-9, atol=1e-9)
