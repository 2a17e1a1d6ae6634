"""Shared test helpers: tiny models and an INDEPENDENT brute-force path enumerator (pure Python, walks the
public Transition interface of the host mirror; shares no code with oracle/hmm_oracle.c)."""
from __future__ import annotations

import math

import numpy as np

from jstacs_b200 import (DNA, BaumWelchParameterSet, DataSet, HigherOrderDiscreteEmission, HigherOrderHMM,
                         IterationCondition, Sequence, SilentEmission, TransitionElement)
from jstacs_b200.workloads import ergodic_hmm, set_closed_form_parameters  # noqa: F401

NEG_INF = float("-inf")


def logsumexp(vals):
    vals = [v for v in vals if v > NEG_INF]
    if not vals:
        return NEG_INF
    m = max(vals)
    return m + math.log(sum(math.exp(v - m) for v in vals))


def emission_logp(hmm, state, x, pos):
    """log P(x[pos] | context) of `state`, straight from the parameter tables (ACDE:431-457 semantics)."""
    e = hmm.emission[hmm.emissionIdx[state]]
    if e.order < 0:
        return 0.0
    a = e.params.shape[1]
    if pos < e.order:
        return -math.log(a)
    cond = sum((a ** i) * int(x[pos - 1 - i]) for i in range(e.order))
    return float(e.params[cond, x[pos]] - e.logNorm[cond])


def enumerate_paths(hmm, x):
    """All complete state paths of `hmm` for sequence x with their joint log probability.
    Returns list of (path tuple, logp, edges) where edges = [(layer, ctx, child, state, pos)]."""
    tr = hmm.transition
    L = len(x)
    out = []

    def rec(layer, ctx, logp, path, edges):
        last = tr.getLastContextState(layer, ctx)
        if layer == L and (tr.getMaximalMarkovOrder() == 0 or (last >= 0 and hmm.finalState[last])):  # HOH:498 'order 0' rule
            out.append((tuple(path), logp, list(edges)))
        for ch in range(tr.getNumberOfChildren(layer, ctx)):
            state, nxt, adv = tr.fillTransitionInformation(layer, ctx, ch)
            if adv == 1 and layer >= L:
                continue
            lp = logp + tr.getLogScoreFor(layer, ctx, ch) + emission_logp(hmm, state, x, layer)
            if lp == NEG_INF:
                continue
            path.append(state)
            edges.append((layer, ctx, ch, state, layer))
            rec(layer + adv, nxt, lp, path, edges)
            path.pop()
            edges.pop()

    rec(0, 0, 0.0, [], [])
    return out


def brute_force(hmm, x):
    """(loglik, best score, set of best paths, posterior[S][L], trans counts dict, emis counts dict)."""
    paths = enumerate_paths(hmm, x)
    ll = logsumexp([p[1] for p in paths])
    best = max(p[1] for p in paths)
    S, L = hmm.getNumberOfStates(), len(x)
    post = np.zeros((S, L))
    tcount, ecount = {}, {}
    tr = hmm.transition
    for path, lp, edges in paths:
        w = math.exp(lp - ll)
        for (layer, ctx, ch, state, pos) in edges:
            el = tr.transIndex[tr.getTransitionElementIndex(layer, ctx)]
            tcount[(el, ch)] = tcount.get((el, ch), 0.0) + w
            e = hmm.emission[hmm.emissionIdx[state]]
            if e.order >= 0:
                post[state, pos] += w
                if pos >= e.order:
                    a = e.params.shape[1]
                    cond = sum((a ** i) * int(x[pos - 1 - i]) for i in range(e.order))
                    key = (hmm.emissionIdx[state], cond, int(x[pos]))
                    ecount[key] = ecount.get(key, 0.0) + w
    return ll, best, paths, post, tcount, ecount


def flat_counts(hmm, tcount, ecount):
    ir = hmm.ir()
    ts = np.zeros(ir.n_child_total)
    es = np.zeros(ir.n_cond_total * ir.alphabet)
    for (el, ch), v in tcount.items():
        ts[ir.elem_child_ptr[el] + ch] += v
    for (e, cond, sym), v in ecount.items():
        es[(ir.emis_cond_ptr[e] + cond) * ir.alphabet + sym] += v
    return ts, es


def random_dna(rng, L):
    return rng.integers(0, 4, L).astype(np.uint8)


def dataset_of(seqs):
    return DataSet("t", *[Sequence(DNA, np.asarray(s, dtype=np.uint8)) for s in seqs], con=DNA)


def sparse_hmm(seed=3, ess=0.0):
    """3 states, order-1 emissions, sparse transitions with a non-sorted child order (tie rule / child order tests)."""
    pars = BaumWelchParameterSet(1, IterationCondition(3), 1)
    em = [HigherOrderDiscreteEmission(DNA, ess, 1) for _ in range(3)]
    te = [
        TransitionElement(None, [2, 0], [ess, ess]),
        TransitionElement([0], [1, 0], [ess, ess]),
        TransitionElement([1], [2, 1, 0], [ess, ess, ess]),
        TransitionElement([2], [0, 2], [ess, ess]),
    ]
    hmm = HigherOrderHMM(pars, ["a", "b", "c"], em, *te)
    hmm.setOutputStream(None)
    return set_closed_form_parameters(hmm)


def order0_hmm(ess=0.0, k=1):
    """Transition order 0 (one transition element with the empty context: the states are drawn independently per position,
    HigherOrderHMM.java:295-308, 498), unsorted child order, emission order k."""
    pars = BaumWelchParameterSet(1, IterationCondition(3), 1)
    em = [HigherOrderDiscreteEmission(DNA, ess, k) for _ in range(3)]
    hmm = HigherOrderHMM(pars, ["a", "b", "c"], em, TransitionElement(None, [2, 0, 1], [ess] * 3))
    hmm.setOutputStream(None)
    return set_closed_form_parameters(hmm)


def order2_hmm(S=2, k=1, ess=0.0):
    """Transition order 2 (contexts = last two states), emission order k."""
    return ergodic_hmm(S, k, n_iter=3, ess=ess, transition_order=2)


def silent_hmm(ess=0.0):
    """Small profile-like HMM with silent states: B(silent) -> M1 | D1(silent) ; M1/D1 -> M2 | D2(silent); ... -> E(silent, final).
    States: 0=M1 1=M2 2=D1(silent) 3=D2(silent) 4=I (insert, emitting, self loop) 5=E(silent, absorbing)."""
    pars = BaumWelchParameterSet(1, IterationCondition(3), 1)
    em = [HigherOrderDiscreteEmission(DNA, ess, 0), HigherOrderDiscreteEmission(DNA, ess, 0), SilentEmission(), SilentEmission(),
          HigherOrderDiscreteEmission(DNA, ess, 0), SilentEmission()]
    h = lambda n: [ess] * n
    te = [
        TransitionElement(None, [0, 2], h(2)),      # start -> M1 | D1
        TransitionElement([0], [1, 3, 4], h(3)),    # M1 -> M2 | D2 | I
        TransitionElement([2], [1, 3], h(2)),       # D1 -> M2 | D2
        TransitionElement([4], [4, 1, 3], h(3)),    # I -> I | M2 | D2
        TransitionElement([1], [5, 1], h(2)),       # M2 -> E | M2 (self loop so long sequences are possible)
        TransitionElement([3], [5], h(1)),          # D2 -> E
    ]
    hmm = HigherOrderHMM(pars, ["M1", "M2", "D1", "D2", "I", "E"], em, *te)
    hmm.setOutputStream(None)
    return set_closed_form_parameters(hmm)


def sunflower_hmm(motifs=(2, 3), ess=0.0, start_central=True, seed=5, expected_length=30):
    """HMMFactory.createSunflowerHMM (background + motif chains, no silent states) with random parameters."""
    from jstacs_b200.factory import createSunflowerHMM
    from jstacs_b200.hmm import BaumWelchParameterSet, IterationCondition

    hmm = createSunflowerHMM(BaumWelchParameterSet(1, IterationCondition(4), 1), None, ess, expected_length, start_central, list(motifs))
    rng = np.random.default_rng(seed)
    for t in hmm.transition.transitions:
        t.setLogProbabilities(np.log(rng.dirichlet(np.ones(len(t.states)))))
    for e in hmm.emission:
        e.setLogProbabilities(np.log(rng.dirichlet(np.ones(4), size=e.params.shape[0])))
    return hmm
