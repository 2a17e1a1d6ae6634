"""Profile-HMM builders of the reference, restated for the Python host side.

WHAT THIS FILE IS.  A statement-level Python restatement (not from-scratch product code) of the host-side MODEL BUILDERS
  HMM/HMMFactory.java:601-760   createProfileHMM (silent start/end chains, D/I/M layers, end state; no joining states)
  HMM/HMMFactory.java:844-1060  createProfileHMMCore, addProfileTransitions, addConditional, getInitFromTo
  HMM/HMMFactory.java:1113-1262 createTransition, propagateESS, PseudoTransitionElement
  HMM/HMMFactory.java:419-582   parseProfileHMMFromHMMer (HMMER3 text -> parameters of a PLAN7 profile HMM)
(HMM/ = de/jstacs/sequenceScores/statisticalModels/trainable/hmm/).  State order, transition-element order and parameter
order have to be the reference's so that parameters parsed from a HMMER file land in the right slots
(DifferentiableHigherOrderHMM.setParameters, :339-344) — hence a restatement.  Nothing here is on the hot path: the
builders produce the same HigherOrderHMM objects every other model goes through (context graph -> jh_model_desc ->
kernels_general.cuh for models with silent states).  Also restated: HMM/HMMFactory.java:270-400 createSunflowerHMM (a
first-order model WITHOUT silent states: background state + motif chains, used by DeNovoSunflower).  What is NOT mirrored:
joining states (closeCircle), reference-sequence / mismatch / phylogenetic emissions, uniform insert emissions.
"""
from __future__ import annotations

import math

import numpy as np

from .data import DNA, AlphabetContainer
from .hmm import (DifferentiableHigherOrderHMM, DiscreteEmission, IterationCondition, SilentEmission, TransitionElement,
                  ViterbiParameterSet)

NAN = float("nan")


def getInitFromTo(hmm_type: str, ess: float):
    """HMMFactory.getInitFromTo (:765-791): rows D, I, M of the current layer; columns D, I, M of the same layer, then of the
    next layer; NaN = no such transition."""
    e2, e3 = ess / 2.0, ess / 3.0
    if hmm_type == "PLAN9":
        row = [NAN, e3, NAN, e3, NAN, e3]
        return [list(row), list(row), list(row)]
    if hmm_type == "PLAN7":
        return [[NAN, NAN, NAN, e2, NAN, e2], [NAN, e2, NAN, NAN, NAN, e2], [NAN, e3, NAN, e3, NAN, e3]]
    if hmm_type == "PLAN8I":
        return [[NAN, e3, NAN, e3, NAN, e3], [NAN, e2, NAN, NAN, NAN, e2], [NAN, e3, NAN, e3, NAN, e3]]
    if hmm_type == "PLAN8D":
        return [[NAN, NAN, NAN, e2, NAN, e2], [NAN, e3, NAN, e3, NAN, e3], [NAN, e3, NAN, e3, NAN, e3]]
    raise ValueError("unknown HMMType")


class PseudoTransitionElement:
    """HMMFactory.PseudoTransitionElement (:1264-1318)."""

    def __init__(self, context, states, posScore=None, weights=None):
        self.context = [] if context is None else list(context)
        self.states = [] if states is None else list(states)
        n = len(self.states)
        if posScore is None:
            self.prob = [1.0 / n] * n if n else []
        else:
            s = posScore[0]
            for v in posScore[1:]:
                s += v
            self.prob = [v / s for v in posScore]  # Normalisation.sumNormalisation
        self.weights = [1.0] * n if weights is None else list(weights)
        self.child = [-1] * n


def _add_conditional(context, lst) -> bool:
    """HMMFactory.addConditional (:1021-1033)."""
    for cur in lst:
        if list(cur) == list(context):
            return False
    lst.append(list(context))
    return True


def _add_profile_transitions(initFromTo, order, states, allTE, lastLayer, newLayer):
    """HMMFactory.addProfileTransitions (:973-1018): lastLayer grows while it is traversed (same-layer children)."""
    o = order - 1
    s = 0
    while s < len(lastLayer):
        current = lastLayer[s]
        k = 0
        while current[o] != states[k]:
            k += 1
        hyper, weights, children = [], [], []
        a = 0
        for j in range(len(initFromTo[k])):
            if not math.isnan(initFromTo[k][j]) and states[j] >= 0:
                hyper.append(initFromTo[k][j])
                children.append(states[j])
                if j < 3:
                    weights.append(1000.0)
                    a = len(children)
                else:
                    weights.append(1.0)
        if children:
            allTE.append(PseudoTransitionElement(current, children, hyper, weights))
        nxt = list(current)
        nxt[:-1] = nxt[1:]  # shiftContext
        for j, ch in enumerate(children):
            nxt[o] = ch
            _add_conditional(nxt, lastLayer if j < a else newLayer)
        s += 1
    lastLayer[:] = newLayer
    del newLayer[:]


def _shift(states, s):
    states[:len(states) - s] = states[s:]


def propagateESS(ess: float, lst):
    """HMMFactory.propagateESS (:1130-1262) -> (hyper-parameters per element, ess per state)."""
    maxOrder, mx, startIdx = 0, 0, -1
    for i, te in enumerate(lst):
        for st in te.states:
            mx = max(mx, st)
        if len(te.context) == 0:
            if startIdx >= 0:
                raise ValueError("Multiple start transitions!")
            startIdx = i
        else:
            maxOrder = max(maxOrder, len(te.context))
    end = []
    t = 0
    while t < len(lst):
        te = lst[t]
        if len(te.states) == 0:
            end.append(t)
        for child in range(len(te.states)):
            n = min(maxOrder, len(te.context) + 1)
            nextContext = [0] * n
            nextContext[n - 1] = te.states[child]
            k = 1
            for j in range(n - 2, -1, -1):
                nextContext[j] = te.context[len(te.context) - k]
                k += 1
            for s2, te2 in enumerate(lst):
                if len(te2.context) == n and te2.context == nextContext:
                    te.child[child] = s2
                    break
            if te.child[child] == -1:
                if maxOrder == 0:
                    te.child[child] = 0
                else:
                    te.child[child] = len(lst)
                    lst.append(PseudoTransitionElement(nextContext, None, None))
        t += 1
    if not end:
        raise ValueError("No absorbing state!")
    n = len(lst)
    cumulated, current, nxt = [0.0] * n, [0.0] * n, [0.0] * n
    current[startIdx] = ess
    while True:
        for t, te in enumerate(lst):
            for child in range(len(te.states)):
                nxt[te.child[child]] += te.prob[child] * current[t]
        total = 0.0
        for i in range(n):
            cumulated[i] += current[i]
            current[i] = nxt[i]
            nxt[i] = 0.0
        for v in current:
            total += v
        if not total > 1e-12:
            break
    hyper = [[cumulated[i] * p for p in te.prob] for i, te in enumerate(lst)]
    stateESS = [0.0] * (mx + 1)
    for i, te in enumerate(lst):
        if te.context:
            stateESS[te.context[-1]] += cumulated[i]
    return hyper, stateESS


def createProfileHMM(trainingParameterSet, hmm_type_or_initFromTo, order: int, numLayers: int, con: AlphabetContainer | None, ess: float):
    """HMMFactory.createProfileHMM (:601-760) with MState.UNCONDITIONAL (DiscreteEmission match states), discrete insert
    states and without joining states: states E0..E(order-1), S0.., D0, I0, [D, I, M] per layer (the last layer without I),
    D(numLayers+1), F.  Returns a DifferentiableHigherOrderHMM over TransitionElements, as HMMFactory.getHMM (:103-119)."""
    if order < 1:
        raise ValueError("The order of a profile HMM has to be at least 1.")
    con = con or DNA
    initFromTo = getInitFromTo(hmm_type_or_initFromTo, ess) if isinstance(hmm_type_or_initFromTo, str) else hmm_type_or_initFromTo
    em, names, lst = [], [], []  # em: "S" silent, "I" insert (DiscreteEmission), "M" match (DiscreteEmission)
    endContext = []
    for i in range(order):
        names.append(f"E{i}")
        em.append("S")
        endContext.append(i)
    for i in range(order):
        names.append(f"S{i}")
        em.append("S")
        lst.append(PseudoTransitionElement([order + k for k in range(i)], [order + i], [ess]))
    lastLayer = []
    lastStartNodeIndex = len(em) - 1
    offset = len(em)
    # --- createProfileHMMCore (:844-917)
    newLayer = []
    em.append("S"); names.append("D0")
    em.append("I"); names.append("I0")
    context = [lastStartNodeIndex - order + i + 1 for i in range(order)]
    if math.isnan(initFromTo[0][1]):
        lst.append(PseudoTransitionElement(context, [len(em) - 2, len(em) - 1], [0.5, 0.5]))
    else:
        lst.append(PseudoTransitionElement(context, [len(em) - 2], [1.0]))
    del lastLayer[:]
    context = list(context)
    context[:-1] = context[1:]
    context[order - 1] = len(em) - 2
    lastLayer.append(list(context))
    if math.isnan(initFromTo[0][1]):
        context[order - 1] = len(em) - 1
        lastLayer.append(list(context))
    states = [-1, -1, -1, offset, offset + 1, -1]
    create = [True, True, True]
    for i in range(numLayers):
        if i == numLayers - 1:
            create = [not math.isnan(initFromTo[j][3]) for j in range(3)]
        last = len(em)
        for j, (kind, letter) in enumerate((("S", "D"), ("I", "I"), ("M", "M"))):  # createEmissions (:1070-1102)
            if create[j]:
                em.append(kind)
                names.append(f"{letter}{i + 1}")
        _shift(states, 3)
        for j in range(3, 6):
            if create[j - 3]:
                states[j] = last
                last += 1
            else:
                states[j] = -1
        _add_profile_transitions(initFromTo, order, states, lst, lastLayer, newLayer)
    em.append("S"); names.append(f"D{numLayers + 1}")
    _shift(states, 3)
    states[3], states[4], states[5] = len(em) - 1, -1, -1
    _add_profile_transitions(initFromTo, order, states, lst, lastLayer, newLayer)
    # --- connect with the end chain (:683-694)
    states = [-1] * 6
    states[3] = len(em) - 1
    for i in range(order):
        _shift(states, 3)
        states[3] = i
        _add_profile_transitions(initFromTo, order, states, lst, lastLayer, newLayer)
    names.append("F")
    em.append("S")
    lst.append(PseudoTransitionElement(endContext, [len(em) - 1], [ess]))
    hyper, stateESS = propagateESS(ess, lst)
    emissions = []
    for i, kind in enumerate(em):  # getEmissions (:793-842): DiscreteEmission(con, ess of the state)
        emissions.append(SilentEmission(con) if kind == "S" else DiscreteEmission(con, float(stateESS[i]) if i < len(stateESS) else 0.0))
    te = [TransitionElement(p.context, p.states, hyper[i]) for i, p in enumerate(lst)]  # createTransition (:1113-1120)
    hmm = DifferentiableHigherOrderHMM(trainingParameterSet, names, None, None, emissions, ess, *te)
    hmm.setOutputStream(None)
    return hmm


def createSunflowerHMM(pars, con: AlphabetContainer | None, ess: float, expectedSequenceLength: int, startCentral: bool, motifLength,
                       motifProb=None):
    """HMMFactory.createSunflowerHMM (HMMFactory.java:291-400, without phylogenetic emissions): state 0 = background "bg",
    then for every motif m a chain "motif m position p"; bg -> bg | first state of a motif, the last state of a motif -> bg.
    The equivalent sample size is propagated over expectedSequenceLength positions exactly like :346-362."""
    con = con or DNA
    motifLength = [int(x) for x in motifLength]
    M = len(motifLength)
    motifProb = [0.1 / M] * M if motifProb is None else [float(x) for x in motifProb]
    anz = 1
    states = [0] * (1 + M)
    for i in range(M):
        states[1 + i] = anz
        anz += motifLength[i]
    hyperNext = [[0.0] * motifLength[i] for i in range(M)] + [[0.0]]
    self_ = 1.0 - sum(motifProb)  # summed like the reference: self += motifProb[i]; self = 1 - self
    stateESS = [list(r) for r in hyperNext]
    if startCentral:
        hyperNext[M][0] = ess
        hyperParams = [ess]
    else:
        hyperNext[M][0] = ess * self_
        hyperParams = [0.0] * anz
        hyperParams[0] = hyperNext[M][0]
        for i in range(M):
            hyperNext[i] = [motifProb[i] * ess / len(hyperNext[i])] * len(hyperNext[i])
            for k in range(states[1 + i], states[1 + i] + motifLength[i]):
                hyperParams[k] = hyperNext[i][0]
    te = [TransitionElement(None, [0] if startCentral else list(range(anz)), hyperParams)]
    hyperParams = [0.0] * (M + 1)
    for _ in range(int(expectedSequenceLength)):  # propagate ESS (:346-362)
        d = 0.0
        for i in range(M):
            d += hyperNext[i][motifLength[i] - 1]
            for j in range(motifLength[i] - 1, 0, -1):
                stateESS[i][j] += hyperNext[i][j]
                hyperNext[i][j] = hyperNext[i][j - 1]
            stateESS[i][0] += hyperNext[i][0]
            hyperNext[i][0] = motifProb[i] * hyperNext[M][0]
            hyperParams[1 + i] += hyperNext[i][0]
        stateESS[M][0] += hyperNext[M][0]
        hyperParams[0] += self_ * hyperNext[M][0]
        hyperNext[M][0] = d + self_ * hyperNext[M][0]
    em = [DiscreteEmission(con, stateESS[M][0])]
    name = ["bg"]
    te.append(TransitionElement([0], states, hyperParams))  # transition to the petals
    idx = 1
    for m in range(M):
        for p_ in range(motifLength[m]):
            em.append(DiscreteEmission(con, stateESS[m][p_]))
            name.append(f"motif {m} position {p_}")
            te.append(TransitionElement([idx], [0] if p_ == motifLength[m] - 1 else [idx + 1], [ess]))  # "does not matter"
            idx += 1
    hmm = DifferentiableHigherOrderHMM(pars, name, None, None, em, ess, *te)
    hmm.setOutputStream(None)
    return hmm


def _pd(s: str) -> float:
    return math.inf if s == "*" else float(s)


def _log(x: float) -> float:
    return -math.inf if x == 0 else math.log(x)  # Math.log(0) = -Infinity


def parseProfileHMMFromHMMer(lines, consensus: list | None = None):
    """HMMFactory.parseProfileHMMFromHMMer (:434-582) for DNA profile HMMs in HMMER3 text format -> (hmm, background
    frequencies as log values).  `lines`: an iterable of text lines (an open file).  Parameters are set in the reference's
    flat order: emissions in emission-index order, then the transition elements with more than one child in element order."""
    it = iter(lines)
    for s in it:
        if s.startswith("HMM "):
            break
    else:
        raise ValueError("no HMM section")
    alph = s.replace("HMM", "", 1).split()
    if [a.upper() for a in alph] != list("ACGT"):
        raise NotImplementedError("only DNA profile HMMs are mirrored")
    al = 4
    next(it)  # column headers of the transitions
    hom = next(it).split("COMPO")[-1].split()
    emissionPars, transitionPars = [], []

    def add(dst, parts):
        for i in range(al):
            dst.append(-_pd(parts[i]))

    add(emissionPars, next(it).split())  # I0
    t0 = next(it).split()
    BI0 = math.exp(-_pd(t0[1]))
    transitionPars += [_log(1 - BI0), _log(BI0)]
    BM1, BD1 = math.exp(-_pd(t0[0])), math.exp(-_pd(t0[2]))
    transitionPars += [_log(BD1 / (BD1 + BM1)), _log(BM1 / (BD1 + BM1))]
    transitionPars += [-_pd(t0[4]), -_pd(t0[3])]  # I0->I0, I0->M1
    n_nodes = 0
    for s in it:
        if s.strip() == "//":
            break
        me = s.split()[1:]  # drop the node number
        ie = next(it).split()
        add(emissionPars, ie)
        add(emissionPars, me)
        if consensus is not None:
            consensus.append(me[al + 1])
        tr = next(it).split()
        mm, mi, md, im, ii, dm, dd = (-_pd(x) for x in tr[:7])
        transitionPars += [dd, dm, mi, md, mm, ii, im]
        n_nodes += 1
    for _ in range(3):
        add(emissionPars, hom)
    transitionPars += [0.0, 0.0, 0.0, -math.inf, -math.inf, math.log(0.99), math.log(0.01)]
    pars = ViterbiParameterSet(1, IterationCondition(10), 1)  # the reference uses a numerical parameter set here (gradient training: out of scope)
    hmm = createProfileHMM(pars, "PLAN7", 1, n_nodes + 2, DNA, 16.0)
    p = hmm.getCurrentParameterValues()
    p[:len(emissionPars)] = emissionPars
    p[len(emissionPars):len(emissionPars) + len(transitionPars)] = transitionPars
    hmm.setParameters(p, 0)
    hmm.setSkiptInit(True)
    return hmm, np.array([-_pd(x) for x in hom[:al]])
