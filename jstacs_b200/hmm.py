"""Host-side mirror of the reference's HMM operator interface, backed by the native sm_100a library.

Same class / method names, argument meaning and error behaviour as the reference package
  HMM/ = de/jstacs/sequenceScores/statisticalModels/trainable/hmm/     (paths relative to /root/reference)
so the parity tests read like Jstacs user code (cf. supplementary/cookbook/Cookbook.java:410-419).

Everything numeric is done by libjstacs_hmm_b200.so through the C ABI of include/jstacs_hmm.h; this file only
holds the model description (context graph construction = BasicHigherOrderTransition.init) and parameter
bookkeeping.  There is no CPU fallback: without the CUDA library every compute method raises.

WHAT IS A RESTATEMENT HERE (and not from-scratch product code): `BasicHigherOrderTransition._init` /
`_add_and_topsort` follow BasicHigherOrderTransition.java:191-431 statement by statement (same variable names, same
control flow), and `HMMFactory.createErgodicHMM` / `_addTransitions` follow HMMFactory.java:123-181.  They are the Python
stand-in for the Java host side (the image has no JVM): the ORDER of the contexts inside a layer decides the summation
order of the reference's log-sum-exp, so it has to be reproduced exactly for 1e-9 parity, and the hyper-parameters of the
factory have to be the reference's.  In a Jstacs deployment this code does not exist — GpuTransitionExport.java reads
Jstacs' own lookup tables.  tests/test_independent_ir.py checks the result against a first-principles construction that
shares nothing with it.
"""
from __future__ import annotations

import math
import sys
from collections import deque

import numpy as np

from .data import AlphabetContainer, DataSet, Sequence, WrongAlphabetException, WrongLengthException

NEG_INF = float("-inf")


# --------------------------------------------------------------------------------------------------
# numerics used for parameter bookkeeping (never on the data path)
def getLogSum(values) -> float:
    """Normalisation.getLogSum (de/jstacs/utils/Normalisation.java:65-99)."""
    offset = NEG_INF
    for v in values:
        if v > offset:
            offset = v
    if math.isinf(offset):
        return NEG_INF
    s = 0.0
    for v in values:
        s += math.exp(v - offset)
    return offset + math.log(s)


# --------------------------------------------------------------------------------------------------
# termination conditions (de/jstacs/algorithms/optimization/termination/)
class IterationCondition:
    """IterationCondition.java:83-86 — doNextIteration = iteration < maxIter."""

    def __init__(self, maxIter: int):
        if maxIter < 0:
            raise ValueError("maxIter has to be non-negative")
        self.maxIter = int(maxIter)

    def isSimple(self) -> bool:
        return True

    def doNextIteration(self, iteration, f_last, f_current, *unused) -> bool:
        return iteration < self.maxIter


class SmallDifferenceOfFunctionEvaluationsCondition:
    """SmallDifferenceOfFunctionEvaluationsCondition.java:86-89 — eps <= |f_last - f_current|."""

    def __init__(self, eps: float):
        if eps < 0:
            raise ValueError("eps has to be non-negative")
        self.eps = float(eps)

    def isSimple(self) -> bool:
        return True

    def doNextIteration(self, iteration, f_last, f_current, *unused) -> bool:
        return self.eps <= abs(f_last - f_current)


# --------------------------------------------------------------------------------------------------
# training parameter sets (HMM/training/*.java)
class MaxHMMTrainingParameterSet:
    def __init__(self, starts: int, tc, threads: int = 1):
        if starts < 1:
            raise ValueError("The number of starts has to be positive.")
        self.starts, self.tc, self.threads = int(starts), tc, int(threads)

    def getNumberOfStarts(self) -> int:
        return self.starts

    def getTerminationCondition(self):
        return self.tc

    def getNumberOfThreads(self) -> int:
        return self.threads

    def hasDefaultOrIsSet(self) -> bool:  # BaumWelchParameterSet.java:66-72
        return self.tc is not None and self.tc.isSimple()


class BaumWelchParameterSet(MaxHMMTrainingParameterSet):
    """HMM/training/BaumWelchParameterSet.java — (starts, terminationCondition, threads).
    `threads` only sized the Java worker pool (MultiThreadedTrainingParameterSet.java:84-86); the native
    path ignores it."""


class ViterbiParameterSet(MaxHMMTrainingParameterSet):
    """HMM/training/ViterbiParameterSet.java"""


# --------------------------------------------------------------------------------------------------
# emissions (HMM/states/emissions/)
class SilentEmission:
    """HMM/states/emissions/SilentEmission.java — score 0, no statistics."""

    order = -1

    def __init__(self, con: AlphabetContainer | None = None):
        self.con = con

    def clone(self):
        return SilentEmission(self.con)


class AbstractConditionalDiscreteEmission:
    """HMM/states/emissions/discrete/AbstractConditionalDiscreteEmission.java:100-300."""

    def __init__(self, con: AlphabetContainer, hyperParams, initHyperParams=None, norm: bool = True, order: int = 0):
        if not con.isDiscrete() or not con.isSimple():
            raise WrongAlphabetException("The AlphabetContainer has to be discrete and simple.")
        hp = np.array(hyperParams, dtype=np.float64)
        if (hp < 0).any():
            raise ValueError("Please check the hyper-parameters.")
        self.con = con
        self.order = int(order)
        self.hyperParams = hp
        self.initHyperParams = hp if initHyperParams is None else np.array(initHyperParams, dtype=np.float64)
        self.params = np.zeros_like(hp)
        self.logNorm = np.zeros(hp.shape[0])
        self.norm = norm
        self.precompute()

    @staticmethod
    def getHyperParams(ess: float, numConditions: int, numEmissions: int):
        """ACDE:161-198 — ess spread evenly: ess/numConditions/numEmissions per cell."""
        return np.full((numConditions, numEmissions), (ess / float(numConditions)) / float(numEmissions))

    def precompute(self):
        """ACDE:466-480 precompute -> Normalisation.logSumNormalisation per condition."""
        if self.norm:
            for i in range(self.params.shape[0]):
                self.logNorm[i] = getLogSum(self.params[i])
        else:
            self.logNorm[:] = 0

    def setParameter(self, params, offset: int):
        """ACDE:707-714."""
        n = self.params.size
        self.params[...] = np.asarray(params[offset:offset + n], dtype=np.float64).reshape(self.params.shape)
        self.precompute()

    def setLogProbabilities(self, logp):
        self.params[...] = np.asarray(logp, dtype=np.float64).reshape(self.params.shape)
        self.precompute()

    def getNumberOfParameters(self) -> int:
        return int(self.params.size)

    def initializeFunctionRandomly(self, rng: np.random.Generator):
        """ACDE:460-462 drawParameters(initHyperParams): Dirichlet draw per condition (the reference RNG is
        unseeded, de/jstacs/utils/random/DirichletMRG.java:41, so only the distribution is mirrored)."""
        for i in range(self.params.shape[0]):
            alpha = self.initHyperParams[i]
            if alpha.sum() == 0:
                alpha = np.ones_like(alpha)
            self.params[i] = np.log(rng.dirichlet(np.maximum(alpha, 1e-300)))
        self.precompute()

    def clone(self):
        c = object.__new__(type(self))
        c.__dict__.update(self.__dict__)
        for k in ("hyperParams", "initHyperParams", "params", "logNorm"):
            setattr(c, k, getattr(self, k).copy())
        return c


class DiscreteEmission(AbstractConditionalDiscreteEmission):
    """HMM/states/emissions/discrete/DiscreteEmission.java:42-70 — order 0 (condition index 0)."""

    def __init__(self, con: AlphabetContainer, ess_or_hyper, initHyperParams=None):
        a = con.getAlphabetLengthAt(0)
        if np.isscalar(ess_or_hyper):
            hyper = self.getHyperParams(float(ess_or_hyper), 1, a)
        else:
            hyper = np.asarray(ess_or_hyper, dtype=np.float64).reshape(1, a)
        super().__init__(con, hyper, None if initHyperParams is None else np.asarray(initHyperParams).reshape(1, a), True, 0)


class HigherOrderDiscreteEmission(AbstractConditionalDiscreteEmission):
    """HMM/states/emissions/discrete/HigherOrderDiscreteEmission.java:84-91 — a^order conditions,
    cond = sum_i a^i x[pos-1-i] (:138-146)."""

    def __init__(self, con: AlphabetContainer, ess: float, order: int, initESS: float | None = None, norm: bool = True):
        if order < 0:
            raise ValueError("The order must be non-negative")
        a = con.getAlphabetLengthAt(0)
        n_cond = int(a ** order)
        hyper = self.getHyperParams(float(ess), n_cond, a)
        init = None if initESS is None else self.getHyperParams(float(initESS), n_cond, a)
        super().__init__(con, hyper, init, norm, order)


# --------------------------------------------------------------------------------------------------
# transition elements (HMM/transitions/BasicHigherOrderTransition.java:695-1411, elements/TransitionElement.java)
class BasicTransitionElement:
    """AbstractTransitionElement(context, states, hyperParameters[, weight, norm]) (BHOT:757-813).
    Child order is the user's order; it is NOT sorted (BHOT:794-803 sorts a copy only to detect duplicates)."""

    def __init__(self, context, states, hyperParameters=None, weight=None, norm: bool = True):
        self.context = [] if context is None else [int(c) for c in context]
        self.states = [] if states is None else [int(s) for s in states]
        s = len(self.states)
        h = s if hyperParameters is None else len(hyperParameters)
        if s != h:
            raise ValueError("You have to provide the same number of states and hyperparameters.")
        if len(set(self.states)) != s:
            raise ValueError("It is not allowed to have several edges to the same child.")
        self.hyperParameters = np.zeros(s) if hyperParameters is None else np.array(hyperParameters, dtype=np.float64)
        if (self.hyperParameters < 0).any():
            raise ValueError("Please check the hyper-parameters.")
        self.parameters = np.zeros(s)
        self.norm = norm
        self.descendants = [-1] * s
        self.logNorm = 0.0
        self.precompute()

    def precompute(self):
        """BHOT:935-941 / TransitionElement.java:119-126."""
        self.logNorm = getLogSum(self.parameters) if self.norm else 0.0

    def getNumberOfChildren(self) -> int:
        return len(self.states)

    def getLastContextState(self) -> int:
        return self.context[-1]

    def getNextContext(self, index: int, maximalMarkovOrder: int):
        """BHOT:1097-1105."""
        n = len(self.context)
        my_ord = n + 1 if n < maximalMarkovOrder else maximalMarkovOrder
        if my_ord == 0:
            return []
        src = self.context[0 if n < maximalMarkovOrder else 1:]
        return list(src[:my_ord - 1]) + [self.states[index]]

    def setLogProbabilities(self, logp):
        self.parameters[...] = np.asarray(logp, dtype=np.float64)
        self.precompute()

    def initializeRandomly(self, rng: np.random.Generator):
        """BHOT:1205-1224, 1281-1285 drawParametersFromStatistic with empty statistic (prior draw)."""
        if len(self.states) > 1:
            alpha = self.hyperParameters.copy()
            if alpha.sum() == 0:
                alpha = np.ones_like(alpha)
            self.parameters[...] = np.log(rng.dirichlet(np.maximum(alpha, 1e-300)))
        else:
            self.parameters[...] = 0
        self.precompute()

    def clone(self):
        c = object.__new__(type(self))
        c.__dict__.update(self.__dict__)
        c.context = list(self.context)
        c.states = list(self.states)
        c.descendants = list(self.descendants)
        c.hyperParameters = self.hyperParameters.copy()
        c.parameters = self.parameters.copy()
        return c


class TransitionElement(BasicTransitionElement):
    """HMM/transitions/elements/TransitionElement.java — the differentiable element HMMFactory uses."""


# --------------------------------------------------------------------------------------------------
class BasicHigherOrderTransition:
    """HMM/transitions/BasicHigherOrderTransition.java — context graph over transition elements."""

    def __init__(self, isSilent, transIndex, *transitions):
        self.isSilent = [bool(b) for b in isSilent]
        self.transitions = [t.clone() for t in transitions]
        n = len(self.transitions)
        if transIndex is None:
            self.transIndex = list(range(n))
        else:
            if len(transIndex) != n:
                raise ValueError("The transIndex has to have the same length as the TransitionElements.")
            for i, ti in enumerate(transIndex):
                if ti < 0 or ti > i:
                    raise ValueError(f"The constraint of the transIndex is not met: 0<={ti}<={i}")
                if ti != i and self.transitions[i].getNumberOfChildren() != self.transitions[ti].getNumberOfChildren():
                    raise ValueError(f"Error in transIndex: TransitionElements {i} and {ti} don't have the same number of children.")
            self.transIndex = [int(t) for t in transIndex]
        self._init()

    # BHOT:191-257
    def _add_and_topsort(self, currentLayer, nextLayer, lookup, canBeUsed, used, inDeg, nextUsed):
        T = self.transitions
        for k in range(len(T)):
            canBeUsed[k] = False
            used[k] = False
            inDeg[k] = 0
        roots = deque()
        thresh = len(currentLayer)
        i = 0
        while i < len(currentLayer):
            idx = currentLayer[i]
            for s in range(T[idx].getNumberOfChildren()):
                d = T[idx].descendants[s]
                if self.isSilent[T[idx].states[s]]:
                    if not canBeUsed[d]:
                        currentLayer.append(d)
                        canBeUsed[d] = True
                    if i >= thresh:
                        inDeg[d] += 1
                else:
                    if not nextUsed[d]:
                        nextLayer.append(d)
            i += 1
        for t in range(len(T)):
            if canBeUsed[t]:
                if inDeg[t] == 0:
                    roots.append(t)
            else:
                used[t] = True
        while roots:
            idx = roots.popleft()
            used[idx] = True
            lookup.append(idx)
            for j in range(T[idx].getNumberOfChildren()):
                child = T[idx].descendants[j]
                if canBeUsed[child]:
                    inDeg[child] -= 1
                    if inDeg[child] == 0:
                        roots.append(child)
        for t in range(len(T)):
            if canBeUsed[t] and not used[t]:
                raise ValueError(f"Check transition element {t}: {T[t].context} -> {T[t].states}")

    # BHOT:268-431
    def _init(self):
        self.maximalMarkovOrder = max((len(t.context) for t in self.transitions), default=0)
        m = self.maximalMarkovOrder
        elements = list(self.transitions)
        by_ctx = {}
        for s, te2 in enumerate(elements):
            by_ctx[tuple(te2.context)] = s  # the last element with a context wins, duplicates are rejected below
        t = 0
        while t < len(elements):
            te = elements[t]
            for child in range(te.getNumberOfChildren()):
                nxt = tuple(te.getNextContext(child, m))
                if nxt in by_ctx:
                    te.descendants[child] = by_ctx[nxt]
                if te.descendants[child] == -1:
                    if m == 0:
                        te.descendants[child] = 0
                    else:
                        te.descendants[child] = len(elements)
                        by_ctx[nxt] = len(elements)
                        elements.append(TransitionElement(list(nxt), None, None))
            t += 1
        if len(elements) > len(self.transitions):
            self.transIndex = self.transIndex + list(range(len(self.transIndex), len(elements)))
            self.transitions = elements
        T = self.transitions
        n = len(T)
        inDeg = [0] * n
        seen = {}
        for t in range(n):
            for d in T[t].descendants:
                inDeg[d] += 1
            key = tuple(T[t].context)
            if key in seen:
                raise ValueError(f"The context {T[t].context} is used by more than one TransitionElements")
            seen[key] = t
        self.maxInDegree = max(inDeg) if n else 0
        if n > 1 and min(inDeg[1:]) <= 0:
            raise ValueError("The in-degree of some TransitionElements is zero.")

        self.lookup0 = [None] * (m + 1)
        self.lookup1 = [None] * (m + 1)

        def createLookup(layer, current):
            self.lookup0[layer] = list(current)
            l1 = [-1] * n
            for i, e in enumerate(current):
                l1[e] = i
            self.lookup1[layer] = l1

        canBeUsed, used, nextUsed = [False] * n, [False] * n, [False] * n
        currentLayer, nextLayer, current = [], [], []
        if m == 0:
            if any(self.isSilent):
                raise ValueError("A hidden Markov model of order zero is not alllowed to have any silent states.")
            createLookup(0, [0])
        else:
            for t in range(n):
                if len(T[t].context) == 0:
                    currentLayer.append(t)
                    current.append(t)
            self._add_and_topsort(currentLayer, nextLayer, current, canBeUsed, used, inDeg, nextUsed)
            createLookup(0, current)
            currentLayer, nextLayer = nextLayer, currentLayer
            for p in range(1, m):
                current.clear()
                nextLayer.clear()
                for k in range(n):
                    nextUsed[k] = False
                for i in range(len(currentLayer)):
                    idx = currentLayer[i]
                    current.append(idx)
                    for s in range(T[idx].getNumberOfChildren()):
                        d = T[idx].descendants[s]
                        if not self.isSilent[T[idx].states[s]] and not nextUsed[d]:
                            nextLayer.append(d)
                            nextUsed[d] = True
                self._add_and_topsort(currentLayer, nextLayer, current, canBeUsed, used, inDeg, nextUsed)
                createLookup(p, current)
                currentLayer, nextLayer = nextLayer, currentLayer
            currentLayer.clear()
            nextLayer.clear()
            current.clear()
            for t in range(n):
                if len(T[t].context) == m and not self.isSilent[T[t].getLastContextState()]:
                    currentLayer.append(t)
                    current.append(t)
            self._add_and_topsort(currentLayer, nextLayer, current, canBeUsed, used, inDeg, nextUsed)
            createLookup(m, current)

    # --- public Transition interface (HMM/transitions/Transition.java:63-199) ---
    def getMaximalMarkovOrder(self) -> int:
        return self.maximalMarkovOrder

    def _lc(self, layer: int) -> int:
        return layer if layer < self.maximalMarkovOrder else self.maximalMarkovOrder

    def getNumberOfIndexes(self, layer: int) -> int:
        return len(self.lookup0[self._lc(layer)])

    def getTransitionElementIndex(self, layer: int, index: int) -> int:
        return self.lookup0[self._lc(layer)][index]

    def getNumberOfChildren(self, layer: int, index: int) -> int:
        return self.transitions[self.getTransitionElementIndex(layer, index)].getNumberOfChildren()

    def fillTransitionInformation(self, layer: int, index: int, childIdx: int):
        te = self.transitions[self.getTransitionElementIndex(layer, index)]
        state = te.states[childIdx]
        adv = 0 if self.isSilent[state] else 1
        return state, self.lookup1[self._lc(layer + adv)][te.descendants[childIdx]], adv

    def getLogScoreFor(self, layer: int, index: int, childIdx: int, seq=None, pos: int = 0) -> float:
        te = self.transitions[self.getTransitionElementIndex(layer, index)]
        return float(te.parameters[childIdx] - te.logNorm)

    def getLastContextState(self, layer: int, index: int) -> int:
        te = self.transitions[self.getTransitionElementIndex(layer, index)]
        return -1 if len(te.context) == 0 else te.getLastContextState()

    def isAbsorbing(self):
        """BHOT:667-676."""
        absorbing = [True] * len(self.isSilent)
        for te in self.transitions:
            if len(te.context) > 0 and len(te.states) > 0:
                absorbing[te.getLastContextState()] = False
        return absorbing

    def getMaximalInDegree(self) -> int:
        return self.maxInDegree


class HigherOrderTransition(BasicHigherOrderTransition):
    """HMM/transitions/HigherOrderTransition.java — adds flat parameter get/set (:219-286):
    elements with transIndex[t]==t and more than one child, in element order."""

    def getNumberOfParameters(self) -> int:
        return sum(t.getNumberOfChildren() for i, t in enumerate(self.transitions)
                   if self.transIndex[i] == i and t.getNumberOfChildren() > 1)

    def fillParameters(self, params, offset: int) -> int:
        for i, t in enumerate(self.transitions):
            if self.transIndex[i] == i and t.getNumberOfChildren() > 1:
                n = t.getNumberOfChildren()
                params[offset:offset + n] = t.parameters
                offset += n
        return offset

    def setParameters(self, params, start: int) -> int:
        for i, t in enumerate(self.transitions):
            if self.transIndex[i] == i:
                n = t.getNumberOfChildren()
                if n > 1:
                    t.parameters[...] = params[start:start + n]
                    t.precompute()
                    start += n
            else:
                t.parameters[...] = self.transitions[self.transIndex[i]].parameters
                t.precompute()
        return start


# --------------------------------------------------------------------------------------------------
class ModelIR:
    """Flattened model in the layout of jh_model_desc (include/jstacs_hmm.h)."""

    FIELDS_I32 = ("lc_ctx_ptr", "ctx_elem", "ctx_last_state", "elem_child_ptr", "child_state", "child_desc",
                  "elem_ctx", "trans_index", "state_emis", "emis_order", "emis_cond_ptr")
    FIELDS_F64 = ("trans_param", "trans_lognorm", "trans_hyper", "emis_param", "emis_lognorm", "emis_hyper")
    FIELDS_U8 = ("state_silent", "state_final")

    def __init__(self):
        self.n_states = self.alphabet = self.max_order = self.n_elem = self.n_emis = 0

    @property
    def n_child_total(self) -> int:
        return int(self.elem_child_ptr[-1])

    @property
    def n_cond_total(self) -> int:
        return int(self.emis_cond_ptr[-1])


def flatten(transition: BasicHigherOrderTransition, emission, emissionIdx, finalState, alphabet: int) -> ModelIR:
    """What the Java shim does through the public Transition interface (SURVEY §8b "Model export")."""
    ir = ModelIR()
    T = transition.transitions
    m = transition.getMaximalMarkovOrder()
    ir.n_states, ir.alphabet, ir.max_order, ir.n_elem, ir.n_emis = len(emissionIdx), int(alphabet), m, len(T), len(emission)
    lc_ptr, ctx_elem, ctx_last = [0], [], []
    for lc in range(m + 1):
        for c in range(transition.getNumberOfIndexes(lc)):
            ctx_elem.append(transition.getTransitionElementIndex(lc, c))
            ctx_last.append(transition.getLastContextState(lc, c))
        lc_ptr.append(len(ctx_elem))
    ir.lc_ctx_ptr = np.array(lc_ptr, np.int32)
    ir.ctx_elem = np.array(ctx_elem, np.int32)
    ir.ctx_last_state = np.array(ctx_last, np.int32)
    ptr = [0]
    for t in T:
        ptr.append(ptr[-1] + t.getNumberOfChildren())
    ir.elem_child_ptr = np.array(ptr, np.int32)
    cat_i = lambda xs: np.array([v for x in xs for v in x], np.int32)
    cat_f = lambda xs: np.array([v for x in xs for v in x], np.float64)
    ir.child_state = cat_i(t.states for t in T)
    ir.child_desc = cat_i(t.descendants for t in T)
    ir.elem_ctx = np.array([transition.lookup1[lc][e] for lc in range(m + 1) for e in range(len(T))], np.int32)
    ir.trans_index = np.array(transition.transIndex, np.int32)
    ir.trans_param = cat_f(t.parameters for t in T)
    ir.trans_lognorm = np.array([t.logNorm for t in T], np.float64)
    ir.trans_hyper = cat_f(t.hyperParameters for t in T)
    ir.state_emis = np.array(emissionIdx, np.int32)
    ir.emis_order = np.array([e.order for e in emission], np.int32)
    cptr = [0]
    for e in emission:
        cptr.append(cptr[-1] + (0 if e.order < 0 else e.params.shape[0]))
    ir.emis_cond_ptr = np.array(cptr, np.int32)
    real = [e for e in emission if e.order >= 0]
    for e in real:
        if e.params.shape[1] != alphabet:
            raise WrongAlphabetException("All emissions have to use the same alphabet.")
    ir.emis_param = np.concatenate([e.params.reshape(-1) for e in real]) if real else np.zeros(0)
    ir.emis_lognorm = np.concatenate([e.logNorm for e in real]) if real else np.zeros(0)
    ir.emis_hyper = np.concatenate([e.hyperParams.reshape(-1) for e in real]) if real else np.zeros(0)
    ir.state_silent = np.array([1 if emission[i].order < 0 else 0 for i in emissionIdx], np.uint8)
    ir.state_final = np.array([1 if f else 0 for f in finalState], np.uint8)
    return ir


def _native_error():
    from . import _native

    return _native.NativeError


# --------------------------------------------------------------------------------------------------
class _Stream:
    """SafeOutputStream (de/jstacs/utils/SafeOutputStream.java:41-61); None silences."""

    def __init__(self, o):
        self.o = o

    def writeln(self, s):
        if self.o is not None:
            self.o.write(str(s) + "\n")


class HigherOrderHMM:
    """HMM/models/HigherOrderHMM.java + HMM/AbstractHMM.java public surface, native-backed.

    HigherOrderHMM(trainingParameterSet, name, emission, *te, emissionIdx=None, forward=None, transIndex=None)
    mirrors HigherOrderHMM.java:156-187.
    """

    def __init__(self, trainingParameterSet, name, emission, *te, emissionIdx=None, forward=None, transIndex=None,
                 filter=None, statesGroups=None, sequenceAnnotationType=None):
        if not trainingParameterSet.hasDefaultOrIsSet():
            raise ValueError("Please check the training parameters.")  # AbstractHMM.java:247-249
        self.trainingParameter = trainingParameterSet
        n = len(name)
        if len(set(name)) != n:
            raise ValueError("The state names should be unique.")  # AbstractHMM.java:264-270
        self.name = list(name)
        self.emissionIdx = list(range(n)) if emissionIdx is None else [int(e) for e in emissionIdx]
        if len(self.emissionIdx) != n:
            raise ValueError("emissionIdx must have one entry per state")
        if forward is not None and not all(forward):
            raise NotImplementedError("reverse-strand states are outside the native path (JH_ERR_UNSUPPORTED)")
        if filter is not None and any(f is not None for f in filter):
            raise NotImplementedError("state filters are outside the native path (JH_ERR_UNSUPPORTED)")
        if len(emission) > n:
            raise ValueError("more emissions than states")
        self.emission = [e.clone() for e in emission]  # ArrayHandler.clone, AbstractHMM.java:301
        cons = [e.con for e in self.emission if e.order >= 0]
        if not cons:
            raise ValueError("at least one emitting state is needed")
        self.alphabets = cons[0]
        for c in cons[1:]:
            if not c.checkConsistency(self.alphabets):
                raise WrongAlphabetException("All emissions have to use the same AlphabetContainer.")
        isSilent = [self.emission[i].order < 0 for i in self.emissionIdx]
        # AbstractHMM.initTransition (:359-382)
        cls = HigherOrderTransition if all(isinstance(t, TransitionElement) for t in te) else BasicHigherOrderTransition
        self.transition = cls(isSilent, transIndex, *te)
        for t in self.transition.transitions:
            for s in t.states:
                if not 0 <= s < n:
                    raise ValueError("transition to an unknown state")
        # determineFinalStates (AbstractHMM.java:1101-1113)
        self.finalState = self.transition.isAbsorbing()
        if not any(self.finalState):
            self.finalState = [not s for s in isSilent]
        # allowed states groups (HigherOrderHMM.java:183-187, AbstractHMM.java:173-228): groups of states that per-position
        # annotations of the training data (DataSet.allowedStatesGroups, the reference's AllowedStatesGroups sequence
        # annotation of type `sequenceAnnotationType`) can restrict the paths to
        self.statesGroups = [] if statesGroups is None else [[int(s) for s in g] for g in statesGroups]
        for g in self.statesGroups:
            if any(not 0 <= s < n for s in g):
                raise ValueError("states group names an unknown state")
        self.type = sequenceAnnotationType
        self.sostream = _Stream(sys.stdout)
        self.skipInit = False
        self.threads = trainingParameterSet.getNumberOfThreads()
        self._rng = np.random.default_rng()
        self._native = None

    # ---- plumbing ---------------------------------------------------------------------------
    def setOutputStream(self, o):
        """AbstractHMM.setOutputStream (:1079)."""
        self.sostream = _Stream(o)

    def setSkiptInit(self, skip: bool):
        """HigherOrderHMM.setSkiptInit (:875) [sic]."""
        self.skipInit = bool(skip)

    def getNumberOfStates(self) -> int:
        return len(self.name)

    def toXML(self) -> str:
        """AbstractHMM.toXML (:394-415) in the reference's XML dialect (jstacs_b200/xmlio.py)."""
        from .xmlio import hmm_to_xml

        return hmm_to_xml(self)

    @staticmethod
    def fromXML(xml: str) -> "HigherOrderHMM":
        """AbstractHMM(StringBuffer xml) (:347-357, fromXML :424-465)."""
        from .xmlio import hmm_from_xml

        return hmm_from_xml(xml)

    def getAlphabetContainer(self):
        return self.alphabets

    def getLength(self) -> int:
        return 0

    def ir(self) -> ModelIR:
        return flatten(self.transition, self.emission, self.emissionIdx, self.finalState,
                       self.alphabets.getAlphabetLengthAt(0))

    def _engine(self):
        from . import _native

        if self._native is None:
            self._native = _native.NativeModel(self.ir())
            if self.statesGroups:
                self._native.set_states_groups(self.statesGroups)
        return self._native

    def _push_params(self):
        if self._native is not None:
            ir = self.ir()
            self._native.set_params(ir.trans_param, ir.trans_lognorm, ir.emis_param, ir.emis_lognorm)

    def _pull_params(self):
        tp, tl, ep, el = self._native.get_params()
        ptr = 0
        for t in self.transition.transitions:
            n = t.getNumberOfChildren()
            t.parameters[...] = tp[ptr:ptr + n]
            ptr += n
        for i, t in enumerate(self.transition.transitions):
            t.logNorm = float(tl[i])
        row = 0
        a = self.alphabets.getAlphabetLengthAt(0)
        for e in self.emission:
            if e.order < 0:
                continue
            nc = e.params.shape[0]
            e.params[...] = ep[row * a:(row + nc) * a].reshape(nc, a)
            e.logNorm[...] = el[row:row + nc]
            row += nc

    def _data(self, data: DataSet, groups=None):
        from . import _native

        if not data.getAlphabetContainer().checkConsistency(self.alphabets):
            raise WrongAlphabetException("The AlphabetContainer of the data set and the model do not match.")
        if data._native is None:
            if hasattr(data, "_make_native"):  # shard files: streamed to the device, never materialised on the host
                data._native = data._make_native(self.alphabets.getAlphabetLengthAt(0))
            else:
                data._native = _native.NativeData(data.codes, data.offsets, None, self.alphabets.getAlphabetLengthAt(0))
        data._native.set_groups(groups)  # None: unconstrained (what every reference method without a groups argument computes)
        return data._native

    # ---- initialisation ------------------------------------------------------------------------
    def initializeRandomly(self):
        """HigherOrderHMM.initializeRandomly (:880-885)."""
        for t in self.transition.transitions:
            t.initializeRandomly(self._rng)
        for i, t in enumerate(self.transition.transitions):
            if self.transition.transIndex[i] != i:
                t.setLogProbabilities(self.transition.transitions[self.transition.transIndex[i]].parameters)
        for e in self.emission:
            if e.order >= 0:
                e.initializeFunctionRandomly(self._rng)
        self._push_params()

    def getLogPriorTerm(self) -> float:
        """HigherOrderHMM.getLogPriorTerm (:253-259), evaluated by the library."""
        self._engine()
        self._push_params()
        return self._native.log_prior()

    # ---- training ------------------------------------------------------------------------------
    def train(self, data: DataSet, weights=None):
        """HigherOrderHMM.train (:733-789): numberOfStarts x { init; do E-step, (M-step) while tc }."""
        from . import _native

        tp = self.trainingParameter
        if not isinstance(tp, MaxHMMTrainingParameterSet):
            raise ValueError("This kind of training is currently not supported.")
        mode = _native.MODE_VITERBI if isinstance(tp, ViterbiParameterSet) else _native.MODE_BAUM_WELCH
        tc = tp.getTerminationCondition()
        eng = self._engine()
        # semi-supervised training (doOneStep, HigherOrderHMM.java:812-820): sequences annotated with allowed states groups
        groups = getattr(data, "allowedStatesGroups", None) if (self.type is not None and self.statesGroups) else None
        nd = self._data(data, groups)
        nd.set_weights(weights)
        best, best_params = NEG_INF, None
        self.last_objectives = []
        for start in range(tp.getNumberOfStarts()):
            self.sostream.writeln(f"start {start} ============================")
            if not self.skipInit:
                self.initializeRandomly()
            self._push_params()
            if isinstance(tc, IterationCondition):
                obj = eng.baum_welch(nd, mode, _native.TERM_ITERATIONS, tc.maxIter, 0.0)
            elif isinstance(tc, SmallDifferenceOfFunctionEvaluationsCondition):
                obj = eng.baum_welch(nd, mode, _native.TERM_SMALL_DIFFERENCE, 0, tc.eps)
            else:
                raise ValueError("only simple termination conditions are supported (BaumWelchParameterSet.java:66-72)")
            old = NEG_INF
            for it, v in enumerate(obj):  # same progress line as HigherOrderHMM.java:764
                self.sostream.writeln(f"{it}\t-\t{v}\t{v - old}")
                old = v
            self.last_objectives.append(list(obj))
            new_value = eng.last_objective()  # the history buffer may be shorter than the run
            if new_value > best:
                best = new_value
                if tp.getNumberOfStarts() > 1:
                    best_params = eng.get_params()
        self.sostream.writeln(f"best result: {best}")
        if best_params is not None:
            eng.set_params(*best_params)
        self._pull_params()
        return best

    # ---- scoring -------------------------------------------------------------------------------
    def getLogScoreFor(self, data, weights=None):
        """HigherOrderHMM.getLogScoreFor(DataSet) (:921-941) -> double[N]; also accepts one Sequence."""
        if isinstance(data, Sequence):
            return self.getLogProbFor(data)
        eng = self._engine()
        self._push_params()
        return eng.loglik(self._data(data))

    def getLogProbFor(self, sequence, startpos: int = 0, endpos: int | None = None, statesGroups=None):
        """AbstractHMM.getLogProbFor (:985-998; with statesGroups :989-998). A DataSet argument scores every element."""
        if isinstance(sequence, DataSet):
            return self.getLogScoreFor(sequence)
        if statesGroups is not None:
            if startpos != 0 or (endpos is not None and endpos != sequence.getLength() - 1):
                raise NotImplementedError("allowed states groups on sub-sequences are outside the native path")
            ds = DataSet("", sequence, con=self.alphabets)
            eng = self._engine()
            self._push_params()
            return float(eng.loglik(self._data(ds, np.asarray(statesGroups, dtype=np.int32)))[0])
        if not sequence.getAlphabetContainer().checkConsistency(self.alphabets):
            raise WrongAlphabetException("The AlphabetContainer of the sequence and the model do not match.")
        if endpos is None:
            endpos = sequence.getLength() - 1
        if startpos < 0 or endpos < startpos or endpos >= sequence.getLength():
            raise WrongLengthException("invalid start/end position")
        ds = DataSet("", sequence, con=self.alphabets)
        return float(self.getLogProbForWindows(ds, [0], [startpos], [endpos])[0])

    def getLogProbForWindows(self, data: DataSet, seq_index, start, end):
        """getLogProbFor(data.getElementAt(seq_index[i]), start[i], end[i]) for many windows in one call (the sliding
        window scan of projects/xanthogenomes/NHMMer.java:225-244); higher-order emissions see the residues before the
        window start, like HigherOrderDiscreteEmission.getConditionIndex (:138-146)."""
        eng = self._engine()
        self._push_params()
        return eng.loglik_windows(self._data(data), seq_index, start, end)

    # ---- decoding ------------------------------------------------------------------------------
    def getViterbiPathsFor(self, data: DataSet, narrow: bool = False):
        """AbstractHMM.getViterbiPathsFor (:894-900) -> list of (IntList path, Double score).
        narrow: uint8 paths (255 = no state) through jh_viterbi_u8: a quarter of the device->host traffic."""
        eng = self._engine()
        self._push_params()
        n_silent = sum(1 for i in self.emissionIdx if self.emission[i].order < 0)
        if n_silent:  # paths contain the silent states: variable length (HigherOrderHMM.java:617-708)
            lens = np.diff(data.offsets)
            paths, scores, plen = eng.viterbi_paths(self._data(data), lens * (1 + n_silent) + n_silent)
            if (plen < 0).any():
                raise ValueError("Impossible path")
            return [(p, float(sc)) for p, sc in zip(paths, scores)]
        try:
            paths, scores = eng.viterbi(self._data(data), narrow=narrow and self.getNumberOfStates() <= 255)
        except _native_error() as e:
            if e.status == -6:  # JH_ERR_IMPOSSIBLE
                raise ValueError("Impossible path") from e  # IllegalArgumentException, HigherOrderHMM.java:659
            raise
        off = data.offsets
        return [(paths[off[i]:off[i + 1]], float(scores[i])) for i in range(data.getNumberOfElements())]

    def getViterbiPathFor(self, seq: Sequence, startPos: int = 0, endPos: int | None = None, allowedStatesGroup=None):
        """AbstractHMM.getViterbiPathFor (:879-881) / HigherOrderHMM.getViterbiPathFor (:590-598; with allowedStatesGroup :594-598)."""
        if endPos is None:
            endPos = seq.getLength() - 1
        if allowedStatesGroup is not None:
            if startPos != 0 or endPos != seq.getLength() - 1:
                raise NotImplementedError("allowed states groups on sub-sequences are outside the native path")
            ds = DataSet("", seq, con=self.alphabets)
            eng = self._engine()
            self._push_params()
            try:
                paths, scores = eng.viterbi(self._data(ds, np.asarray(allowedStatesGroup, dtype=np.int32)))
            except _native_error() as e:
                if e.status == -6:
                    raise ValueError("Impossible path") from e
                raise
            return paths, float(scores[0])
        if startPos == 0 and endPos == seq.getLength() - 1:
            return self.getViterbiPathsFor(DataSet("", seq, con=self.alphabets))[0]
        # a window: the emissions see the residues before startPos (absolute positions, HigherOrderDiscreteEmission.java:138-146)
        eng = self._engine()
        self._push_params()
        ds = DataSet("", seq, con=self.alphabets)
        n_silent = sum(1 for i in self.emissionIdx if self.emission[i].order < 0)
        paths, scores, plen = eng.viterbi_windows(self._data(ds), [0], [startPos], [endPos], n_silent)
        if plen[0] < 0:
            raise ValueError("Impossible path")
        return paths[0], float(scores[0])

    def getLogStatePosteriorMatrixFor(self, seq_or_data, startPos: int = 0, endPos: int | None = None):
        """AbstractHMM.getLogStatePosteriorMatrixFor (:776-838) -> double[S][L] (list of them for a DataSet)."""
        eng = self._engine()
        self._push_params()
        if isinstance(seq_or_data, DataSet):
            nd = self._data(seq_or_data)
            lens = np.diff(seq_or_data.offsets)
            return [eng.posterior(nd, i, self.getNumberOfStates(), int(lens[i])) for i in range(seq_or_data.getNumberOfElements())]
        seq = seq_or_data
        if endPos is None:
            endPos = seq.getLength() - 1
        ds = DataSet("", seq, con=self.alphabets)
        if startPos == 0 and endPos == seq.getLength() - 1:
            return eng.posterior(self._data(ds), 0, self.getNumberOfStates(), ds.getNumberOfResidues())
        return eng.posterior_window(self._data(ds), 0, startPos, endPos, self.getNumberOfStates())

    def getStatePosteriorMatrixFor(self, seq_or_data):
        """AbstractHMM.getStatePosteriorMatrixFor (:811-820)."""
        r = self.getLogStatePosteriorMatrixFor(seq_or_data)
        return [np.exp(x) for x in r] if isinstance(r, list) else np.exp(r)

    @staticmethod
    def decodeStatePosterior(*statePosterior):
        """AbstractHMM.decodeStatePosterior (:1125-1139): np.argmax returns the first maximum == strict '<'."""
        return [np.argmax(p, axis=0).astype(np.int32) for p in statePosterior]

    def decodePosteriorPathsFor(self, data: DataSet, narrow: bool = False):
        """Fused getLogStatePosteriorMatrixFor + decodeStatePosterior over a DataSet (no S x L matrix leaves the GPU)."""
        eng = self._engine()
        self._push_params()
        paths, ll = eng.posterior_decode(self._data(data), narrow=narrow and self.getNumberOfStates() <= 255)
        off = data.offsets
        return [(paths[off[i]:off[i + 1]], float(ll[i])) for i in range(data.getNumberOfElements())]


class DifferentiableHigherOrderHMM(HigherOrderHMM):
    """HMM/models/DifferentiableHigherOrderHMM.java — what HMMFactory returns for discrete emissions
    (HMMFactory.java:111-119). Only the Baum-Welch/Viterbi entry point and the flat parameter accessors are
    mirrored (:316-344, :495-518); the gradient path is out of scope."""

    def __init__(self, trainingParameterSet, name, emissionIdx, forward, emission, ess, *te, transIndex=None):
        super().__init__(trainingParameterSet, name, emission, *te, emissionIdx=emissionIdx, forward=forward,
                         transIndex=transIndex)
        if ess < 0:
            raise ValueError("ess must be non-negative")
        self.ess = ess

    def getNumberOfParameters(self) -> int:
        n = sum(e.getNumberOfParameters() for e in self.emission if e.order >= 0)
        return n + self.transition.getNumberOfParameters()

    def getCurrentParameterValues(self):
        p = np.zeros(self.getNumberOfParameters())
        o = 0
        for e in self.emission:
            if e.order >= 0:
                p[o:o + e.params.size] = e.params.reshape(-1)
                o += e.params.size
        self.transition.fillParameters(p, o)
        return p

    def setParameters(self, params, start: int = 0):
        o = start
        for e in self.emission:
            if e.order >= 0:
                e.setParameter(params, o)
                o += e.params.size
        self.transition.setParameters(np.asarray(params, dtype=np.float64), o)
        self._push_params()


class HMMFactory:
    """HMM/HMMFactory.java — ergodic constructor (:123-181)."""

    @staticmethod
    def _addTransitions(states, ess, selfTransitionPart, o, out):
        n = len(states)
        e = ess * (1 - selfTransitionPart) / (n - 1) if n > 1 else 0.0
        ctx = [0] * o
        while True:  # SequenceIterator(o) with bounds n: position 0 runs fastest
            hyper = [e] * n
            hyper[ctx[o - 1]] = ess * selfTransitionPart
            out.append(TransitionElement(list(ctx), states, hyper))
            i = 0
            while i < o:
                ctx[i] += 1
                if ctx[i] < n:
                    break
                ctx[i] = 0
                i += 1
            if i == o:
                break

    @staticmethod
    def createErgodicHMM(pars, order: int, ess: float, selfTranistionPart: float, expectedSequenceLength: float, *emission):
        n = len(emission)
        for em in emission:
            if isinstance(em, SilentEmission):
                raise ValueError("An ergodic HMM can not contain silent states.")
        name = [str(i) for i in range(n)]
        states = list(range(n))
        te = [TransitionElement(None, states, [ess / n] * n)]
        e = ess / n
        for o in range(1, order):
            HMMFactory._addTransitions(states, e, selfTranistionPart, o, te)
            e /= n
        HMMFactory._addTransitions(states, ess * (expectedSequenceLength - order), selfTranistionPart, order, te)
        return DifferentiableHigherOrderHMM(pars, name, None, None, list(emission), ess, *te)
