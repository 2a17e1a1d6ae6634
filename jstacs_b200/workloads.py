"""Synthetic workloads of BASELINE.json / BASELINE.md §3: models with deterministic closed-form initial parameters
and java.util.Random DNA, so a Jstacs (Java) harness and this repo evaluate identical inputs."""
from __future__ import annotations

import math

import numpy as np

from .data import DNA, DataSet, JavaRandom, java_random_dna
from .hmm import (BaumWelchParameterSet, HigherOrderDiscreteEmission, HMMFactory, IterationCondition,
                  TransitionElement, HigherOrderHMM, DifferentiableHigherOrderHMM)


def closed_form_rows(n_rows: int, n_cols: int, salt: float) -> np.ndarray:
    """Normalised rows of 1 + 0.5*sin(i*0.7 + j*1.3 + salt) as LOG probabilities (SURVEY §8d "Initial parameters")."""
    i = np.arange(n_rows, dtype=np.float64)[:, None]
    j = np.arange(n_cols, dtype=np.float64)[None, :]
    w = 1.0 + 0.5 * np.sin(i * 0.7 + j * 1.3 + salt)
    return np.log(w / w.sum(axis=1, keepdims=True))


def set_closed_form_parameters(hmm: HigherOrderHMM):
    """Deterministic, RNG-free initial parameters for every emission and transition element."""
    for s, e in enumerate(hmm.emission):
        if e.order >= 0:
            e.setLogProbabilities(closed_form_rows(e.params.shape[0], e.params.shape[1], float(s)))
    T = hmm.transition.transitions
    for t, te in enumerate(T):
        n = te.getNumberOfChildren()
        if n > 0:
            te.setLogProbabilities(closed_form_rows(1, n, 0.37 * t + 5.0)[0])
    for t, te in enumerate(T):
        if hmm.transition.transIndex[t] != t:
            te.setLogProbabilities(T[hmm.transition.transIndex[t]].parameters)
    hmm.setSkiptInit(True)
    hmm._push_params()
    return hmm


def ergodic_hmm(n_states: int, emission_order: int, n_iter: int = 10, ess: float = 0.0, transition_order: int = 1,
                self_part: float = 0.9, expected_len: float = 1000.0, threads: int = 1) -> DifferentiableHigherOrderHMM:
    """HMMFactory.createErgodicHMM with HigherOrderDiscreteEmission(DNA, ess, order) per state and closed-form parameters."""
    pars = BaumWelchParameterSet(1, IterationCondition(n_iter), threads)
    em = [HigherOrderDiscreteEmission(DNA, ess, emission_order) for _ in range(n_states)]
    hmm = HMMFactory.createErgodicHMM(pars, transition_order, ess, self_part, expected_len, *em)
    hmm.setOutputStream(None)
    return set_closed_form_parameters(hmm)


def banded_hmm(n_states: int = 128, out_degree: int = 4, emission_order: int = 0, ess: float = 0.0, offsets=None) -> HigherOrderHMM:
    """cfg5: each state -> itself + the next out_degree-1 states (mod S), E = S*out_degree edges.
    offsets: the children of state s are (s + d) mod S for d in offsets, in that (child) order, instead of 0..out_degree-1."""
    pars = BaumWelchParameterSet(1, IterationCondition(1), 1)
    em = [HigherOrderDiscreteEmission(DNA, ess, emission_order) for _ in range(n_states)]
    te = [TransitionElement(None, list(range(n_states)), [ess / n_states] * n_states)]
    offsets = list(range(out_degree)) if offsets is None else list(offsets)
    for s in range(n_states):
        te.append(TransitionElement([s], [(s + d) % n_states for d in offsets], [ess] * len(offsets)))
    hmm = DifferentiableHigherOrderHMM(pars, [str(i) for i in range(n_states)], None, None, em, ess, *te)
    hmm.setOutputStream(None)
    return set_closed_form_parameters(hmm)


def ragged_lengths(n_seq: int, seed: int = 42, lo: float = 100.0, decades: float = 3.0) -> np.ndarray:
    """cfg4: lengths floor(100 * 10^(3u)), u ~ U[0,1) from java.util.Random(42) (log-uniform 100 bp - 100 kb)."""
    r = JavaRandom(seed)
    return np.array([int(math.floor(lo * 10.0 ** (decades * r.nextDouble()))) for _ in range(n_seq)], dtype=np.int64)


def planted_dna(n_seq: int, length: int, n_states: int = 2, seed: int = 7) -> DataSet:
    """DNA sampled from a fixed sticky HMM so EM has signal (SURVEY §8d 'planted' variant); numpy RNG (not on the Java stream)."""
    rng = np.random.default_rng(seed)
    emis = rng.dirichlet(np.ones(4) * 0.7, size=n_states)
    stay = 0.98
    states = np.zeros((n_seq, length), dtype=np.int64)
    states[:, 0] = rng.integers(0, n_states, n_seq)
    switch = rng.random((n_seq, length)) > stay
    jump = rng.integers(1, max(n_states, 2), (n_seq, length))
    for p in range(1, length):
        states[:, p] = np.where(switch[:, p], (states[:, p - 1] + jump[:, p]) % n_states, states[:, p - 1])
    u = rng.random((n_seq, length))
    cdf = np.cumsum(emis, axis=1)[states]
    codes = (u[..., None] > cdf).sum(axis=-1).clip(0, 3).astype(np.uint8)
    offsets = np.arange(n_seq + 1, dtype=np.int64) * length
    return DataSet.from_packed(codes.reshape(-1), offsets, DNA, f"planted({n_seq}x{length})")


# name -> (states, emission order, N, L) of BASELINE.json:configs
CONFIGS = {
    "cfg1": dict(n_states=2, emission_order=1, n_seq=10_000, length=1000, n_iter=10),
    "cfg2": dict(n_states=8, emission_order=2, n_seq=1_000_000, length=1000, n_iter=1),
    "cfg3": dict(n_states=32, emission_order=3, n_seq=8_000_000, length=1000, n_iter=20),
    "cfg3_1gpu": dict(n_states=32, emission_order=3, n_seq=1_000_000, length=1000, n_iter=20),
    "cfg4": dict(n_states=16, emission_order=1, n_seq=100_000, length=None, n_iter=1),
}


def make_data(n_seq: int, length, start_index: int = 0) -> DataSet:
    return java_random_dna(n_seq, length, start_index=start_index)
