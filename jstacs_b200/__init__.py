"""jstacs_b200 — B200-native (sm_100a) engine for the Jstacs HMM hot path.

Layout: csrc/ (CUDA kernels + the C ABI of include/jstacs_hmm.h), _native.py (ctypes binding),
hmm.py / data.py (host-side mirror of the reference's operator interface), parallel.py (one process per GPU),
workloads.py (BASELINE.json configs).
"""
from .data import DNA, AlphabetContainer, DataSet, DNAAlphabet, JavaRandom, Sequence, WrongAlphabetException, WrongLengthException  # noqa: F401
from .hmm import (BasicTransitionElement, BaumWelchParameterSet, DifferentiableHigherOrderHMM, DiscreteEmission,  # noqa: F401
                  HigherOrderDiscreteEmission, HigherOrderHMM, HMMFactory, IterationCondition, SilentEmission,
                  SmallDifferenceOfFunctionEvaluationsCondition, TransitionElement, ViterbiParameterSet)

__version__ = "0.1.0"
