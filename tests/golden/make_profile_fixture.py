"""Generates tests/golden/profile_repeats.json from a REAL fixture of the reference:
projects/xanthogenomes/data/repeats.hmm (HMMER3 profile of TALE repeats, 102 nodes), parsed by the mirror of
HMMFactory.parseProfileHMMFromHMMer (jstacs_b200/factory.py) into a PLAN7 profile HMM with 317 states (109 silent).

Stored: the flat parameter vector (reference order), synthetic and consensus-derived test sequences, and their
log-likelihoods computed by an INDEPENDENT method — dense linear algebra with the silent states eliminated in closed form
((I - T_SS)^-1), not the oracle and not the context-graph code — so the fixture pins the oracle and the GPU path for a
realistic silent-state model.  usage (container with /root/reference): python tests/golden/make_profile_fixture.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from jstacs_b200.factory import parseProfileHMMFromHMMer  # noqa: E402

SRC = "/root/reference/projects/xanthogenomes/data/repeats.hmm"


def independent_loglik(hmm, x):
    """log P(x) by matrix algebra: silent states eliminated through (I - T_SS)^-1, scaled forward recursion."""
    n = len(hmm.name)
    T = np.zeros((n, n))
    pi = np.zeros(n)
    for te in hmm.transition.transitions:
        p = np.exp(te.parameters - te.logNorm)
        if len(te.context) == 0:
            for st, v in zip(te.states, p):
                pi[st] += v
        else:
            for st, v in zip(te.states, p):
                T[te.context[-1], st] += v
    silent = np.array([hmm.emission[i].order < 0 for i in hmm.emissionIdx])
    S, E = np.flatnonzero(silent), np.flatnonzero(~silent)
    C = np.linalg.inv(np.eye(S.size) - T[np.ix_(S, S)])
    pi_eff = pi[E] + pi[S] @ C @ T[np.ix_(S, E)]
    T_eff = T[np.ix_(E, E)] + T[np.ix_(E, S)] @ C @ T[np.ix_(S, E)]
    f = np.array(hmm.finalState, dtype=float)
    g = f[E] + T[np.ix_(E, S)] @ C @ f[S]
    em = np.stack([np.exp(hmm.emission[hmm.emissionIdx[i]].params[0] - hmm.emission[hmm.emissionIdx[i]].logNorm[0]) for i in E])  # [state][symbol]
    a = pi_eff * em[:, x[0]]
    ll = 0.0
    for t in range(1, len(x)):
        s = a.sum()
        ll += np.log(s)
        a = (a / s) @ T_eff * em[:, x[t]]
    return ll + np.log(a @ g)


def main():
    consensus = []
    hmm, bg = parseProfileHMMFromHMMer(open(SRC), consensus)
    rng = np.random.default_rng(2024)
    cons = np.array(["acgt".index(c.lower()) for c in consensus], dtype=np.uint8)
    seqs = [cons.copy()]
    for k in range(7):  # consensus with substitutions, deletions and insertions
        s = list(cons)
        for _ in range(3 * k):
            i = int(rng.integers(0, len(s)))
            op = rng.integers(0, 3)
            if op == 0:
                s[i] = int(rng.integers(0, 4))
            elif op == 1 and len(s) > 5:
                del s[i]
            else:
                s.insert(i, int(rng.integers(0, 4)))
        seqs.append(np.array(s, dtype=np.uint8))
    seqs += [rng.integers(0, 4, L).astype(np.uint8) for L in (1, 2, 7, 60, 150, 233)]
    seqs.append(np.concatenate([cons, cons[:40]]))
    out = {"source": "projects/xanthogenomes/data/repeats.hmm via jstacs_b200/factory.py:parseProfileHMMFromHMMer",
           "n_nodes": len(consensus), "n_states": len(hmm.name), "consensus": "".join(consensus),
           "params": [float(v).hex() for v in hmm.getCurrentParameterValues()],
           "sequences": ["".join("ACGT"[c] for c in s) for s in seqs],
           "loglik_independent": [float(independent_loglik(hmm, s)).hex() for s in seqs]}
    with open(os.path.join(ROOT, "tests", "golden", "profile_repeats.json"), "w") as f:
        json.dump(out, f)
    print("states", len(hmm.name), "sequences", len(seqs), [round(independent_loglik(hmm, s), 3) for s in seqs][:5])


if __name__ == "__main__":
    main()
