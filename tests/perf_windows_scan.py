"""The NHMMer call pattern (projects/xanthogenomes/NHMMer.java:225-244): a profile HMM with silent states — the PLAN7 model
parsed from the reference's repeats.hmm (tests/golden/profile_repeats.json, 105 states) — scores EVERY window of a fixed
width along a sequence, hmm.getLogProbFor(seq, i, i + w - 1) for all i, here as one jh_loglik_windows call.
usage: python tests/perf_windows_scan.py [sequence length] [window width]
(lives under tests/ because it checks a sample of the windows against the oracle, which only tests may load)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np

import __graft_entry__ as g

g.build()
from jstacs_b200 import _native
from jstacs_b200.workloads import make_data
from test_profile_hmm import fixture_model  # builds the model from the committed fixture
from oracle import binding

L = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 102
fx, hmm, seqs, _ = fixture_model()
data = make_data(1, L)
eng, nd = hmm._engine(), hmm._data(data)
n = L - W + 1
start = np.arange(n, dtype=np.int64)
idx = np.zeros(n, dtype=np.int64)
eng.loglik_windows(nd, idx[:1000], start[:1000], start[:1000] + W - 1)  # warm-up
t0 = time.perf_counter()
ll = eng.loglik_windows(nd, idx, start, start + W - 1)
_native.check(_native.lib().jh_synchronize())
wall = time.perf_counter() - t0
kern = eng.last_kernel_ms()
# the oracle (single thread) on a sample of the windows: parity and the CPU time per window
om = binding.OracleModel(hmm.ir())
x = data.codes
sample = np.linspace(0, n - 1, 200).astype(np.int64)
t0 = time.perf_counter()
ref = np.array([om.loglik_window(x, int(s), int(s) + W - 1) for s in sample])
cpu_per_window = (time.perf_counter() - t0) / sample.size
fin = np.isfinite(ref)
assert np.array_equal(np.isfinite(ll[sample]), fin)
err = float(np.max(np.abs(ll[sample][fin] - ref[fin]) / np.abs(ref[fin]))) if fin.any() else 0.0
print(json.dumps({"config": "nhmmer_window_scan", "states": hmm.getNumberOfStates(), "sequence_length": L, "window": W, "windows": n,
                  "kernel_ms": round(kern, 2), "call_ms": round(wall * 1e3, 2), "windows_per_s": n / (kern * 1e-3),
                  "window_residues_per_s": n * W / (kern * 1e-3), "oracle_cpu_us_per_window_1_thread": round(cpu_per_window * 1e6, 1),
                  "max_rel_err_vs_oracle_on_200_windows": err}), flush=True)
