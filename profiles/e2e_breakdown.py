"""Where the end-to-end time of HigherOrderHMM.train(DataSet) goes (bench.py's e2e leg): wall-clock per phase with a
device synchronize after each.  usage: python profiles/e2e_breakdown.py [n_seq]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as g

g.build()
from jstacs_b200 import _native
from jstacs_b200.workloads import ergodic_hmm, make_data

n_seq = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
L = _native.lib()
hmm = ergodic_hmm(32, 3, n_iter=2, ess=4.0)
data = make_data(n_seq, 1000)
pinned = torch.from_numpy(data.codes).pin_memory().numpy()
eng = hmm._engine()
sync = lambda: _native.check(L.jh_synchronize())


def timed(fn):
    sync()
    t0 = time.perf_counter()
    r = fn()
    sync()
    return r, (time.perf_counter() - t0) * 1e3


for rep in range(3):
    out = {}
    nd, out["dataset_create_ms"] = timed(lambda: _native.NativeData(pinned, data.offsets, None, 4))
    _, out["push_params_ms"] = timed(hmm._push_params)
    _, out["estep1_ms"] = timed(lambda: eng.estep(nd, fetch=False))
    out["estep1_kernel_ms"] = eng.last_main_kernel_ms()
    _, out["mstep_ms"] = timed(eng.mstep)
    _, out["estep2_ms"] = timed(lambda: eng.estep(nd, fetch=False))
    _, out["pull_params_ms"] = timed(hmm._pull_params)
    _, out["dataset_destroy_ms"] = timed(nd.close)
    out["total_ms"] = sum(v for k, v in out.items() if k.endswith("_ms") and k != "estep1_kernel_ms")
    print(json.dumps({k: round(v, 2) for k, v in out.items()}), flush=True)
