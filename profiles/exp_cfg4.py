"""Scratch experiment: 16-state E-step on uniform-length vs ragged data (tail / makespan check)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
g.build()
from jstacs_b200 import _native
from jstacs_b200.workloads import ergodic_hmm, make_data

import ctypes
if os.environ.get("EXP_STREAM") == "legacy":
    _native.check(_native.lib().jh_set_stream(ctypes.c_void_p(1)))  # cudaStreamLegacy
for kv in filter(None, os.environ.get("JH_OPTS", "").split(",")):
    k, v = kv.split("=")
    _native.check(_native.lib().jh_set_option(k.encode(), int(v)))


def run(name, lens):
    hmm = ergodic_hmm(int(os.environ.get('EXP_STATES', 16)), int(os.environ.get('EXP_ORDER', 1)))
    data = make_data(len(lens), np.asarray(lens))
    eng, nd = hmm._engine(), hmm._data(data)
    eng.estep(nd, fetch=False)
    ms = []
    for _ in range(3):
        eng.estep(nd, fetch=False)
        _native.check(_native.lib().jh_synchronize())
        ms.append(eng.last_main_kernel_ms())
    print(json.dumps({"case": name, "n_seq": len(lens), "residues": int(np.sum(lens)), "kernel_ms": round(float(np.median(ms)), 3)}), flush=True)
    nd.close()

run("uniform_1184oct_x_12500", [12500] * (1184 * 8))
if not os.environ.get("EXP_SHORT"):
    run("uniform_1243oct_x_12500", [12500] * (1243 * 8))
    run("uniform_2368oct_x_12500", [12500] * (2368 * 8))
    run("linear_1243oct_12500_to_6250", np.linspace(12500, 6250, 1243 * 8).astype(np.int64))
