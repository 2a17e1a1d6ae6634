"""Device-time throughput of every BASELINE.json config shape (kernel time from the library's CUDA events; the
host<->device copies of the C-ABI call are excluded).  usage: python profiles/bench_configs.py [scale]
scale < 1 shrinks the sequence counts (default 0.1).  One JSON line per config."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import __graft_entry__ as g

g.build()
from jstacs_b200 import _native
from jstacs_b200.workloads import banded_hmm, ergodic_hmm, make_data, ragged_lengths

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
only = sys.argv[2:] or None
for kv in filter(None, os.environ.get("JH_OPTS", "").split(",")):  # library options, e.g. JH_OPTS=oct_chunk=256,oct_chunk_force=1
    k, v = kv.split("=")
    _native.check(_native.lib().jh_set_option(k.encode(), int(v)))


_bufs = {}


def _buf(n):
    if n not in _bufs:
        _bufs.clear()
        _bufs[n] = np.zeros(max(n, 1), dtype=np.uint8)
    return _bufs[n]


def run(name, hmm, data, what, reps=3):
    if (only and name not in only) or (name.startswith("full_") and not (only and name in only)):
        return
    hmm = hmm()  # models and data are built lazily: only for the selected lines
    if callable(data):
        data = data()
    eng, nd = hmm._engine(), hmm._data(data)
    res, S = data.getNumberOfResidues(), hmm.getNumberOfStates()
    fn = {"estep": lambda: eng.estep(nd, fetch=False), "loglik": lambda: eng.loglik(nd), "viterbi": lambda: eng.viterbi(nd),
          "decode": lambda: eng.posterior_decode(nd), "viterbi_u8": lambda: eng.viterbi(nd, narrow=True),
          "decode_u8": lambda: eng.posterior_decode(nd, narrow=True),
          # the caller's path buffer reused (already touched pages: no first-touch faults inside the call)
          "viterbi_u8_reuse": lambda: eng.viterbi(nd, narrow=True, out=_buf(nd.n_res)),
          "vit_estep": lambda: eng.estep(nd, mode=_native.MODE_VITERBI, fetch=False),  # E-step of Viterbi training
          "train": lambda: hmm.train(data)}[what]  # the whole HigherOrderHMM.train call (data set resident; parameters read back)
    if not name.startswith("full_"):
        fn()  # warm-up
    ms = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        _native.check(_native.lib().jh_synchronize())
        wall = (time.perf_counter() - t0) * 1e3
        ms.append((eng.last_main_kernel_ms() if what in ("estep", "vit_estep") else wall if what == "train" else eng.last_kernel_ms(), wall))
    k = float(np.median([m[0] for m in ms]))
    print(json.dumps({"config": name, "op": what, "states": S, "n_seq": data.getNumberOfElements(), "residues": res, "kernel_ms": round(k, 3),
                      "call_ms": round(float(np.median([m[1] for m in ms])), 3), "residues_states_per_s": res * S / (k * 1e-3),
                      "residues_per_s": res / (k * 1e-3)}), flush=True)
    nd.close()
    data._native = None


n = lambda x: max(64, int(x * scale))
run("cfg1_bw_s2_k1", lambda: ergodic_hmm(2, 1), lambda: make_data(10_000, 1000), "estep")
run("cfg1_train10_s2_k1", lambda: ergodic_hmm(2, 1, n_iter=10), lambda: make_data(10_000, 1000), "train")
run("cfg1_decode_u8_s2_k1_100k", lambda: ergodic_hmm(2, 1), lambda: make_data(100_000, 1000), "decode_u8")
run("cfg1_viterbi_u8_s2_k1_100k", lambda: ergodic_hmm(2, 1), lambda: make_data(100_000, 1000), "viterbi_u8")
run("cfg2_viterbi_s8_k2", lambda: ergodic_hmm(8, 2), lambda: make_data(n(1_000_000), 1000), "viterbi")
run("cfg2_bw_s8_k2", lambda: ergodic_hmm(8, 2), lambda: make_data(n(1_000_000), 1000), "estep")
run("cfg2_viterbi_training_estep_s8_k2", lambda: ergodic_hmm(8, 2), lambda: make_data(n(1_000_000), 1000), "vit_estep")
run("cfg3_bw_s32_k3", lambda: ergodic_hmm(32, 3, ess=4.0), lambda: make_data(n(1_000_000), 1000), "estep")
run("cfg3_loglik_s32_k3", lambda: ergodic_hmm(32, 3, ess=4.0), lambda: make_data(n(1_000_000), 1000), "loglik")
run("cfg3_decode_s32_k3", lambda: ergodic_hmm(32, 3, ess=4.0), lambda: make_data(n(200_000), 1000), "decode")
_lens = []


def ragged():
    if not _lens:
        _lens.append(ragged_lengths(n(100_000), seed=42))
    return _lens[0]


run("cfg4_loglik_s16_k1_ragged", lambda: ergodic_hmm(16, 1), lambda: make_data(ragged().size, ragged()), "loglik")
run("cfg4_bw_s16_k1_ragged", lambda: ergodic_hmm(16, 1), lambda: make_data(ragged().size, ragged()), "estep")
run("cfg5_loglik_s128_banded_1Mb", lambda: banded_hmm(128, 4, 0), lambda: make_data(24, 1_000_000), "loglik")
run("long_decode_s16_dense_8x2Mb", lambda: ergodic_hmm(16, 1), lambda: make_data(8, 2_000_000), "decode")
run("cfg5_decode_s128_banded_1Mb", lambda: banded_hmm(128, 4, 0), lambda: make_data(24, 1_000_000), "decode")
run("cfg5_bw_s128_banded_1Mb", lambda: banded_hmm(128, 4, 0), lambda: make_data(24, 1_000_000), "estep")
run("long_bw_s16_dense_8x2Mb", lambda: ergodic_hmm(16, 1), lambda: make_data(8, 2_000_000), "estep")
run("cfg5_viterbi_u8_s128_banded_1Mb", lambda: banded_hmm(128, 4, 0), lambda: make_data(24, 1_000_000), "viterbi_u8")
run("cfg4_decode_s16_k1_ragged", lambda: ergodic_hmm(16, 1), lambda: make_data(ragged().size, ragged()), "decode")
run("cfg2_viterbi_u8_s8_k2", lambda: ergodic_hmm(8, 2), lambda: make_data(n(1_000_000), 1000), "viterbi_u8")
run("cfg4_decode_u8_s16_k1_ragged", lambda: ergodic_hmm(16, 1), lambda: make_data(ragged().size, ragged()), "decode_u8")
# small dense models on chromosome-scale sequences: checkpoint pass parallel in the position (kernels_scan.cuh; JH_OPTS=scan=0: sequential)
run("chr_cpg_bw_s2_k1_2x50Mb", lambda: ergodic_hmm(2, 1), lambda: make_data(2, 50_000_000), "estep")
run("chr_cpg_decode_u8_s2_k1_2x50Mb", lambda: ergodic_hmm(2, 1), lambda: make_data(2, 50_000_000), "decode_u8")
run("chr_cpg_loglik_s2_k1_2x50Mb", lambda: ergodic_hmm(2, 1), lambda: make_data(2, 50_000_000), "loglik")
run("chr_decode_u8_s8_k2_4x25Mb", lambda: ergodic_hmm(8, 2), lambda: make_data(4, 25_000_000), "decode_u8")
run("long_bw_s32_k3_dense_8x2Mb", lambda: ergodic_hmm(32, 3, ess=4.0), lambda: make_data(8, 2_000_000), "estep")
run("chr_bw_s8_k2_4x25Mb", lambda: ergodic_hmm(8, 2), lambda: make_data(4, 25_000_000), "estep")
# emission tables larger than shared memory (EG kernels): 32 states order 4 = 256 KB, 8 states order 5 = 256 KB
run("hiorder_bw_s32_k4", lambda: ergodic_hmm(32, 4, ess=4.0), lambda: make_data(n(1_000_000), 1000), "estep")
run("hiorder_bw_s8_k5", lambda: ergodic_hmm(8, 5), lambda: make_data(n(1_000_000), 1000), "estep")
run("hiorder_viterbi_u8_s8_k5", lambda: ergodic_hmm(8, 5), lambda: make_data(n(1_000_000), 1000), "viterbi_u8")
# BASELINE configs[4] at FULL size (GRCh38 chromosome lengths in Mb, 3.085e9 residues); only when named explicitly
GRCH38_MB = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57]
_full5 = []


def full5():
    if not _full5:
        _full5.append(make_data(24, [mb * 1_000_000 for mb in GRCH38_MB]))
    return _full5[0]


mid5 = lambda: make_data(24, [mb * 20_000 for mb in GRCH38_MB])  # the same length profile at 1/50 (61.7e6 residues)
run("full_mid_cfg5_loglik", lambda: banded_hmm(128, 4, 0), mid5, "loglik", reps=3)
run("full_mid_cfg5_bw", lambda: banded_hmm(128, 4, 0), mid5, "estep", reps=3)
run("full_mid_cfg5_decode_u8", lambda: banded_hmm(128, 4, 0), mid5, "decode_u8", reps=3)
run("full_mid_cfg5_viterbi_u8", lambda: banded_hmm(128, 4, 0), mid5, "viterbi_u8", reps=3)
run("full_cfg2_viterbi_u8_s8_k2", lambda: ergodic_hmm(8, 2), lambda: make_data(1_000_000, 1000), "viterbi_u8", reps=3)
run("full_cfg2_viterbi_u8_reuse_s8_k2", lambda: ergodic_hmm(8, 2), lambda: make_data(1_000_000, 1000), "viterbi_u8_reuse", reps=4)
run("full_cfg5_viterbi_u8", lambda: banded_hmm(128, 4, 0), full5, "viterbi_u8", reps=1)
run("full_cfg5_loglik", lambda: banded_hmm(128, 4, 0), full5, "loglik", reps=1)
run("full_cfg5_decode_u8", lambda: banded_hmm(128, 4, 0), full5, "decode_u8", reps=1)
run("full_cfg5_bw", lambda: banded_hmm(128, 4, 0), full5, "estep", reps=1)
