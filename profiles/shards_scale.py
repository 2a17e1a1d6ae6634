"""cfg5-scale data from DISK: a FASTA file with the 24 GRCh38 chromosome lengths (x scale) is written in pieces, streamed
once into 2-bit shard files (jstacs_b200/shards.py), and a Baum-Welch E-step of the 128-state banded model runs on the
sharded data set (jh_dataset_builder_*: mmap -> pinned staging -> device).  Reports the timings and the PEAK RESIDENT SET of
the process — the point being that no host array of the size of the data ever exists.
usage: python profiles/shards_scale.py [scale] [workdir]"""
import json
import os
import resource
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import __graft_entry__ as g

g.build()
from jstacs_b200 import _native
from jstacs_b200.data import DataSet
from jstacs_b200.workloads import banded_hmm

GRCH38_MB = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57]
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
work = sys.argv[2] if len(sys.argv) > 2 else tempfile.mkdtemp(prefix="jh_shards_")
os.makedirs(work, exist_ok=True)
fasta = os.path.join(work, "genome.fa")
rss = lambda: resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1024.0  # MB


t0 = time.perf_counter()
rng = np.random.default_rng(1)
letters = np.frombuffer(b"ACGT", dtype=np.uint8)
total = 0
with open(fasta, "wb") as fh:  # written in 8 MB pieces, 80 columns per line
    for i, mb in enumerate(GRCH38_MB):
        L = int(mb * 1_000_000 * scale)
        total += L
        fh.write(f">chr{i + 1}\n".encode())
        done = 0
        while done < L:
            n = min(8_000_000, L - done)
            piece = letters[rng.integers(0, 4, n, dtype=np.uint8)]
            rows = piece[:n - n % 80].reshape(-1, 80)
            out = np.concatenate([rows, np.full((rows.shape[0], 1), 10, np.uint8)], axis=1).tobytes()
            fh.write(out)
            if n % 80:
                fh.write(piece[n - n % 80:].tobytes() + b"\n")
            done += n
t_write = time.perf_counter() - t0
rss_after_write = rss()

t0 = time.perf_counter()
sharded = DataSet.from_fasta_stream(fasta, os.path.join(work, "shards"))
t_shard = time.perf_counter() - t0
rss_after_shard = rss()
shard_bytes = sum(os.path.getsize(os.path.join(work, "shards", f)) for f in os.listdir(os.path.join(work, "shards")))

hmm = banded_hmm(128, 4, 0)
eng = hmm._engine()
t0 = time.perf_counter()
nd = hmm._data(sharded)  # streamed upload
_native.check(_native.lib().jh_synchronize())
t_upload = time.perf_counter() - t0
rss_after_upload = rss()
eng.estep(nd, fetch=False)  # warm-up
t0 = time.perf_counter()
ts, es, sc = eng.estep(nd)
t_estep = time.perf_counter() - t0
print(json.dumps({"config": "cfg5_from_disk", "scale": scale, "residues": total, "fasta_bytes": os.path.getsize(fasta), "shard_bytes": shard_bytes,
                  "write_fasta_s": round(t_write, 2), "fasta_to_shards_s": round(t_shard, 2), "shards_to_device_s": round(t_upload, 3),
                  "estep_ms": round(t_estep * 1e3, 1), "estep_main_kernel_ms": round(eng.last_main_kernel_ms(), 1),
                  "count_check": float(ts.sum() + 0.0), "score": sc,
                  "peak_rss_mb": {"after_writing_fasta": round(rss_after_write), "after_sharding": round(rss_after_shard),
                                  "after_upload": round(rss_after_upload), "end": round(rss())},
                  "note": "peak resident set of the whole python process (torch is not imported); data size in host memory would be residues bytes"}), flush=True)
if len(sys.argv) <= 2:
    shutil.rmtree(work, ignore_errors=True)
