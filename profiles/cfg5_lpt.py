"""BASELINE configs[4] (128 states, band 4, the 24 GRCh38 chromosome lengths, 3.085e9 residues) on N GPUs: every rank takes
the sequences that longest-processing-time packing gives it (jstacs_b200/parallel.py: shard_lpt), one Baum-Welch E-step with
the count all-reduce.  usage: torchrun --nproc-per-node N profiles/cfg5_lpt.py [scale]   (scale < 1 shrinks the lengths)
Prints one JSON line on rank 0: per-rank residues and device times, the job time (max over ranks)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import __graft_entry__ as g

g.build()
from jstacs_b200 import _native
from jstacs_b200.parallel import broadcast_comm_id, shard_lpt
from jstacs_b200.workloads import banded_hmm, make_data

GRCH38_MB = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57]
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
L = _native.lib()
_native.check(L.jh_set_device(local))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lens = np.array([int(mb * 1_000_000 * scale) for mb in GRCH38_MB], dtype=np.int64)
offsets = np.concatenate(([0], np.cumsum(lens)))
mine = shard_lpt(offsets, rank, world)
hmm = banded_hmm(128, 4, 0)
data = make_data(len(mine), lens[mine], start_index=int(mine[0]) if len(mine) else 0)
eng, nd = hmm._engine(), hmm._data(data)
if world > 1:
    import ctypes as C

    def make_id():
        buf = (C.c_uint8 * 128)()
        _native.check(L.jh_comm_unique_id(buf))
        return bytes(buf)

    cid = broadcast_comm_id(make_id, device=torch.device("cuda", local))
    _native.check(L.jh_comm_init(eng.h, (C.c_uint8 * 128).from_buffer_copy(cid), world, rank))


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


eng.estep(nd, fetch=False)  # warm-up (scratch, streams)
times = []
for _ in range(2):
    barrier()
    t0 = time.perf_counter()
    eng.estep(nd, fetch=False)
    _native.check(L.jh_synchronize())
    mine_ms = (time.perf_counter() - t0) * 1e3
    kern_ms = eng.last_main_kernel_ms()
    barrier()
    job_ms = (time.perf_counter() - t0) * 1e3
    times.append((job_ms, mine_ms, kern_ms))
job_ms, mine_ms, kern_ms = min(times)
rec = torch.tensor([float(lens[mine].sum()), mine_ms, kern_ms, job_ms], dtype=torch.float64, device="cuda")
if world > 1:
    allr = [torch.zeros_like(rec) for _ in range(world)]
    dist.all_gather(allr, rec)
else:
    allr = [rec]
if rank == 0:
    rows = [[float(x) for x in t.cpu()] for t in allr]
    total = float(lens.sum())
    job = max(r[3] for r in rows)
    print(json.dumps({"config": "cfg5_lpt", "n_gpus": world, "scale": scale, "residues": total, "states": 128,
                      "job_ms": round(job, 1), "residues_states_per_s": total * 128 / (job * 1e-3),
                      "per_rank": [{"residues": r[0], "local_estep_ms": round(r[1], 1), "main_kernel_ms": round(r[2], 1)} for r in rows],
                      "note": "one Baum-Welch E-step incl. the count all-reduce; job time = max over ranks (barrier to barrier)"}), flush=True)
if world > 1:
    dist.destroy_process_group()
