"""Multi-rank host logic on CPU (gloo, world_size 2): residue-balanced sharding, ONE all-reduce of the
[transition | emission | score] statistics block per EM iteration, M-step replicated on every rank — the
reference's Compute.oneIteration / estimateFromStatistics (HigherOrderHMM.java:1250-1288) across processes.

The E-step of each rank is done by the oracle here (tests may use it; no GPU in this suite); on the GPU box the same
sharding + collective + M-step sequence runs inside jh_baum_welch with ncclAllReduce (tests/test_gpu_parity.py and
bench.py --gpus N cover that leg).
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from jstacs_b200.parallel import shard_bounds, shard_lpt  # noqa: E402
from jstacs_b200.workloads import ergodic_hmm, make_data, ragged_lengths  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_bounds_partition_and_balance():
    lens = ragged_lengths(5000, seed=7, lo=100.0, decades=3.0)
    offsets = np.zeros(lens.size + 1, np.int64)
    np.cumsum(lens, out=offsets[1:])
    for world in (1, 2, 3, 4, 8):
        cuts = [shard_bounds(offsets, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == lens.size
        for a, b in zip(cuts[:-1], cuts[1:]):
            assert a[1] == b[0]  # contiguous, nothing lost, nothing twice
        res = np.array([offsets[hi] - offsets[lo] for lo, hi in cuts], dtype=np.float64)
        assert res.sum() == offsets[-1]
        # balanced by residues to within one (longest) sequence
        assert res.max() - res.min() <= 2 * lens.max()


def test_shard_bounds_more_ranks_than_sequences():
    offsets = np.array([0, 10, 20], np.int64)
    seen = []
    for r in range(4):
        lo, hi = shard_bounds(offsets, r, 4)
        assert 0 <= lo <= hi <= 2
        seen += list(range(lo, hi))
    assert seen == [0, 1]
    with pytest.raises(ValueError):
        shard_bounds(offsets, 4, 4)


def test_lpt_sharding_of_chromosome_scale_lengths():
    """cfg5 (GRCh38 chromosome sizes in Mb): LPT bin packing is a partition and balances what a contiguous cut cannot."""
    mb = np.array([248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57], np.int64)
    offsets = np.zeros(mb.size + 1, np.int64)
    np.cumsum(mb, out=offsets[1:])
    for world in (1, 2, 4, 8):
        parts = [shard_lpt(offsets, r, world) for r in range(world)]
        assert sorted(np.concatenate(parts).tolist()) == list(range(mb.size))
        load = np.array([mb[p].sum() for p in parts])
        assert load.max() <= max(mb.max(), 1.1 * mb.sum() / world)  # within 10 % of the mean, or one chromosome
        cont = np.array([offsets[hi] - offsets[lo] for lo, hi in (shard_bounds(offsets, r, world) for r in range(world))])
        assert load.max() <= cont.max()
    data = make_data(5, [30, 10, 50, 20, 40])
    a, b = data.shard(0, 2, policy="lpt"), data.shard(1, 2, policy="lpt")
    assert a.getNumberOfResidues() + b.getNumberOfResidues() == 150 and abs(a.getNumberOfResidues() - b.getNumberOfResidues()) <= 10
    i0 = shard_lpt(data.offsets, 0, 2)
    assert np.array_equal(a.codes[:int(np.diff(data.offsets)[i0[0]])], data.getElementAt(int(i0[0])).codes)


def _worker(rank, world, port, n_iter, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    from jstacs_b200.parallel import allreduce_stats_
    from oracle import binding

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hmm = ergodic_hmm(4, 2, ess=4.0)
        lens = ragged_lengths(40, seed=11, lo=20.0, decades=1.0)
        data = make_data(40, lens)
        weights = 0.5 + (np.arange(40) % 3).astype(np.float64)
        lo, hi = shard_bounds(data.offsets, rank, world)
        shard = data.shard(rank, world)
        om = binding.OracleModel(hmm.ir())
        objective = []
        for it in range(n_iter):
            ts, es, sc = om.estep(shard.codes, shard.offsets, weights=weights[lo:hi])
            block = torch.from_numpy(np.concatenate([ts, es, [sc]]))
            allreduce_stats_(block)  # the ONE collective of an EM iteration
            block = block.numpy()
            objective.append(om.log_prior() + block[-1])
            if it + 1 < n_iter:  # N E-steps, N-1 M-steps (HigherOrderHMM.java:760-770)
                om.mstep(block[:ts.size], block[ts.size:-1])  # replicated M-step: no broadcast
        tp, tl, ep, el = om.get_params()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), tp=tp, ep=ep, obj=np.array(objective), n=hi - lo)
    finally:
        dist.destroy_process_group()


def test_two_rank_em_equals_single_process(tmp_path, oracle_lib):
    import torch.multiprocessing as mp

    n_iter, world = 4, 2
    mp.spawn(_worker, args=(world, _free_port(), n_iter, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{i}.npz") for i in range(world)]
    assert int(r[0]["n"]) + int(r[1]["n"]) == 40 and int(r[0]["n"]) > 0 and int(r[1]["n"]) > 0
    # every rank ends with the same parameters (replicated deterministic M-step) ...
    assert np.array_equal(r[0]["tp"], r[1]["tp"]) and np.array_equal(r[0]["ep"], r[1]["ep"])
    assert np.array_equal(r[0]["obj"], r[1]["obj"])
    # ... and they are the single-process result up to the summation order of the statistics (~1e-13)
    hmm = ergodic_hmm(4, 2, ess=4.0)
    lens = ragged_lengths(40, seed=11, lo=20.0, decades=1.0)
    data = make_data(40, lens)
    weights = 0.5 + (np.arange(40) % 3).astype(np.float64)
    om = oracle_lib.OracleModel(hmm.ir())
    obj = om.train(data.codes, data.offsets, weights=weights, max_iter=n_iter)
    tp, tl, ep, el = om.get_params()
    np.testing.assert_allclose(r[0]["obj"], obj, rtol=1e-12)
    np.testing.assert_allclose(r[0]["tp"], tp, rtol=0, atol=1e-10)
    np.testing.assert_allclose(r[0]["ep"], ep, rtol=0, atol=1e-10)
