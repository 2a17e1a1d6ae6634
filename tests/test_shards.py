"""Wire/disk formats: streamed FASTA -> packed shard files -> device (jstacs_b200/shards.py,
jh_dataset_builder_*, jh_dataset_create_packed2)."""
import os

import numpy as np
import pytest

from jstacs_b200 import DataSet, WrongAlphabetException
from jstacs_b200.shards import ShardedDataSet, dataset_to_shards, fasta_to_shards, pack2, unpack2
from jstacs_b200.workloads import ergodic_hmm, make_data


def write_fasta(path, records, width=60, newline="\n", final_newline=True):
    with open(path, "w", newline="") as f:
        for k, (name, seq) in enumerate(records):
            f.write(">" + name + newline)
            for i in range(0, len(seq), width):
                last = k == len(records) - 1 and i + width >= len(seq)
                f.write(seq[i:i + width] + ("" if last and not final_newline else newline))


def random_records(rng, lens, alphabet="ACGT"):
    return [(f"chr{i} test record", "".join(rng.choice(list(alphabet), L))) for i, L in enumerate(lens)]


def test_pack2_round_trip():
    rng = np.random.default_rng(0)
    for n in (0, 1, 3, 4, 5, 17, 1000, 1001, 1002, 1003):
        c = rng.integers(0, 4, n).astype(np.uint8)
        p = pack2(c)
        assert p.size == (n + 3) // 4
        assert np.array_equal(unpack2(p, n), c)
    assert pack2(np.array([1, 2, 3, 0, 3], np.uint8)).tolist() == [1 | (2 << 2) | (3 << 4), 3]


@pytest.mark.parametrize("packed2", [True, False])
@pytest.mark.parametrize("block_bytes", [37, 256, 1 << 20])
def test_fasta_stream_equals_in_memory_reader(tmp_path, packed2, block_bytes):
    rng = np.random.default_rng(1)
    recs = random_records(rng, [1, 59, 60, 61, 500, 3, 1234, 4, 7])
    fa = str(tmp_path / "a.fa")
    write_fasta(fa, recs, final_newline=(block_bytes != 256))
    sh = fasta_to_shards(fa, str(tmp_path / "shards"), shard_residues=100, packed2=packed2, block_bytes=block_bytes)
    ref = DataSet.from_fasta(fa)
    got = sh.materialize()
    assert np.array_equal(got.offsets, ref.offsets) and np.array_equal(got.codes, ref.codes)
    assert sh.names == [n for n, _ in recs]
    assert len(sh.manifest["shards"]) >= 3  # several shards, whole sequences each
    for s in sh.manifest["shards"]:
        assert s["bytes"] == os.path.getsize(os.path.join(sh.dir, s["file"]))
    assert sh.getNumberOfResidues() == ref.getNumberOfResidues() and sh.getMaximalElementLength() == 1234
    again = ShardedDataSet(sh.dir)
    assert np.array_equal(again.offsets, sh.offsets)


def test_fasta_stream_unknown_symbols(tmp_path):
    rng = np.random.default_rng(2)
    a, b, c = ("".join(rng.choice(list("ACGT"), L)) for L in (40, 3, 75))
    recs = [("one", "NN" + a + "NNNNNN" + b + "N" + c), ("two", "acgtacgtnnnn"), ("three", "NNNN")]
    fa = str(tmp_path / "n.fa")
    write_fasta(fa, recs, width=17, newline="\r\n")
    with pytest.raises(WrongAlphabetException):
        fasta_to_shards(fa, str(tmp_path / "err"))
    for block in (11, 1 << 16):
        sh = fasta_to_shards(fa, str(tmp_path / f"split{block}"), on_unknown="split", min_length=4, block_bytes=block, packed2=True)
        ds = sh.materialize()
        assert sh.names == ["one:2", "one:52", "two"]  # the 3-residue piece and the all-N record are dropped
        texts = ["".join("ACGT"[k] for k in ds.getElementAt(i).codes) for i in range(3)]
        assert texts == [a, c, "ACGTACGT"]


def test_dataset_to_shards_round_trip(tmp_path):
    data = make_data(9, [5, 1, 64, 63, 2, 1000, 7, 8, 9])
    for p2 in (True, False):
        sh = dataset_to_shards(data, str(tmp_path / f"d{int(p2)}"), shard_residues=100, packed2=p2)
        m = sh.materialize()
        assert np.array_equal(m.codes, data.codes) and np.array_equal(m.offsets, data.offsets)


@pytest.mark.gpu
@pytest.mark.parametrize("packed2", [True, False])
def test_sharded_data_set_on_the_device_equals_the_in_memory_one(tmp_path, packed2):
    import __graft_entry__ as g

    g.build()
    from jstacs_b200 import _native

    rng = np.random.default_rng(3)
    lens = np.concatenate([rng.integers(1, 40, 20), rng.integers(500, 3000, 6), [70_001, 16, 17, 15]])
    data = make_data(lens.size, lens)
    sh = dataset_to_shards(data, str(tmp_path / "sh"), shard_residues=20_000, packed2=packed2)
    hmm = ergodic_hmm(8, 2, n_iter=3, ess=1.0)
    ll_mem = hmm.getLogScoreFor(data)
    ll_sh = hmm.getLogScoreFor(sh)
    assert np.array_equal(ll_mem, ll_sh)  # the same residues on the device: the same bits out
    v_mem = hmm.getViterbiPathsFor(data)
    v_sh = hmm.getViterbiPathsFor(sh)
    for (p1, s1), (p2_, s2) in zip(v_mem, v_sh):
        assert np.array_equal(p1, p2_) and s1 == s2
    eng = hmm._engine()
    ts1, es1, sc1 = eng.estep(hmm._data(data))
    ts2, es2, sc2 = eng.estep(hmm._data(sh))
    np.testing.assert_allclose(ts1, ts2, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(es1, es2, rtol=1e-12, atol=1e-12)
    # jh_dataset_create_packed2: one call, 2-bit input
    if packed2:
        packed, bp = [], [0]
        for i in range(lens.size):
            p = pack2(data.codes[data.offsets[i]:data.offsets[i + 1]])
            packed.append(p)
            bp.append(bp[-1] + p.size)
        nd = _native.NativeData(np.concatenate(packed), data.offsets, None, 4, byte_ptr=np.array(bp, np.int64))
        assert np.array_equal(eng.loglik(nd), ll_mem)
        nd.close()
    # a code outside the model's alphabet is refused on the device, whatever the route
    bad = data.codes.copy()
    bad[12345] = 3
    hmm2 = ergodic_hmm(2, 0)
    hmm2.alphabets = type(hmm2.alphabets)(__import__("jstacs_b200").data.DiscreteAlphabet(3))
    with pytest.raises(_native.NativeError) as ei:
        _native.NativeData.from_shards([(bad, np.diff(data.offsets), False)], lens.size, int(data.offsets[-1]), None, 3)
    assert ei.value.status == -4
    # announcing fewer residues than delivered is an error, not an overflow
    with pytest.raises(_native.NativeError):
        _native.NativeData.from_shards([(data.codes, np.diff(data.offsets), False)], lens.size, int(data.offsets[-1]) - 1, None, 4)
