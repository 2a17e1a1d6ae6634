"""Wire/disk formats around the HMM path: genome-scale FASTA -> packed shard files -> device, streamed.

The reference keeps DNA either one byte per symbol (ByteSequence) or 2 bits per symbol, 32 symbols per long
(de/jstacs/data/sequences/SparseSequence.java:59-73), and reads FASTA through a line extractor
(de/jstacs/io/SparseStringExtractor).  Here a FASTA file is read ONCE, line by line with bounded memory, into shard
files of packed residues:

  <dir>/manifest.json                      alphabet, packing, totals, one record per shard, sequence names
  <dir>/shard_00000.bin                    residues of whole sequences: uint8 codes, or 2 bits per residue (four per
                                           byte, lowest bits first, every sequence starting on a byte boundary)
  <dir>/shard_00000.len                    int64 length of every sequence of the shard

A `ShardedDataSet` memory-maps the shards and hands them one at a time to the library's streamed constructor
(jh_dataset_builder_*: pageable mmap -> pinned staging buffers -> device, 2-bit shards are widened on the device), so a
3 Gb genome is trained from disk without ever holding 3 GB of codes in a host array.
"""
from __future__ import annotations

import json
import os

import numpy as np

from .data import DNA, AlphabetContainer, DataSet, WrongAlphabetException

_LUT = np.full(256, 255, dtype=np.uint8)
for _i, _ch in enumerate("ACGT"):
    _LUT[ord(_ch)] = _i
    _LUT[ord(_ch.lower())] = _i


def pack2(codes: np.ndarray) -> np.ndarray:
    """uint8 codes 0..3 -> 2 bits per residue, four per byte, lowest bits first (the last byte is zero padded)."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    n = codes.size
    pad = (-n) % 4
    if pad:
        codes = np.concatenate([codes, np.zeros(pad, np.uint8)])
    q = codes.reshape(-1, 4)
    return (q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)).astype(np.uint8)


def unpack2(packed: np.ndarray, n: int) -> np.ndarray:
    packed = np.asarray(packed, dtype=np.uint8)
    out = np.empty(packed.size * 4, dtype=np.uint8)
    for k in range(4):
        out[k::4] = (packed >> (2 * k)) & 3
    return out[:n]


class _ShardWriter:
    def __init__(self, out_dir: str, shard_residues: int, packed2: bool):
        self.dir, self.limit, self.packed2 = out_dir, int(shard_residues), packed2
        self.shards, self.names = [], []
        self.fh = None
        self.lens = []
        self.res_in_shard = 0
        self.cur_len = 0          # residues of the sequence being written
        self.carry = np.zeros(0, np.uint8)  # residues of the current sequence not yet flushed (< 4 when packing)

    def _open(self):
        if self.fh is None:
            self.path = os.path.join(self.dir, f"shard_{len(self.shards):05d}.bin")
            self.fh = open(self.path, "wb")
            self.lens, self.res_in_shard = [], 0

    def _close(self):
        if self.fh is None:
            return
        self.fh.close()
        np.asarray(self.lens, dtype=np.int64).tofile(self.path[:-4] + ".len")
        self.shards.append({"file": os.path.basename(self.path), "n_seq": len(self.lens), "n_res": int(sum(self.lens)),
                            "bytes": os.path.getsize(self.path)})
        self.fh = None

    def add(self, codes: np.ndarray):
        """residues of the current sequence (any chunking)"""
        if codes.size == 0:
            return
        self._open()
        self.cur_len += codes.size
        if not self.packed2:
            self.fh.write(codes.tobytes())
            return
        buf = np.concatenate([self.carry, codes]) if self.carry.size else codes
        full = buf.size - buf.size % 4
        if full:
            self.fh.write(pack2(buf[:full]).tobytes())
        self.carry = buf[full:].copy()

    def end_sequence(self, name: str):
        if self.cur_len == 0:
            return
        if self.packed2 and self.carry.size:
            self.fh.write(pack2(self.carry).tobytes())  # every sequence starts on a byte boundary
        self.carry = np.zeros(0, np.uint8)
        self.lens.append(self.cur_len)
        self.names.append(name)
        self.res_in_shard += self.cur_len
        self.cur_len = 0
        if self.res_in_shard >= self.limit:
            self._close()

    def finish(self, con_len: int):
        self._close()
        man = {"format": "jstacs_b200 shards v1", "alphabet": con_len, "packed2": bool(self.packed2),
               "n_seq": int(sum(s["n_seq"] for s in self.shards)), "n_res": int(sum(s["n_res"] for s in self.shards)),
               "shards": self.shards, "names": self.names}
        with open(os.path.join(self.dir, "manifest.json"), "w") as f:
            json.dump(man, f)
        return man


def fasta_to_shards(path: str, out_dir: str, shard_residues: int = 64 << 20, packed2: bool = True, on_unknown: str = "error",
                    min_length: int = 1, block_bytes: int = 16 << 20) -> "ShardedDataSet":
    """Streams a DNA FASTA file into shard files with bounded memory (one block of `block_bytes` at a time).
    on_unknown: "error" — a symbol outside ACGT raises WrongAlphabetException (what DNAAlphabet does);
                "split" — runs of other symbols (N, IUPAC codes) end the current sequence, the next ACGT run starts a new one
                (named name:start), pieces shorter than min_length are dropped."""
    if on_unknown not in ("error", "split"):
        raise ValueError("on_unknown must be 'error' or 'split'")
    os.makedirs(out_dir, exist_ok=True)
    w = _ShardWriter(out_dir, shard_residues, packed2)
    name, pos = None, 0          # current record name, residues consumed of it (incl. unknown symbols)
    piece_start = 0
    pending = []                 # residues of the current piece not yet known to reach min_length
    pending_len = 0

    def flush_piece(final: bool):
        nonlocal pending, pending_len
        total = w.cur_len + pending_len
        if total >= min_length or w.cur_len > 0:
            for c in pending:
                w.add(c)
            pending, pending_len = [], 0
        if final:
            if w.cur_len >= min_length:
                w.end_sequence(name if piece_start == 0 else f"{name}:{piece_start}")
            else:
                assert w.cur_len == 0
            pending, pending_len = [], 0

    def feed(codes: np.ndarray):
        """codes of consecutive residues of the current record (255 = unknown symbol)"""
        nonlocal pos, piece_start, pending, pending_len
        if codes.size == 0:
            return
        good = codes != 255
        if good.all():
            runs = [(0, codes.size)]
        else:
            if on_unknown == "error":
                raise WrongAlphabetException("Symbol from input not defined in alphabet.")
            d = np.diff(np.concatenate([[0], good.astype(np.int8), [0]]))
            runs = list(zip(np.flatnonzero(d == 1).tolist(), np.flatnonzero(d == -1).tolist()))
        cur = 0
        for s0, e0 in runs:
            if s0 > cur:  # unknown symbols in between: the running piece ends, a new one starts at s0
                if w.cur_len + pending_len > 0:
                    flush_piece(True)
                piece_start = pos + s0
            pending.append(codes[s0:e0])
            pending_len += e0 - s0
            if w.cur_len + pending_len >= min_length:
                flush_piece(False)
            cur = e0
        if cur < codes.size:  # the block ends inside a run of unknown symbols
            if w.cur_len + pending_len > 0:
                flush_piece(True)
            piece_start = pos + codes.size
        pos += codes.size

    with open(path, "rb") as fh:
        tail = b""
        while True:
            blk = fh.read(block_bytes)
            if not blk:
                break
            blk = tail + blk
            cut = blk.rfind(b"\n")
            if cut < 0:
                tail = blk
                continue
            tail, blk = blk[cut + 1:], blk[:cut + 1]
            p = 0
            while p < len(blk):
                h = blk.find(b">", p)
                seg_end = len(blk) if h < 0 else h
                if seg_end > p and name is not None:
                    raw = np.frombuffer(blk, dtype=np.uint8, count=seg_end - p, offset=p)
                    raw = raw[(raw != 10) & (raw != 13) & (raw != 32) & (raw != 9)]
                    feed(_LUT[raw])
                if h < 0:
                    break
                # a header line: finish the previous record, start the next one
                if name is not None:
                    flush_piece(True)
                eol = blk.find(b"\n", h)
                name = blk[h + 1:eol].decode("ascii", "replace").strip()
                pos, piece_start = 0, 0
                p = eol + 1
        if tail.strip():
            if tail.startswith(b">"):
                if name is not None:
                    flush_piece(True)
                name, pos, piece_start = tail[1:].decode("ascii", "replace").strip(), 0, 0
            elif name is not None:
                raw = np.frombuffer(tail, dtype=np.uint8)
                raw = raw[(raw != 10) & (raw != 13) & (raw != 32) & (raw != 9)]
                feed(_LUT[raw])
    if name is not None:
        flush_piece(True)
    w.finish(4)
    return ShardedDataSet(out_dir)


def dataset_to_shards(data: DataSet, out_dir: str, shard_residues: int = 64 << 20, packed2: bool = True) -> "ShardedDataSet":
    """An in-memory DataSet -> shard files (tests, conversions)."""
    os.makedirs(out_dir, exist_ok=True)
    if packed2 and data.getAlphabetContainer().getAlphabetLengthAt(0) > 4:
        raise ValueError("2-bit packing needs an alphabet of at most 4 symbols")
    w = _ShardWriter(out_dir, shard_residues, packed2)
    for i in range(data.getNumberOfElements()):
        w.add(data.codes[data.offsets[i]:data.offsets[i + 1]])
        w.end_sequence(f"seq{i}")
    w.finish(data.getAlphabetContainer().getAlphabetLengthAt(0))
    return ShardedDataSet(out_dir, con=data.getAlphabetContainer())


class ShardedDataSet:
    """A data set that lives in shard files; quacks like DataSet for the native-backed model methods (train,
    getLogScoreFor, getViterbiPathsFor, decodePosteriorPathsFor).  The residues are memory-mapped per shard and streamed
    to the device; `materialize()` builds an ordinary DataSet (tests, small data)."""

    def __init__(self, directory: str, con: AlphabetContainer | None = None):
        self.dir = directory
        with open(os.path.join(directory, "manifest.json")) as f:
            self.manifest = json.load(f)
        if not str(self.manifest.get("format", "")).startswith("jstacs_b200 shards"):
            raise ValueError("not a shard directory")
        self.con = con or DNA
        if self.con.getAlphabetLengthAt(0) != self.manifest["alphabet"]:
            raise WrongAlphabetException("The AlphabetContainer does not match the shard files.")
        self.annotation = directory
        self.names = self.manifest["names"]
        lens = [np.fromfile(os.path.join(directory, s["file"][:-4] + ".len"), dtype=np.int64) for s in self.manifest["shards"]]
        self._lens = lens
        all_lens = np.concatenate(lens) if lens else np.zeros(0, np.int64)
        self.offsets = np.zeros(all_lens.size + 1, dtype=np.int64)
        np.cumsum(all_lens, out=self.offsets[1:])
        if all_lens.size != self.manifest["n_seq"] or int(self.offsets[-1]) != self.manifest["n_res"]:
            raise ValueError("manifest and length files disagree")
        self._native = None

    # --- DataSet surface -------------------------------------------------------------------------------------------
    def getAlphabetContainer(self) -> AlphabetContainer:
        return self.con

    def getNumberOfElements(self) -> int:
        return int(self.offsets.size - 1)

    def getNumberOfResidues(self) -> int:
        return int(self.offsets[-1])

    def getMaximalElementLength(self) -> int:
        return int(np.diff(self.offsets).max()) if self.getNumberOfElements() else 0

    def getElementLength(self) -> int:
        d = np.diff(self.offsets)
        return int(d[0]) if d.size and (d == d[0]).all() else 0

    @property
    def packed2(self) -> bool:
        return bool(self.manifest["packed2"])

    def iter_shards(self):
        """(memory-mapped bytes, lengths, packed2) per shard."""
        for s, lens in zip(self.manifest["shards"], self._lens):
            path = os.path.join(self.dir, s["file"])
            mm = np.memmap(path, dtype=np.uint8, mode="r") if s["bytes"] else np.zeros(0, np.uint8)
            yield mm, lens, self.packed2

    def _make_native(self, alphabet: int):
        from . import _native

        return _native.NativeData.from_shards(self.iter_shards(), self.getNumberOfElements(), self.getNumberOfResidues(), None, alphabet)

    def materialize(self) -> DataSet:
        parts = []
        for mm, lens, p2 in self.iter_shards():
            if not p2:
                parts.append(np.asarray(mm))
                continue
            o = 0
            for L in lens:
                nb = (int(L) + 3) // 4
                parts.append(unpack2(np.asarray(mm[o:o + nb]), int(L)))
                o += nb
        codes = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
        return DataSet.from_packed(codes, self.offsets.copy(), self.con, self.annotation)
