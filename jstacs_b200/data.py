"""Host-side mirror of the reference's data layer, reduced to what the HMM hot path reads.

Reference (paths relative to /root/reference):
  de/jstacs/data/alphabets/DNAAlphabet.java:72      A,C,G,T -> 0..3
  de/jstacs/data/sequences/Sequence.java:102,390    discreteVal(pos), getLength()
  de/jstacs/data/DataSet.java:1001,1199,1222        getElementAt, getMaximalElementLength, getNumberOfElements

B200-first difference: a DataSet is stored PACKED (one uint8 code buffer + int64 offsets), the form the
native library consumes (include/jstacs_hmm.h: jh_dataset_create); Sequence objects are views.
"""
from __future__ import annotations

import numpy as np


class WrongAlphabetException(Exception):
    """de/jstacs/data/WrongAlphabetException.java"""


class WrongLengthException(Exception):
    """de/jstacs/data/WrongLengthException.java"""


class DNAAlphabet:
    """de/jstacs/data/alphabets/DNAAlphabet.java:72 — symbols A,C,G,T with codes 0..3."""

    SYMBOLS = "ACGT"
    _LUT = None

    def length(self) -> int:
        return 4

    def getCode(self, sym: str) -> int:
        i = self.SYMBOLS.find(sym.upper())
        if i < 0:
            raise WrongAlphabetException(f"Symbol \"{sym}\" from input not defined in alphabet.")
        return i

    def getSymbolAt(self, code: int) -> str:
        return self.SYMBOLS[code]

    @classmethod
    def encode(cls, text: str) -> np.ndarray:
        if cls._LUT is None:
            lut = np.full(256, 255, dtype=np.uint8)
            for i, ch in enumerate(cls.SYMBOLS):
                lut[ord(ch)] = i
                lut[ord(ch.lower())] = i
            cls._LUT = lut
        codes = cls._LUT[np.frombuffer(text.encode("ascii"), dtype=np.uint8)]
        if codes.size and codes.max() == 255:
            raise WrongAlphabetException("Symbol from input not defined in alphabet.")
        return codes


class DiscreteAlphabet:
    """de/jstacs/data/alphabets/DiscreteAlphabet.java — integer alphabet of a given size."""

    def __init__(self, size: int):
        self._n = int(size)

    def length(self) -> int:
        return self._n


class AlphabetContainer:
    """de/jstacs/data/AlphabetContainer.java — only simple, discrete containers reach the HMM path
    (AbstractConditionalDiscreteEmission.java:266-268)."""

    def __init__(self, alphabet=None):
        self.alphabet = alphabet if alphabet is not None else DNAAlphabet()

    def getAlphabetLengthAt(self, pos: int) -> int:
        return self.alphabet.length()

    def isDiscrete(self) -> bool:
        return True

    def isSimple(self) -> bool:
        return True

    def checkConsistency(self, other: "AlphabetContainer") -> bool:
        return self.getAlphabetLengthAt(0) == other.getAlphabetLengthAt(0)


DNA = AlphabetContainer(DNAAlphabet())


class Sequence:
    """View of one sequence (Sequence.java:102 discreteVal, :390 getLength)."""

    __slots__ = ("con", "codes")

    def __init__(self, con: AlphabetContainer, codes: np.ndarray):
        self.con = con
        self.codes = np.ascontiguousarray(codes, dtype=np.uint8)

    @staticmethod
    def create(con: AlphabetContainer, text: str) -> "Sequence":
        """Sequence.create(con, String) (Sequence.java:635-659) for a DNA container."""
        return Sequence(con, DNAAlphabet.encode(text))

    def discreteVal(self, pos: int) -> int:
        return int(self.codes[pos])

    def getLength(self) -> int:
        return int(self.codes.shape[0])

    def getAlphabetContainer(self) -> AlphabetContainer:
        return self.con

    def __len__(self):
        return self.getLength()


class DataSet:
    """Packed DataSet: uint8 codes + int64 offsets (sequence n = codes[offsets[n]:offsets[n+1]])."""

    def __init__(self, annotation: str, *seqs: Sequence, con: AlphabetContainer | None = None):
        self.annotation = annotation
        self.con = con or (seqs[0].con if seqs else DNA)
        lens = np.fromiter((s.getLength() for s in seqs), dtype=np.int64, count=len(seqs))
        self.offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum(lens, out=self.offsets[1:])
        self.codes = (np.concatenate([s.codes for s in seqs]) if seqs else np.zeros(0, np.uint8)).astype(np.uint8, copy=False)
        self._native = None  # cached jh_dataset handle (owned by hmm._NativeData)
        self.allowedStatesGroups = None  # int32 per residue, -1 = unconstrained (AbstractHMM.AllowedStatesGroups annotations)

    @classmethod
    def from_packed(cls, codes: np.ndarray, offsets: np.ndarray, con: AlphabetContainer | None = None,
                    annotation: str = "packed") -> "DataSet":
        ds = cls.__new__(cls)
        ds.annotation = annotation
        ds.con = con or DNA
        ds.codes = np.ascontiguousarray(codes, dtype=np.uint8)
        ds.offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        ds._native = None
        ds.allowedStatesGroups = None
        if ds.offsets.ndim != 1 or ds.offsets.size < 1 or ds.offsets[0] != 0 or ds.offsets[-1] != ds.codes.size:
            raise ValueError("offsets must start at 0 and end at len(codes)")
        if ds.codes.size and int(ds.codes.max()) >= ds.con.getAlphabetLengthAt(0):
            raise WrongAlphabetException("code outside the alphabet")
        return ds

    @classmethod
    def from_fasta(cls, path: str, con: AlphabetContainer | None = None) -> "DataSet":
        """DataSet(con, SparseStringExtractor(file, '>')) for plain ACGT FASTA files."""
        con = con or DNA
        seqs, cur = [], []
        with open(path) as fh:
            for line in fh:
                line = line.strip()
                if not line:
                    continue
                if line.startswith(">"):
                    if cur:
                        seqs.append(Sequence.create(con, "".join(cur)))
                        cur = []
                else:
                    cur.append(line)
        if cur:
            seqs.append(Sequence.create(con, "".join(cur)))
        return cls(path, *seqs, con=con)

    @staticmethod
    def from_fasta_stream(path: str, shard_dir: str, **kw):
        """Genome-scale FASTA: streamed once into packed shard files (2 bits per residue by default) and used from there
        (jstacs_b200/shards.py) — returns a ShardedDataSet."""
        from .shards import fasta_to_shards

        return fasta_to_shards(path, shard_dir, **kw)

    def getAlphabetContainer(self) -> AlphabetContainer:
        return self.con

    def getNumberOfElements(self) -> int:
        return int(self.offsets.size - 1)

    def getElementAt(self, n: int) -> Sequence:
        return Sequence(self.con, self.codes[self.offsets[n]:self.offsets[n + 1]])

    def getMaximalElementLength(self) -> int:
        return int(np.diff(self.offsets).max()) if self.getNumberOfElements() else 0

    def getMinimalElementLength(self) -> int:
        return int(np.diff(self.offsets).min()) if self.getNumberOfElements() else 0

    def getElementLength(self) -> int:
        """DataSet.getElementLength: common length, 0 if the lengths differ."""
        if self.getNumberOfElements() == 0:
            return 0
        d = np.diff(self.offsets)
        return int(d[0]) if (d == d[0]).all() else 0

    def getNumberOfResidues(self) -> int:
        return int(self.offsets[-1])

    def shard(self, rank: int, world: int, policy: str = "contiguous") -> "DataSet":
        """Shard balanced by RESIDUE count, not by sequence count like the reference's thread partition
        (HigherOrderHMM.java:1241-1248).  policy "contiguous": a range of sequences; "lpt": longest-processing-time bin
        packing (few very long sequences)."""
        from .parallel import shard_bounds, shard_lpt

        if policy == "lpt":
            idx = shard_lpt(self.offsets, rank, world)
            lens = np.diff(self.offsets)[idx]
            off = np.zeros(idx.size + 1, dtype=np.int64)
            np.cumsum(lens, out=off[1:])
            codes = np.concatenate([self.codes[self.offsets[i]:self.offsets[i + 1]] for i in idx]) if idx.size else np.zeros(0, np.uint8)
            return DataSet.from_packed(codes, off, self.con, f"{self.annotation}[lpt {rank}/{world}]")
        if policy != "contiguous":
            raise ValueError("unknown sharding policy")
        lo, hi = shard_bounds(self.offsets, rank, world)
        off = self.offsets[lo:hi + 1] - self.offsets[lo]
        return DataSet.from_packed(self.codes[self.offsets[lo]:self.offsets[hi]], off, self.con,
                                   f"{self.annotation}[{rank}/{world}]")


class JavaRandom:
    """java.util.Random (48-bit LCG), bit-exact: the synthetic workloads of BASELINE.md §3 are defined
    on this stream so a Java harness and this repo see byte-identical data without shipping files."""

    MULT = 0x5DEECE66D
    MASK = (1 << 48) - 1

    def __init__(self, seed: int):
        self.seed = (seed ^ self.MULT) & self.MASK

    def next(self, bits: int) -> int:
        self.seed = (self.seed * self.MULT + 0xB) & self.MASK
        r = self.seed >> (48 - bits)  # (int)(seed >>> (48 - bits))
        if r >= 1 << 31:
            r -= 1 << 32
        return r

    def nextInt(self, bound: int) -> int:
        if bound & (-bound) == bound:
            return (bound * self.next(31)) >> 31
        while True:
            bits = self.next(31)
            val = bits % bound
            if bits - val + (bound - 1) < (1 << 31):
                return val

    def nextDouble(self) -> float:
        return ((self.next(26) << 27) + self.next(27)) * (1.0 / (1 << 53))


def java_random_dna(n_seq: int, length, first_seed: int = 0x5EED0000, start_index: int = 0) -> DataSet:
    """Sequence n uses Random(0x5EED0000+n).nextInt(4) per residue, vectorised:
    for a power-of-two bound nextInt(4) = top 2 of the 31 returned bits = bits 47..46 of the state."""
    lens = np.full(n_seq, length, dtype=np.int64) if np.isscalar(length) else np.asarray(length, dtype=np.int64)
    offsets = np.zeros(n_seq + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    codes = np.empty(int(offsets[-1]), dtype=np.uint8)
    mult = np.uint64(JavaRandom.MULT)
    mask = np.uint64(JavaRandom.MASK)
    add = np.uint64(0xB)
    max_len = int(lens.max()) if n_seq else 0
    if n_seq and n_seq * 64 < max_len:  # few long sequences: vectorise ALONG each sequence with the LCG jump-ahead
        B = 1 << 20                     # state_{p+k} = state_p * A^k + c * (1 + A + ... + A^(k-1))   (mod 2^48, computed mod 2^64)
        apow = np.multiply.accumulate(np.concatenate([[np.uint64(1)], np.full(B, mult, dtype=np.uint64)]))  # A^0 .. A^B
        cacc = np.concatenate([[np.uint64(0)], np.add.accumulate(apow[:-1]) * add])                          # c * sum_{i<k} A^i
        for n in range(n_seq):
            state = ((np.uint64(n + start_index) + np.uint64(first_seed)) ^ mult) & mask
            L = int(lens[n])
            for p0 in range(0, L, B):
                k = min(B, L - p0)
                st = (state * apow[1:k + 1] + cacc[1:k + 1]) & mask  # states after 1 .. k steps
                codes[offsets[n] + p0:offsets[n] + p0 + k] = (st >> np.uint64(46)).astype(np.uint8)
                state = st[-1]
        return DataSet.from_packed(codes, offsets, DNA, f"javaRandom({n_seq})")
    # process sequences in blocks, stepping all LCGs of a block in lock step
    block = 1 << 16
    for b0 in range(0, n_seq, block):
        b1 = min(n_seq, b0 + block)
        seeds = (np.arange(b0 + start_index, b1 + start_index, dtype=np.uint64) + np.uint64(first_seed)) ^ mult
        state = seeds & mask
        bl = lens[b0:b1]
        bmax = int(bl.max())
        if (bl == bmax).all():
            out = codes[offsets[b0]:offsets[b1]].reshape(b1 - b0, bmax)
            for p in range(bmax):
                state = (state * mult + add) & mask
                out[:, p] = (state >> np.uint64(46)).astype(np.uint8)
        else:
            for p in range(bmax):
                state = (state * mult + add) & mask
                act = np.nonzero(bl > p)[0]
                codes[offsets[b0 + act] + p] = (state[act] >> np.uint64(46)).astype(np.uint8)
    del max_len
    return DataSet.from_packed(codes, offsets, DNA, f"javaRandom({n_seq})")
