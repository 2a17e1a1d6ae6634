"""ctypes binding of libjstacs_hmm_b200.so (C ABI: include/jstacs_hmm.h).

The library is the product; this file only marshals numpy buffers.  It fails loudly when the shared object or
a CUDA device is missing — there is no CPU fallback (BASELINE.json north_star).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

MODE_VITERBI, MODE_BAUM_WELCH = 0, 1
TERM_ITERATIONS, TERM_SMALL_DIFFERENCE = 0, 1
ABI_VERSION = 2
COMM_ID_BYTES = 128

# JSTACS_HMM_LIB: another build of the same library (A/B measurements of two builds in one process environment)
LIB_PATH = os.environ.get("JSTACS_HMM_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libjstacs_hmm_b200.so")

_P_I32 = C.POINTER(C.c_int32)
_P_F64 = C.POINTER(C.c_double)
_P_U8 = C.POINTER(C.c_uint8)
_P_I64 = C.POINTER(C.c_int64)


class ModelDesc(C.Structure):
    """struct jh_model_desc (include/jstacs_hmm.h)."""

    _fields_ = [
        ("abi_version", C.c_int32), ("n_states", C.c_int32), ("alphabet", C.c_int32), ("max_order", C.c_int32),
        ("lc_ctx_ptr", _P_I32), ("ctx_elem", _P_I32), ("ctx_last_state", _P_I32),
        ("n_elem", C.c_int32),
        ("elem_child_ptr", _P_I32), ("child_state", _P_I32), ("child_desc", _P_I32), ("elem_ctx", _P_I32),
        ("trans_index", _P_I32),
        ("trans_param", _P_F64), ("trans_lognorm", _P_F64), ("trans_hyper", _P_F64),
        ("n_emis", C.c_int32),
        ("state_emis", _P_I32), ("emis_order", _P_I32), ("emis_cond_ptr", _P_I32),
        ("emis_param", _P_F64), ("emis_lognorm", _P_F64), ("emis_hyper", _P_F64),
        ("state_silent", _P_U8), ("state_final", _P_U8),
    ]


class TrainOpts(C.Structure):
    """struct jh_train_opts."""

    _fields_ = [("mode", C.c_int32), ("term_kind", C.c_int32), ("max_iter", C.c_int32), ("eps", C.c_double)]


def make_desc(ir):
    """ModelIR -> (ModelDesc, keepalive list)."""
    d = ModelDesc()
    keep = []
    d.abi_version = ABI_VERSION
    d.n_states, d.alphabet, d.max_order = ir.n_states, ir.alphabet, ir.max_order
    d.n_elem, d.n_emis = ir.n_elem, ir.n_emis

    def put(name, dtype, ptype):
        arr = np.ascontiguousarray(getattr(ir, name), dtype=dtype)
        if arr.size == 0:
            arr = np.zeros(1, dtype=dtype)
        keep.append(arr)
        setattr(d, name, arr.ctypes.data_as(ptype))

    for f in ir.FIELDS_I32:
        put(f, np.int32, _P_I32)
    for f in ir.FIELDS_F64:
        put(f, np.float64, _P_F64)
    for f in ir.FIELDS_U8:
        put(f, np.uint8, _P_U8)
    return d, keep


class NativeError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"jstacs_hmm_b200 error {status}: {msg}")
        self.status = status


_lib = None


def lib():
    """Load the shared object once; raise if it was not built (python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                           "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp = C.c_void_p
    sig = {
        "jh_abi_version": (C.c_int, []),
        "jh_last_error": (C.c_char_p, []),
        "jh_device_count": (C.c_int, []),
        "jh_set_device": (C.c_int, [C.c_int]),
        "jh_kernel_launches": (C.c_int64, []),
        "jh_set_stream": (C.c_int, [vp]),
        "jh_synchronize": (C.c_int, []),
        "jh_model_create": (C.c_int, [C.POINTER(ModelDesc), C.POINTER(vp)]),
        "jh_model_destroy": (None, [vp]),
        "jh_model_set_params": (C.c_int, [vp, _P_F64, _P_F64, _P_F64, _P_F64]),
        "jh_model_get_params": (C.c_int, [vp, _P_F64, _P_F64, _P_F64, _P_F64]),
        "jh_model_sizes": (C.c_int, [vp, _P_I64]),
        "jh_model_log_prior": (C.c_int, [vp, _P_F64]),
        "jh_dataset_create": (C.c_int, [_P_U8, _P_I64, C.c_int64, _P_F64, C.c_int32, C.POINTER(vp)]),
        "jh_dataset_destroy": (None, [vp]),
        "jh_dataset_info": (C.c_int, [vp, _P_I64, _P_I64, _P_I64]),
        "jh_dataset_set_weights": (C.c_int, [vp, _P_F64]),
        "jh_loglik": (C.c_int, [vp, vp, _P_F64]),
        "jh_viterbi": (C.c_int, [vp, vp, _P_I32, _P_F64]),
        "jh_viterbi_u8": (C.c_int, [vp, vp, _P_U8, _P_F64]),
        "jh_posterior_decode_u8": (C.c_int, [vp, vp, _P_U8, _P_F64]),
        "jh_viterbi_paths": (C.c_int, [vp, vp, _P_I32, _P_I64, _P_I64, _P_F64]),
        "jh_loglik_windows": (C.c_int, [vp, vp, C.c_int64, _P_I64, _P_I64, _P_I64, _P_F64]),
        "jh_posterior": (C.c_int, [vp, vp, C.c_int64, _P_F64]),
        "jh_posterior_window": (C.c_int, [vp, vp, C.c_int64, C.c_int64, C.c_int64, _P_F64]),
        "jh_viterbi_windows": (C.c_int, [vp, vp, C.c_int64, _P_I64, _P_I64, _P_I64, _P_I32, _P_I64, _P_I64, _P_F64]),
        "jh_posterior_decode": (C.c_int, [vp, vp, _P_I32, _P_F64]),
        "jh_estep": (C.c_int, [vp, vp, C.c_int, _P_F64, _P_F64, _P_F64]),
        "jh_mstep": (C.c_int, [vp]),
        "jh_baum_welch": (C.c_int, [vp, vp, C.POINTER(TrainOpts), _P_F64, C.c_int32, _P_I32]),
        "jh_comm_unique_id": (C.c_int, [_P_U8]),
        "jh_comm_init": (C.c_int, [vp, _P_U8, C.c_int32, C.c_int32]),
        "jh_comm_destroy": (C.c_int, [vp]),
        "jh_comm_broadcast_params": (C.c_int, [vp]),
        "jh_stats_device": (C.c_int, [vp, C.POINTER(vp), _P_I64]),
        "jh_measure_fp64_peak": (C.c_int, [_P_F64]),
        "jh_last_kernel_ms": (C.c_int, [vp, _P_F64]),
        "jh_set_option": (C.c_int, [C.c_char_p, C.c_int64]),
        "jh_probe_fp64": (C.c_int, [C.c_int, C.c_int, C.c_int, _P_F64]),
        "jh_probe_dmma": (C.c_int, [C.c_int, C.c_int, C.c_int, _P_F64]),
        "jh_last_main_kernel_ms": (C.c_int, [vp, _P_F64]),
        "jh_model_set_option": (C.c_int, [vp, C.c_char_p, C.c_int64]),
        "jh_model_set_stream": (C.c_int, [vp, vp]),
        "jh_dataset_create_packed2": (C.c_int, [_P_U8, _P_I64, _P_I64, C.c_int64, _P_F64, C.c_int32, C.POINTER(vp)]),
        "jh_last_objective": (C.c_int, [vp, _P_F64]),
        "jh_model_set_states_groups": (C.c_int, [vp, C.c_int32, _P_I32, _P_I32]),
        "jh_dataset_set_groups": (C.c_int, [vp, _P_I32]),
        "jh_dataset_builder_create": (C.c_int, [C.c_int64, C.c_int64, C.c_int32, C.POINTER(vp)]),
        "jh_dataset_builder_append": (C.c_int, [vp, _P_U8, _P_I64, C.c_int64, C.c_int32]),
        "jh_dataset_builder_finish": (C.c_int, [vp, _P_F64, C.POINTER(vp)]),
        "jh_dataset_builder_destroy": (None, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    if L.jh_abi_version() != ABI_VERSION:
        raise RuntimeError("libjstacs_hmm_b200.so ABI mismatch")
    _lib = L
    return L


EXPORTS = None  # filled by tests from include/jstacs_hmm.h


def check(status: int):
    if status != 0:
        raise NativeError(status, lib().jh_last_error().decode("utf-8", "replace"))


def _f64(a):
    return None if a is None else a.ctypes.data_as(_P_F64)


IMPOSSIBLE = -6  # JH_ERR_IMPOSSIBLE


class NativeData:
    """jh_dataset handle: packed uint8 codes + int64 offsets (+ weights) resident on the device.
    byte_ptr given: `codes` holds 2-bit packed residues (four per byte, every sequence on a byte boundary at byte_ptr[n])."""

    def __init__(self, codes: np.ndarray, offsets: np.ndarray, weights, alphabet: int, byte_ptr=None):
        self.h = C.c_void_p()
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.n_seq = int(offsets.size - 1)
        self.n_res = int(offsets[-1])
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        cp = codes.ctypes.data_as(_P_U8) if codes.size else C.cast(None, _P_U8)
        if byte_ptr is None:
            check(lib().jh_dataset_create(cp, offsets.ctypes.data_as(_P_I64), self.n_seq, _f64(w), int(alphabet), C.byref(self.h)))
        else:
            bp = np.ascontiguousarray(byte_ptr, dtype=np.int64)
            check(lib().jh_dataset_create_packed2(cp, bp.ctypes.data_as(_P_I64), offsets.ctypes.data_as(_P_I64), self.n_seq, _f64(w),
                                                  int(alphabet), C.byref(self.h)))

    @classmethod
    def from_shards(cls, shards, n_seq: int, n_res: int, weights, alphabet: int):
        """Streamed construction (jh_dataset_builder_*): `shards` yields (codes, lengths, packed2) per shard — typically
        numpy memmaps of shard files; nothing but one shard is ever resident on the host."""
        self = cls.__new__(cls)
        self.h = C.c_void_p()
        self.n_seq, self.n_res = int(n_seq), int(n_res)
        b = C.c_void_p()
        check(lib().jh_dataset_builder_create(self.n_seq, self.n_res, int(alphabet), C.byref(b)))
        try:
            for codes, lengths, packed2 in shards:
                lengths = np.ascontiguousarray(lengths, dtype=np.int64)
                codes = np.ascontiguousarray(codes, dtype=np.uint8)  # a memmap stays a memmap: no copy
                cp = codes.ctypes.data_as(_P_U8) if codes.size else C.cast(None, _P_U8)
                check(lib().jh_dataset_builder_append(b, cp, lengths.ctypes.data_as(_P_I64), lengths.size, 1 if packed2 else 0))
            w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
            check(lib().jh_dataset_builder_finish(b, _f64(w), C.byref(self.h)))
            b = None
        finally:
            if b is not None:
                lib().jh_dataset_builder_destroy(b)
        return self

    def set_groups(self, groups):
        """allowedStatesGroup of every position (int32, -1 = unconstrained), None detaches (jh_dataset_set_groups)."""
        if groups is None:
            check(lib().jh_dataset_set_groups(self.h, None))
            return
        g = np.ascontiguousarray(groups, dtype=np.int32)
        if g.size != self.n_res:
            raise ValueError("one group index per residue is needed")
        check(lib().jh_dataset_set_groups(self.h, g.ctypes.data_as(_P_I32)))

    def set_weights(self, weights):
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        if w is not None and w.size != self.n_seq:
            raise ValueError("one weight per sequence is needed")
        check(lib().jh_dataset_set_weights(self.h, _f64(w)))

    def close(self):
        if self.h:
            lib().jh_dataset_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NativeModel:
    """jh_model handle."""

    def __init__(self, ir):
        self.h = C.c_void_p()
        desc, keep = make_desc(ir)
        check(lib().jh_model_create(C.byref(desc), C.byref(self.h)))
        del keep
        sz = (C.c_int64 * 4)()
        check(lib().jh_model_sizes(self.h, sz))
        self.sizes = tuple(int(v) for v in sz)  # n_child_total, n_elem, n_cond_total*a, n_cond_total

    def close(self):
        if self.h:
            lib().jh_model_destroy(self.h)
            self.h = C.c_void_p()

    def set_states_groups(self, groups):
        """statesGroups of the model constructor (AbstractHMM.java:173-228): a list of state-index lists."""
        ptr = np.zeros(len(groups) + 1, np.int32)
        ptr[1:] = np.cumsum([len(g) for g in groups])
        st = np.array([s for g in groups for s in g] or [0], np.int32)
        check(lib().jh_model_set_states_groups(self.h, len(groups), ptr.ctypes.data_as(_P_I32), st.ctypes.data_as(_P_I32)))

    def set_option(self, name: str, value: int):
        """Per-handle override of a library option (jh_model_set_option)."""
        check(lib().jh_model_set_option(self.h, name.encode(), int(value)))

    def set_stream(self, cuda_stream):
        check(lib().jh_model_set_stream(self.h, cuda_stream))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, tp, tl, ep, el):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (tp, tl, ep, el)]
        check(lib().jh_model_set_params(self.h, *[_f64(a) for a in arrs]))

    def get_params(self):
        out = [np.zeros(max(n, 1)) for n in self.sizes]
        check(lib().jh_model_get_params(self.h, *[_f64(a) for a in out]))
        return tuple(a[:n] for a, n in zip(out, self.sizes))

    def log_prior(self) -> float:
        v = C.c_double()
        check(lib().jh_model_log_prior(self.h, C.byref(v)))
        return v.value

    def loglik(self, data: NativeData) -> np.ndarray:
        out = np.zeros(max(data.n_seq, 1))
        check(lib().jh_loglik(self.h, data.h, _f64(out)))
        return out[:data.n_seq]

    def viterbi(self, data: NativeData, narrow: bool = False, allow_impossible: bool = False, out=None):
        """narrow: one byte per residue (uint8 paths, 255 = no state) instead of int32.
        allow_impossible: sequences without a state path of probability > 0 come back with score -inf and states
        -1 / 255 instead of raising JH_ERR_IMPOSSIBLE (the outputs are filled in both cases).
        out: reuse a path buffer of a previous call."""
        scores = np.zeros(max(data.n_seq, 1))
        if narrow:
            paths = out if out is not None else np.empty(max(data.n_res, 1), dtype=np.uint8)
            st = lib().jh_viterbi_u8(self.h, data.h, paths.ctypes.data_as(_P_U8), _f64(scores))
        else:
            paths = out if out is not None else np.empty(max(data.n_res, 1), dtype=np.int32)
            st = lib().jh_viterbi(self.h, data.h, paths.ctypes.data_as(_P_I32), _f64(scores))
        if not (st == IMPOSSIBLE and allow_impossible):
            check(st)
        return paths[:data.n_res], scores[:data.n_seq]

    def viterbi_paths(self, data: NativeData, capacity: np.ndarray):
        """General form (silent states): CSR capacity per sequence -> (list of paths, scores, path_len)."""
        ptr = np.zeros(data.n_seq + 1, dtype=np.int64)
        np.cumsum(np.asarray(capacity, dtype=np.int64), out=ptr[1:])
        paths = np.full(max(int(ptr[-1]), 1), -1, dtype=np.int32)
        plen = np.zeros(max(data.n_seq, 1), dtype=np.int64)
        scores = np.zeros(max(data.n_seq, 1))
        check(lib().jh_viterbi_paths(self.h, data.h, paths.ctypes.data_as(_P_I32), ptr.ctypes.data_as(_P_I64), plen.ctypes.data_as(_P_I64),
                                     _f64(scores)))
        return [paths[ptr[i]:ptr[i] + max(plen[i], 0)] for i in range(data.n_seq)], scores[:data.n_seq], plen[:data.n_seq]

    def loglik_windows(self, data: NativeData, seq_index, start, end) -> np.ndarray:
        si, st, en = (np.ascontiguousarray(a, dtype=np.int64) for a in (seq_index, start, end))
        if not (si.size == st.size == en.size):
            raise ValueError("seq_index, start and end need the same length")
        out = np.zeros(max(si.size, 1))
        check(lib().jh_loglik_windows(self.h, data.h, si.size, si.ctypes.data_as(_P_I64), st.ctypes.data_as(_P_I64), en.ctypes.data_as(_P_I64),
                                      _f64(out)))
        return out[:si.size]

    def viterbi_windows(self, data: NativeData, seq_index, start, end, n_silent: int = 0):
        """getViterbiPathFor(startPos, endPos, seq) for many windows -> (list of paths, scores)."""
        si, st, en = (np.ascontiguousarray(a, dtype=np.int64) for a in (seq_index, start, end))
        if not (si.size == st.size == en.size):
            raise ValueError("seq_index, start and end need the same length")
        n = si.size
        ptr = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.maximum(en - st + 1, 0) * (1 + n_silent) + n_silent, out=ptr[1:])
        paths = np.full(max(int(ptr[-1]), 1), -1, dtype=np.int32)
        plen = np.zeros(max(n, 1), dtype=np.int64)
        scores = np.zeros(max(n, 1))
        check(lib().jh_viterbi_windows(self.h, data.h, n, si.ctypes.data_as(_P_I64), st.ctypes.data_as(_P_I64), en.ctypes.data_as(_P_I64),
                                       paths.ctypes.data_as(_P_I32), ptr.ctypes.data_as(_P_I64), plen.ctypes.data_as(_P_I64), _f64(scores)))
        return [paths[ptr[i]:ptr[i] + max(plen[i], 0)] for i in range(n)], scores[:n], plen[:n]

    def posterior_window(self, data: NativeData, index: int, start: int, end: int, n_states: int) -> np.ndarray:
        length = end - start + 1
        out = np.zeros(max(n_states * length, 1))
        check(lib().jh_posterior_window(self.h, data.h, int(index), int(start), int(end), _f64(out)))
        return out[:n_states * length].reshape(n_states, length)

    def posterior(self, data: NativeData, index: int, n_states: int, length: int) -> np.ndarray:
        out = np.zeros(max(n_states * length, 1))
        check(lib().jh_posterior(self.h, data.h, int(index), _f64(out)))
        return out[:n_states * length].reshape(n_states, length)

    def posterior_decode(self, data: NativeData, narrow: bool = False):
        ll = np.zeros(max(data.n_seq, 1))
        if narrow:
            paths = np.empty(max(data.n_res, 1), dtype=np.uint8)
            check(lib().jh_posterior_decode_u8(self.h, data.h, paths.ctypes.data_as(_P_U8), _f64(ll)))
        else:
            paths = np.zeros(max(data.n_res, 1), dtype=np.int32)
            check(lib().jh_posterior_decode(self.h, data.h, paths.ctypes.data_as(_P_I32), _f64(ll)))
        return paths[:data.n_res], ll[:data.n_seq]

    def estep(self, data: NativeData, mode: int = MODE_BAUM_WELCH, fetch: bool = True):
        score = C.c_double()
        if fetch:
            ts, es = np.zeros(max(self.sizes[0], 1)), np.zeros(max(self.sizes[2], 1))
            check(lib().jh_estep(self.h, data.h, mode, _f64(ts), _f64(es), C.byref(score)))
            return ts[:self.sizes[0]], es[:self.sizes[2]], score.value
        check(lib().jh_estep(self.h, data.h, mode, None, None, None))
        return None

    def mstep(self):
        check(lib().jh_mstep(self.h))

    def baum_welch(self, data: NativeData, mode: int, term_kind: int, max_iter: int, eps: float, cap: int = 0):
        cap = cap or (max_iter if term_kind == TERM_ITERATIONS else 100000)
        cap = max(cap, 1)
        obj = np.zeros(cap)
        n_it = C.c_int32()
        opts = TrainOpts(mode, term_kind, max_iter, eps)
        check(lib().jh_baum_welch(self.h, data.h, C.byref(opts), _f64(obj), cap, C.byref(n_it)))
        self.n_iterations = n_it.value
        return obj[:min(n_it.value, cap)]

    def last_objective(self) -> float:
        """Objective of the last iteration of the most recent baum_welch (the history may be capped)."""
        v = C.c_double()
        check(lib().jh_last_objective(self.h, C.byref(v)))
        return v.value

    def stats_device(self):
        p, n = C.c_void_p(), C.c_int64()
        check(lib().jh_stats_device(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def last_main_kernel_ms(self) -> float:
        v = C.c_double()
        check(lib().jh_last_main_kernel_ms(self.h, C.byref(v)))
        return v.value

    def last_kernel_ms(self) -> float:
        v = C.c_double()
        check(lib().jh_last_kernel_ms(self.h, C.byref(v)))
        return v.value

