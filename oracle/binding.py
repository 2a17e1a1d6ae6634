"""ctypes binding of the CPU oracle (oracle/hmm_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libhmm_oracle.so")

_P_I32 = C.POINTER(C.c_int32)
_P_F64 = C.POINTER(C.c_double)
_P_U8 = C.POINTER(C.c_uint8)
_P_I64 = C.POINTER(C.c_int64)

_lib = None


def build(force: bool = False):
    import hashlib

    src = os.path.join(HERE, "hmm_oracle.c")
    with open(src, "rb") as f:
        digest = hashlib.sha256(f.read()).hexdigest()
    try:
        with open(LIB + ".srchash") as f:
            fresh = os.path.exists(LIB) and f.read().strip() == digest
    except OSError:
        fresh = False
    if force or not fresh:  # decided by content: file times need not survive the copy to the GPU box
        subprocess.check_call(["make", "-C", HERE, "-B"], stdout=subprocess.DEVNULL)
        with open(LIB + ".srchash", "w") as f:
            f.write(digest)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        from jstacs_b200._native import ModelDesc

        L = C.CDLL(LIB)
        vp = C.c_void_p
        L.jo_model_create.restype, L.jo_model_create.argtypes = vp, [C.POINTER(ModelDesc)]
        L.jo_model_destroy.restype, L.jo_model_destroy.argtypes = None, [vp]
        L.jo_model_get_params.restype, L.jo_model_get_params.argtypes = None, [vp, _P_F64, _P_F64, _P_F64, _P_F64]
        L.jo_model_set_params.restype, L.jo_model_set_params.argtypes = None, [vp, _P_F64, _P_F64, _P_F64, _P_F64]
        L.jo_log_prior.restype, L.jo_log_prior.argtypes = C.c_double, [vp]
        L.jo_mstep.restype, L.jo_mstep.argtypes = None, [vp, _P_F64, _P_F64]
        L.jo_loglik.restype, L.jo_loglik.argtypes = C.c_int, [vp, _P_U8, _P_I64, C.c_int64, _P_F64]
        L.jo_loglik_window.restype, L.jo_loglik_window.argtypes = C.c_double, [vp, _P_U8, C.c_int64, C.c_int64]
        L.jo_loglik_forward.restype, L.jo_loglik_forward.argtypes = C.c_double, [vp, _P_U8, C.c_int64]
        L.jo_matrices.restype, L.jo_matrices.argtypes = C.c_int, [vp, _P_U8, C.c_int64, C.c_int, _P_F64, _P_F64]
        L.jo_viterbi.restype = C.c_int
        L.jo_viterbi.argtypes = [vp, _P_U8, _P_I64, C.c_int64, _P_I32, _P_I64, _P_I64, _P_F64]
        L.jo_posterior.restype, L.jo_posterior.argtypes = C.c_int, [vp, _P_U8, C.c_int64, _P_F64]
        L.jo_posterior_window.restype, L.jo_posterior_window.argtypes = C.c_int, [vp, _P_U8, C.c_int64, C.c_int64, _P_F64]
        L.jo_viterbi_window.restype, L.jo_viterbi_window.argtypes = C.c_double, [vp, _P_U8, C.c_int64, C.c_int64, _P_I32, C.c_int64, _P_I64]
        L.jo_decode_posterior.restype, L.jo_decode_posterior.argtypes = None, [_P_F64, C.c_int, C.c_int64, _P_I32]
        L.jo_estep.restype = C.c_int
        L.jo_estep.argtypes = [vp, _P_U8, _P_I64, C.c_int64, _P_F64, C.c_int, C.c_int, _P_F64, _P_F64, _P_F64]
        L.jo_train.restype = C.c_int
        L.jo_train.argtypes = [vp, _P_U8, _P_I64, C.c_int64, _P_F64, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                               _P_F64, C.c_int]
        L.jo_model_set_groups.restype, L.jo_model_set_groups.argtypes = None, [vp, C.c_int, _P_I32, _P_I32]
        L.jo_model_attach_groups.restype, L.jo_model_attach_groups.argtypes = None, [vp, _P_U8, _P_I32]
        L.jo_path_logprob.restype = C.c_double
        L.jo_path_logprob.argtypes = [vp, _P_U8, C.c_int64, _P_I32, C.c_int64]
        _lib = L
    return _lib


def _u8(a):
    return a.ctypes.data_as(_P_U8)


def _i64(a):
    return a.ctypes.data_as(_P_I64)


def _f64(a):
    return None if a is None else a.ctypes.data_as(_P_F64)


class OracleModel:
    """CPU restatement of HigherOrderHMM over a flattened model (jstacs_b200.hmm.ModelIR)."""

    def __init__(self, ir):
        from jstacs_b200._native import make_desc

        self.ir = ir
        desc, keep = make_desc(ir)
        self.h = C.c_void_p(lib().jo_model_create(C.byref(desc)))
        del keep
        self.S = ir.n_states
        self.max_ctx = int(np.diff(ir.lc_ctx_ptr).max())
        self.sizes = (ir.n_child_total, ir.n_elem, ir.n_cond_total * ir.alphabet, ir.n_cond_total)
        self.n_silent = int(np.asarray(ir.state_silent).sum())

    def __del__(self):
        try:
            if self.h:
                lib().jo_model_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_groups(self, groups):
        """statesGroups (AbstractHMM.java:173-228): a list of state lists."""
        ptr = np.zeros(len(groups) + 1, np.int32)
        ptr[1:] = np.cumsum([len(g) for g in groups])
        st = np.array([s for g in groups for s in g] or [0], np.int32)
        lib().jo_model_set_groups(self.h, len(groups), ptr.ctypes.data_as(_P_I32), st.ctypes.data_as(_P_I32))

    def attach_groups(self, codes, groups):
        """allowedStatesGroup of every position (-1 = unconstrained), aligned with `codes` (the SAME contiguous uint8 array
        has to be passed to the following calls); None detaches."""
        if groups is None:
            self._g = None
            lib().jo_model_attach_groups(self.h, None, None)
            return
        assert codes.dtype == np.uint8 and codes.flags.c_contiguous
        self._g = (codes, np.ascontiguousarray(groups, np.int32))
        lib().jo_model_attach_groups(self.h, _u8(codes), self._g[1].ctypes.data_as(_P_I32))

    @staticmethod
    def _pack(codes, offsets):
        return np.ascontiguousarray(codes, np.uint8), np.ascontiguousarray(offsets, np.int64)

    def get_params(self):
        out = [np.zeros(max(n, 1)) for n in self.sizes]
        lib().jo_model_get_params(self.h, *[_f64(a) for a in out])
        return tuple(a[:n] for a, n in zip(out, self.sizes))

    def set_params(self, tp, tl, ep, el):
        arrs = [None if a is None else np.ascontiguousarray(a, np.float64) for a in (tp, tl, ep, el)]
        lib().jo_model_set_params(self.h, *[_f64(a) for a in arrs])

    def log_prior(self) -> float:
        return lib().jo_log_prior(self.h)

    def loglik(self, codes, offsets) -> np.ndarray:
        codes, offsets = self._pack(codes, offsets)
        n = offsets.size - 1
        out = np.zeros(max(n, 1))
        if lib().jo_loglik(self.h, _u8(codes), _i64(offsets), n, _f64(out)) != 0:
            raise ValueError("empty sequence")
        return out[:n]

    def loglik_window(self, x, start, end) -> float:
        x = np.ascontiguousarray(x, np.uint8)
        return lib().jo_loglik_window(self.h, _u8(x), start, end)

    def loglik_forward(self, x) -> float:
        x = np.ascontiguousarray(x, np.uint8)
        return lib().jo_loglik_forward(self.h, _u8(x), x.size)

    def matrices(self, x, viterbi: bool = False):
        x = np.ascontiguousarray(x, np.uint8)
        L = x.size
        fwd = np.zeros((L + 1, self.max_ctx))
        bwd = np.zeros((L + 1, self.max_ctx))
        lib().jo_matrices(self.h, _u8(x), L, 1 if viterbi else 0, _f64(fwd), _f64(bwd))
        return fwd, bwd

    def viterbi(self, codes, offsets):
        codes, offsets = self._pack(codes, offsets)
        n = offsets.size - 1
        lens = np.diff(offsets)
        cap = lens * (1 + self.n_silent) + self.n_silent
        pptr = np.zeros(n + 1, np.int64)
        np.cumsum(cap, out=pptr[1:])
        paths = np.zeros(max(int(pptr[-1]), 1), np.int32)
        plen = np.zeros(max(n, 1), np.int64)
        scores = np.zeros(max(n, 1))
        if lib().jo_viterbi(self.h, _u8(codes), _i64(offsets), n, paths.ctypes.data_as(_P_I32), _i64(pptr), _i64(plen),
                            _f64(scores)) != 0:
            raise ValueError("empty sequence")
        return [paths[pptr[i]:pptr[i] + max(plen[i], 0)] for i in range(n)], scores[:n], plen[:n]

    def posterior(self, x) -> np.ndarray:
        x = np.ascontiguousarray(x, np.uint8)
        out = np.zeros((self.S, x.size))
        if lib().jo_posterior(self.h, _u8(x), x.size, _f64(out)) != 0:
            raise ValueError("order-0 posterior not restated")
        return out

    def posterior_window(self, x, start: int, end: int) -> np.ndarray:
        x = np.ascontiguousarray(x, np.uint8)
        out = np.zeros((self.S, end - start + 1))
        if lib().jo_posterior_window(self.h, _u8(x), start, end, _f64(out)) != 0:
            raise ValueError("order-0 posterior not restated")
        return out

    def viterbi_window(self, x, start: int, end: int):
        x = np.ascontiguousarray(x, np.uint8)
        cap = (end - start + 1) * (1 + self.n_silent) + self.n_silent
        path = np.zeros(max(cap, 1), np.int32)
        plen = C.c_int64()
        sc = lib().jo_viterbi_window(self.h, _u8(x), start, end, path.ctypes.data_as(_P_I32), cap, C.byref(plen))
        return path[:max(plen.value, 0)], sc

    def decode_posterior(self, post) -> np.ndarray:
        post = np.ascontiguousarray(post, np.float64)
        path = np.zeros(post.shape[1], np.int32)
        lib().jo_decode_posterior(_f64(post), post.shape[0], post.shape[1], path.ctypes.data_as(_P_I32))
        return path

    def estep(self, codes, offsets, weights=None, mode: int = 1, threads: int = 1):
        codes, offsets = self._pack(codes, offsets)
        n = offsets.size - 1
        w = None if weights is None else np.ascontiguousarray(weights, np.float64)
        ts, es = np.zeros(max(self.sizes[0], 1)), np.zeros(max(self.sizes[2], 1))
        score = C.c_double()
        if lib().jo_estep(self.h, _u8(codes), _i64(offsets), n, _f64(w), mode, threads, _f64(ts), _f64(es),
                          C.byref(score)) != 0:
            raise ValueError("empty sequence")
        return ts[:self.sizes[0]], es[:self.sizes[2]], score.value

    def mstep(self, ts, es):
        ts = np.array(ts, np.float64)
        es = np.array(es, np.float64)
        lib().jo_mstep(self.h, _f64(ts), _f64(es))

    def train(self, codes, offsets, weights=None, mode: int = 1, term_kind: int = 0, max_iter: int = 10, eps: float = 0.0,
              threads: int = 1, cap: int = 0):
        codes, offsets = self._pack(codes, offsets)
        n = offsets.size - 1
        w = None if weights is None else np.ascontiguousarray(weights, np.float64)
        cap = cap or (max_iter if term_kind == 0 else 100000)
        obj = np.zeros(max(cap, 1))
        it = lib().jo_train(self.h, _u8(codes), _i64(offsets), n, _f64(w), mode, term_kind, max_iter, eps, threads,
                            _f64(obj), cap)
        if it < 0:
            raise ValueError("empty sequence")
        return obj[:min(it, cap)]

    def path_logprob(self, x, path) -> float:
        x = np.ascontiguousarray(x, np.uint8)
        p = np.ascontiguousarray(path, np.int32)
        return lib().jo_path_logprob(self.h, _u8(x), x.size, p.ctypes.data_as(_P_I32), p.size)
