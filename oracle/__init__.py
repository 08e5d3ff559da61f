"""Loader for the CPU checkers.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg and
--impl reference) may import this package; calm_b200/ never does.

Three interchangeable checkers, all driven through the reference's own
`struct Transformer` (calm_b200.cstructs) with HOST pointers:

  kind "port"       oracle/libcalm_oracle.so       our scalar-C restatement of reference src/infer.c
  kind "port_f64"   oracle/libcalm_oracle_f64.so   same, reductions in double (referee)
  kind "reference"  oracle/_ref/libcalm_ref_cpu.so the UNMODIFIED reference src/infer.c compiled with the
                                                    reference's flags by oracle/Makefile (symbols prepare/forward,
                                                    reference run.c:19-20)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from calm_b200.cstructs import FF_UPDATE_KV_ONLY, Transformer

HERE = os.path.dirname(os.path.abspath(__file__))
_fptr = C.POINTER(C.c_float)

_PATHS = {
    "port": os.path.join(HERE, "libcalm_oracle.so"),
    "port_f64": os.path.join(HERE, "libcalm_oracle_f64.so"),
    "reference": os.path.join(HERE, "_ref", "libcalm_ref_cpu.so"),
}


def build(ref: bool = True) -> None:
    """Compile the restatement and, when /root/reference is present, the reference itself."""
    targets = ["all"] + (["ref"] if ref else [])
    subprocess.run(["make", "-C", HERE, "--no-print-directory"] + targets, check=True, stdout=subprocess.DEVNULL)


def available(kind: str) -> bool:
    return os.path.exists(_PATHS[kind])


class Checker:
    """prepare()/forward() of one CPU implementation over a calm_b200.modelgen.HostModel."""

    def __init__(self, kind: str = "port"):
        self.kind = kind
        path = _PATHS[kind]
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle all ref`")
        self.lib = C.CDLL(path)
        pre = "" if kind == "reference" else "oracle_"
        self._prepare = getattr(self.lib, pre + "prepare")
        self._prepare.argtypes = [C.POINTER(Transformer)]
        self._prepare.restype = None
        self._forward = getattr(self.lib, pre + "forward")
        self._forward.argtypes = [C.POINTER(Transformer), C.c_int, C.c_int, C.c_uint]
        self._forward.restype = _fptr
        if kind != "reference":
            L = self.lib
            L.oracle_forward_taps.argtypes = [C.POINTER(Transformer), C.c_int, C.c_int, C.c_uint, _fptr, _fptr, _fptr]
            L.oracle_forward_taps.restype = _fptr
            L.oracle_matvec.argtypes = [C.c_int, C.c_void_p, _fptr, _fptr, _fptr, C.c_int, C.c_int]
            L.oracle_matvec.restype = None
            L.oracle_read_kv.argtypes = [C.POINTER(Transformer), C.c_int, C.c_int, _fptr, _fptr]
            L.oracle_read_kv.restype = None
            L.oracle_release.argtypes = [C.POINTER(Transformer)]
            L.oracle_release.restype = None
            L.oracle_argmax.argtypes = [_fptr, C.c_int]
            L.oracle_argmax.restype = C.c_int

    def prepare(self, model) -> None:
        self._prepare(C.byref(model.transformer))

    def forward(self, model, token: int, pos: int, flags: int = 0):
        """Returns a COPY of the logits (np.float32[vocab]) or None for FF_UPDATE_KV_ONLY."""
        p = self._forward(C.byref(model.transformer), token, pos, flags)
        if not p:
            return None
        return np.ctypeslib.as_array(p, shape=(model.spec.vocab_size,)).copy()

    def forward_taps(self, model, token: int, pos: int):
        s = model.spec
        tq = np.zeros((s.n_layers, s.q_dim), np.float32)
        ta = np.zeros((s.n_layers, s.q_dim), np.float32)
        tx = np.zeros((s.n_layers, s.dim), np.float32)
        p = self.lib.oracle_forward_taps(C.byref(model.transformer), token, pos, 0, tq.ctypes.data_as(_fptr),
                                         ta.ctypes.data_as(_fptr), tx.ctypes.data_as(_fptr))
        logits = np.ctypeslib.as_array(p, shape=(s.vocab_size,)).copy()
        return logits, tq, ta, tx

    def read_kv(self, model, layer: int, kv_pos: int):
        """KV entry as float32 (kv_dim,) x2.  For the unmodified reference the cache is read directly
        (fp16 [layer][pos][kv_dim], reference infer.c:355-381)."""
        s = model.spec
        if self.kind != "reference":
            k = np.zeros(s.kv_dim, np.float32)
            v = np.zeros(s.kv_dim, np.float32)
            self.lib.oracle_read_kv(C.byref(model.transformer), layer, kv_pos, k.ctypes.data_as(_fptr), v.ctypes.data_as(_fptr))
            return k, v
        st = model.transformer.state
        n = s.n_layers * model.seq_len * s.kv_dim
        off = (layer * model.seq_len + kv_pos) * s.kv_dim
        kc = np.ctypeslib.as_array(C.cast(st.key_cache, C.POINTER(C.c_uint16)), shape=(n,)).view(np.float16)
        vc = np.ctypeslib.as_array(C.cast(st.value_cache, C.POINTER(C.c_uint16)), shape=(n,)).view(np.float16)
        return kc[off:off + s.kv_dim].astype(np.float32), vc[off:off + s.kv_dim].astype(np.float32)

    def matvec(self, dbits: int, w: np.ndarray, x: np.ndarray, n: int, d: int, bias=None) -> np.ndarray:
        assert self.kind != "reference"
        y = np.zeros(d, np.float32)
        x = np.ascontiguousarray(x, np.float32)
        b = None if bias is None else np.ascontiguousarray(bias, np.float32).ctypes.data_as(_fptr)
        self.lib.oracle_matvec(dbits, w.ctypes.data, x.ctypes.data_as(_fptr), b, y.ctypes.data_as(_fptr), n, d)
        return y

    def release(self, model) -> None:
        if self.kind != "reference":
            self.lib.oracle_release(C.byref(model.transformer))


def teacher_forced(checker: Checker, model, tokens, pos0: int = 0):
    """Feed `tokens` at positions pos0.. and return the stacked logits (len(tokens), vocab)."""
    checker.prepare(model)
    out = [checker.forward(model, tok, pos0 + i, 0) for i, tok in enumerate(tokens)]
    return np.stack(out)


# ---------------------------------------------------------------------------------------------
# sampler (reference src/sampler.c): our restatement and the unmodified reference, same call shape

class _RefSampler(C.Structure):  # struct Sampler, reference sampler.h:3-9
    _fields_ = [("vocab_size", C.c_int), ("rng_state", C.c_ulonglong), ("temperature", C.c_float), ("minp", C.c_float)]


class Sampler:
    """sample(logits) -> token with an explicit xorshift* state.  kind "port" = oracle_sample (calm_oracle.c),
    kind "reference" = sample() of the unmodified sampler.c (oracle/_ref/libcalm_ref_sampler.so)."""

    def __init__(self, kind: str, temperature: float, minp: float, rng_state: int):
        self.kind, self.temperature, self.minp, self.rng_state = kind, float(temperature), float(minp), int(rng_state)
        path = _PATHS["port"] if kind == "port" else os.path.join(HERE, "_ref", "libcalm_ref_sampler.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        if kind == "port":
            self.lib.oracle_sample.argtypes = [_fptr, C.c_int, C.c_float, C.c_float, C.POINTER(C.c_ulonglong)]
            self.lib.oracle_sample.restype = C.c_int
        else:
            self.lib.sample.argtypes = [C.POINTER(_RefSampler), _fptr]
            self.lib.sample.restype = C.c_int

    def sample(self, logits: np.ndarray) -> int:
        buf = np.array(logits, np.float32, copy=True)  # the reference overwrites its argument with probabilities
        if self.kind == "port":
            rng = C.c_ulonglong(self.rng_state)
            tok = self.lib.oracle_sample(buf.ctypes.data_as(_fptr), len(buf), self.temperature, self.minp, C.byref(rng))
            self.rng_state = int(rng.value)
            return tok
        s = _RefSampler(len(buf), self.rng_state, self.temperature, self.minp)
        tok = self.lib.sample(C.byref(s), buf.ctypes.data_as(_fptr))
        self.rng_state = int(s.rng_state)
        return tok
