"""ctypes binding of libcalm_b200.so -- the host-side mirror of the reference's backend interface.

The function names, argument meaning and error behaviour are the reference's
(run.c:22-25): upload_cuda / prepare_cuda / forward_cuda / perf_cuda.  There is
no CPU path behind these calls: loading fails loudly when the library has not
been built, and prepare_cuda aborts the process when no sm_100 device exists.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np

from .cstructs import FF_UPDATE_KV_ONLY, Transformer
from . import modelgen as mg

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcalm_b200.so")
_fptr = C.POINTER(C.c_float)

# every symbol include/calm_b200.h declares
SYMBOLS = [
    "upload_cuda", "prepare_cuda", "forward_cuda", "perf_cuda", "forward_prefill_cuda",
    "calm_b200_abi_version", "calm_b200_set_device", "calm_b200_free", "calm_b200_release",
    "calm_b200_forward_argmax", "calm_b200_decode_greedy", "calm_b200_timer_start", "calm_b200_timer_stop",
    "calm_b200_stream", "calm_b200_launch_count", "calm_b200_read_kv", "calm_b200_fill_kv", "calm_b200_matvec",
    "calm_b200_set_perf", "calm_b200_stage_stats", "calm_b200_perf_token_ms", "calm_b200_debug_stamps",
    "calm_b200_tp_unique_id", "calm_b200_tp_init", "calm_b200_tp_world", "calm_b200_tp_mode",
    "calm_b200_decode_sample", "calm_b200_forward_sample", "calm_b200_read_device_logits", "calm_b200_sample_logits",
]

_lib = None


def load() -> C.CDLL:
    """Load the product library.  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not built: run `python -m calm_b200.build` (CUDA extension is mandatory, no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    T = C.POINTER(Transformer)
    L.upload_cuda.argtypes, L.upload_cuda.restype = [C.c_void_p, C.c_size_t], C.c_void_p
    L.prepare_cuda.argtypes, L.prepare_cuda.restype = [T], None
    L.forward_cuda.argtypes, L.forward_cuda.restype = [T, C.c_int, C.c_int, C.c_uint], _fptr
    L.perf_cuda.argtypes, L.perf_cuda.restype = [], None
    L.forward_prefill_cuda.argtypes, L.forward_prefill_cuda.restype = [T, C.POINTER(C.c_int), C.c_int, C.c_int], C.c_int
    L.calm_b200_abi_version.argtypes, L.calm_b200_abi_version.restype = [], C.c_int
    L.calm_b200_set_device.argtypes, L.calm_b200_set_device.restype = [C.c_int], None
    L.calm_b200_free.argtypes, L.calm_b200_free.restype = [C.c_void_p], None
    L.calm_b200_release.argtypes, L.calm_b200_release.restype = [T], None
    L.calm_b200_forward_argmax.argtypes, L.calm_b200_forward_argmax.restype = [T, C.c_int, C.c_int], C.c_int
    L.calm_b200_decode_greedy.argtypes, L.calm_b200_decode_greedy.restype = [T, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)], None
    L.calm_b200_timer_start.argtypes, L.calm_b200_timer_start.restype = [], None
    L.calm_b200_timer_stop.argtypes, L.calm_b200_timer_stop.restype = [], C.c_float
    L.calm_b200_stream.argtypes, L.calm_b200_stream.restype = [], C.c_void_p
    L.calm_b200_launch_count.argtypes, L.calm_b200_launch_count.restype = [], C.c_uint64
    L.calm_b200_read_kv.argtypes, L.calm_b200_read_kv.restype = [T, C.c_int, C.c_int, _fptr, _fptr], None
    L.calm_b200_fill_kv.argtypes, L.calm_b200_fill_kv.restype = [T, C.c_int, C.c_uint64], None
    L.calm_b200_matvec.argtypes = [C.c_int, C.c_void_p, _fptr, _fptr, C.c_int, C.c_int, C.c_int, C.c_int]
    L.calm_b200_matvec.restype = C.c_float
    L.calm_b200_set_perf.argtypes, L.calm_b200_set_perf.restype = [C.c_int], None
    L.calm_b200_stage_stats.argtypes = [C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_long)]
    L.calm_b200_stage_stats.restype = C.c_int
    L.calm_b200_perf_token_ms.argtypes, L.calm_b200_perf_token_ms.restype = [], C.c_double
    L.calm_b200_debug_stamps.argtypes, L.calm_b200_debug_stamps.restype = [C.POINTER(C.c_ulonglong)], None
    L.calm_b200_tp_unique_id.argtypes, L.calm_b200_tp_unique_id.restype = [C.c_void_p], None
    L.calm_b200_tp_init.argtypes, L.calm_b200_tp_init.restype = [C.c_int, C.c_int, C.c_void_p], None
    L.calm_b200_tp_world.argtypes, L.calm_b200_tp_world.restype = [], C.c_int
    L.calm_b200_tp_mode.argtypes, L.calm_b200_tp_mode.restype = [], C.c_int
    L.calm_b200_decode_sample.argtypes = [T, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.POINTER(C.c_ulonglong), C.POINTER(C.c_int)]
    L.calm_b200_decode_sample.restype = None
    L.calm_b200_forward_sample.argtypes = [T, C.c_int, C.c_int, C.c_float, C.c_float, C.POINTER(C.c_ulonglong)]
    L.calm_b200_forward_sample.restype = C.c_int
    L.calm_b200_read_device_logits.argtypes, L.calm_b200_read_device_logits.restype = [_fptr], None
    L.calm_b200_sample_logits.argtypes = [_fptr, C.c_int, C.c_float, C.c_float, C.POINTER(C.c_ulonglong), _fptr]
    L.calm_b200_sample_logits.restype = C.c_int
    _lib = L
    return L


def tp_unique_id() -> bytes:
    """128-byte NCCL id for calm_b200_tp_init; rank 0 makes it, every rank of the group receives it."""
    buf = C.create_string_buffer(128)
    load().calm_b200_tp_unique_id(buf)
    return buf.raw


class DeviceModel:
    """A model on the GPU behind the reference boundary.

    Mirrors what the reference driver does around the backend (run.c:552-582):
    upload every model.* tensor with upload_cuda, bind the returned device pointers
    in struct Weights, call prepare_cuda, then forward_cuda once per token.
    `tensors` may hold host tensors (uploaded through upload_cuda) or tensors that
    already live on the device (their data_ptr is used directly, which is what
    struct Weights carries after upload anyway).
    """

    def __init__(self, spec: mg.ModelSpec, tensors: Dict[str, "object"], seq_len: Optional[int] = None, kvbits: int = 16,
                 device: Optional[int] = None, tp=None):
        """tp = (rank, world, id128) turns on tensor parallelism (include/calm_b200.h); tensors stay the FULL model."""
        self.lib = load()
        self.spec = spec
        if device is not None:
            self.lib.calm_b200_set_device(device)
        if tp is not None:
            self.lib.calm_b200_tp_init(int(tp[0]), int(tp[1]), C.c_char_p(tp[2]) if tp[2] else None)
        self._uploaded = []
        self._keep = tensors
        ptrs = {}
        for name, t in tensors.items():
            if getattr(t, "is_cuda", False):
                ptrs[name] = t.data_ptr()
            else:
                t = t.contiguous()
                nbytes = t.numel() * t.element_size()
                p = self.lib.upload_cuda(t.data_ptr(), nbytes)
                self._uploaded.append(p)
                ptrs[name] = p
        self.transformer = mg.fill_transformer(spec, lambda n: ptrs.get(n, 0), seq_len, kvbits)
        self.lib.prepare_cuda(C.byref(self.transformer))
        self.vocab = spec.vocab_size

    @property
    def seq_len(self) -> int:
        return self.transformer.config.seq_len

    def forward(self, token: int, pos: int, flags: int = 0):
        """forward_cuda; returns a COPY of the host logits, or None for FF_UPDATE_KV_ONLY."""
        p = self.lib.forward_cuda(C.byref(self.transformer), token, pos, flags)
        if not p:
            return None
        return np.ctypeslib.as_array(p, shape=(self.vocab,)).copy()

    def prefill(self, tokens, pos0: int = 0) -> int:
        """forward_prefill_cuda: the batched prompt pass; returns 1 (tensor-core pass) or 0 (fed token by token)."""
        arr = np.ascontiguousarray(tokens, np.int32)
        return self.lib.forward_prefill_cuda(C.byref(self.transformer), arr.ctypes.data_as(C.POINTER(C.c_int)), len(arr), pos0)

    def forward_raw(self, token: int, pos: int, flags: int = 0):
        return self.lib.forward_cuda(C.byref(self.transformer), token, pos, flags)

    def forward_argmax(self, token: int, pos: int) -> int:
        return self.lib.calm_b200_forward_argmax(C.byref(self.transformer), token, pos)

    def decode_greedy(self, token0: int, pos0: int, n: int) -> np.ndarray:
        out = np.zeros(n, np.int32)
        self.lib.calm_b200_decode_greedy(C.byref(self.transformer), token0, pos0, n, out.ctypes.data_as(C.POINTER(C.c_int)))
        return out

    def decode_sample(self, token0: int, pos0: int, n: int, temperature: float, minp: float, rng_state: int):
        """Device-resident temperature / min-p sampling; returns (tokens, new rng state)."""
        out = np.zeros(n, np.int32)
        rng = C.c_ulonglong(rng_state)
        self.lib.calm_b200_decode_sample(C.byref(self.transformer), token0, pos0, n, temperature, minp, C.byref(rng), out.ctypes.data_as(C.POINTER(C.c_int)))
        return out, int(rng.value)

    def device_logits(self) -> np.ndarray:
        out = np.zeros(self.vocab, np.float32)
        self.lib.calm_b200_read_device_logits(out.ctypes.data_as(_fptr))
        return out

    def logits_view(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.transformer.state.logits, shape=(self.vocab,))

    def read_kv(self, layer: int, kv_pos: int):
        k = np.zeros(self.spec.kv_dim, np.float32)
        v = np.zeros(self.spec.kv_dim, np.float32)
        self.lib.calm_b200_read_kv(C.byref(self.transformer), layer, kv_pos, k.ctypes.data_as(_fptr), v.ctypes.data_as(_fptr))
        return k, v

    def fill_kv(self, n_pos: int, seed: int = 1) -> None:
        self.lib.calm_b200_fill_kv(C.byref(self.transformer), n_pos, seed)

    def profile(self, token: int, pos0: int, n: int):
        """Run n greedy tokens of the production graph with in-kernel stage stamps; returns
        ({stage: (ms_total, bytes_total, launches)}, mean token span in ms)."""
        self.lib.calm_b200_set_perf(1)
        tok = token
        for i in range(n):
            tok = self.forward_argmax(tok, pos0 + i)
        out = {}
        i = 0
        name = C.create_string_buffer(64)
        ms, by, nl = C.c_double(), C.c_double(), C.c_long()
        while self.lib.calm_b200_stage_stats(i, name, 64, C.byref(ms), C.byref(by), C.byref(nl)):
            out[name.value.decode()] = (ms.value, by.value, nl.value)
            i += 1
        span = self.lib.calm_b200_perf_token_ms()
        self.lib.calm_b200_set_perf(0)
        return out, span

    def close(self) -> None:
        if self.transformer is not None:
            self.lib.calm_b200_release(C.byref(self.transformer))
            for p in self._uploaded:
                self.lib.calm_b200_free(p)
            self._uploaded = []
            self.transformer = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def matvec(dbits: int, w_dev_ptr: int, x: np.ndarray, n: int, d: int, warmup: int = 0, iters: int = 1):
    """Run the production matvec kernel once (y = W.x); returns (y, ms per launch)."""
    L = load()
    x = np.ascontiguousarray(x, np.float32)
    y = np.zeros(d, np.float32)
    ms = L.calm_b200_matvec(dbits, w_dev_ptr, x.ctypes.data_as(_fptr), y.ctypes.data_as(_fptr), n, d, warmup, iters)
    return y, ms


def sample_logits(logits: np.ndarray, temperature: float, minp: float, rng_state: int, timed: bool = False):
    """The device sampler on host logits; returns (token, new rng state[, microseconds of the sampler kernels])."""
    L = load()
    buf = np.ascontiguousarray(logits, np.float32)
    rng = C.c_ulonglong(rng_state)
    us = C.c_float(0)
    tok = L.calm_b200_sample_logits(buf.ctypes.data_as(_fptr), len(buf), temperature, minp, C.byref(rng), C.byref(us) if timed else None)
    return (tok, int(rng.value), us.value) if timed else (tok, int(rng.value))
