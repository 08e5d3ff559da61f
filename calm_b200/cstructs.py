"""ctypes mirror of include/calm_model.h (reference src/model.h:12-89).

Used by the tests and bench.py to drive both the product library
(libcalm_b200.so, device pointers) and the CPU checkers under oracle/
(host pointers) through the same records the reference driver fills in
run.c:32-117.
"""
import ctypes as C

MAX_LAYERS = 128
MAX_EXPERTS = 64
KV_SINKS = 2
FF_UPDATE_KV_ONLY = 1

_fptr = C.POINTER(C.c_float)


class Config(C.Structure):
    _fields_ = [
        ("dim", C.c_int),
        ("hidden_dim", C.c_int),
        ("head_dim", C.c_int),
        ("n_layers", C.c_int),
        ("n_heads", C.c_int),
        ("n_kv_heads", C.c_int),
        ("vocab_size", C.c_int),
        ("seq_len", C.c_int),
        ("rope_theta", C.c_float),
        ("rotary_dim", C.c_int),
        ("n_experts", C.c_int),
        ("n_experts_ac", C.c_int),
        ("norm_eps", C.c_float),
        ("act_gelu", C.c_bool),
        ("norm_ln", C.c_bool),
        ("norm_par", C.c_bool),
        ("qkv_clip", C.c_float),
    ]


class Weights(C.Structure):
    _fields_ = [
        ("dbits", C.c_int),
        ("token_embedding_table", C.c_void_p),
        ("rms_att_weight", C.c_void_p * MAX_LAYERS),
        ("rms_ffn_weight", C.c_void_p * MAX_LAYERS),
        ("wq", C.c_void_p * MAX_LAYERS),
        ("wk", C.c_void_p * MAX_LAYERS),
        ("wv", C.c_void_p * MAX_LAYERS),
        ("wo", C.c_void_p * MAX_LAYERS),
        ("w1", C.c_void_p * MAX_LAYERS),
        ("w2", C.c_void_p * MAX_LAYERS),
        ("w3", C.c_void_p * MAX_LAYERS),
        ("rms_final_weight", C.c_void_p),
        ("wcls", C.c_void_p),
        ("bqkv", C.c_void_p * MAX_LAYERS),
        ("moegate", C.c_void_p * MAX_LAYERS),
    ]


class RunState(C.Structure):
    _fields_ = [
        ("x", _fptr),
        ("xb", _fptr),
        ("xb2", _fptr),
        ("hb", _fptr),
        ("hb2", _fptr),
        ("he", _fptr),
        ("q", _fptr),
        ("k", _fptr),
        ("v", _fptr),
        ("att", _fptr),
        ("exp", _fptr),
        ("logits", _fptr),
        ("kvbits", C.c_int),
        ("key_cache", C.c_void_p),
        ("value_cache", C.c_void_p),
    ]


class Transformer(C.Structure):
    pass


FORWARD_FN = C.CFUNCTYPE(_fptr, C.POINTER(Transformer), C.c_int, C.c_int, C.c_uint)

Transformer._fields_ = [
    ("config", Config),
    ("weights", Weights),
    ("state", RunState),
    ("n_params", C.c_size_t),
    ("n_bytes", C.c_size_t),
    ("n_bandwidth", C.c_size_t),
    ("forward", C.c_void_p),
]

# sizes measured from the reference header (see include/calm_model.h)
assert C.sizeof(Config) == 60
assert C.sizeof(Weights) == 11296
assert C.sizeof(RunState) == 120
assert C.sizeof(Transformer) == 11512
