"""ctypes binding of libnavc.so (C ABI declared in include/navc.h).

The product path has no CPU or PyTorch fallback: if the shared library is missing or the device is
not an sm_100 GPU, loading / initialisation raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libnavc.so")

c_f32p = C.c_void_p
i32, i64, u64, f32, vp = C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_void_p


class Epilogue(C.Structure):
    _fields_ = [("bias", vp), ("residual", vp), ("row_tokens", vp), ("act", i32), ("ld_res", i32),
                ("out_f32", vp), ("out_hi", vp), ("out_lo", vp), ("ld_out", i32), ("reserved", i32),
                ("split_k", i32), ("accumulate", i32), ("res_hi", vp), ("res_lo", vp), ("m_dev", vp),
                ("m_hint", i32), ("pad_", i32)]


class Step(C.Structure):
    _fields_ = [("part_max", vp), ("part_sum", vp), ("part_idx", vp), ("n_tiles", i32),
                ("is_ct", i32), ("merge", i32), ("select", i32), ("q", i32), ("ratio", f32),
                ("win_lo", i32), ("win_hi", i32),
                ("lens", vp), ("teacher", vp), ("given", vp), ("tokens", vp), ("probs", vp),
                ("upd_mask", vp), ("canvas", vp), ("lprobs", vp), ("counters", vp), ("visual", vp),
                ("masked0", vp), ("seq_off", vp), ("part_slot", vp), ("sel_rows", vp), ("sel_count", vp), ("sel_slot", vp)]


ACT = {"none": 0, None: 0, "gelu_new": 1, "gelu": 2, "relu": 3, "swish": 4}
MASK_KIND = {"NARFormer": 0, "ARFormer": 1, "SelfMask": 2}
TC_BF16, TC_BF16X3 = 1, 3
MERGE_NONE, MERGE_ALL, MERGE_MASKED, MERGE_EF = 0, 1, 2, 3
SELECT_NONE, SELECT_WORST, SELECT_MASKTOK, SELECT_GIVEN, SELECT_KEEP, SELECT_WINDOW = 0, 1, 2, 3, 4, 5

# name -> argtypes ; every function returns int except the three noted below
_PROTOS = {
    "navc_init": [i32],
    "navc_linear_f32": [vp, i32, vp, i32, i32, i32, i32, C.POINTER(Epilogue), vp],
    "navc_linear_tc": [i32, vp, vp, i32, vp, vp, i32, i32, i32, i32, C.POINTER(Epilogue), vp],
    "navc_wgrad_tc": [i32, vp, vp, i32, vp, vp, i32, i32, i32, i32, C.POINTER(Epilogue), vp],
    "navc_dgrad_tc": [i32, vp, vp, i32, vp, vp, i32, i32, i32, i32, C.POINTER(Epilogue), vp],
    "navc_split_bf16": [vp, vp, vp, i64, vp],
    "navc_join_bf16": [vp, vp, vp, i64, vp],
    "navc_vocab_tile": [i32],
    "navc_vocab_partials_f32": [vp, i32, vp, i32, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp],
    "navc_vocab_partials_tc": [i32, vp, vp, i32, vp, vp, i32, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp],
    "navc_log_softmax": [vp, vp, i32, i32, i32, vp],
    "navc_log_softmax_ld": [vp, i32, vp, i32, i32, i32, vp],
    "navc_highway_bn": [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, f32, vp, vp, vp, vp, vp],
    "navc_highway_ln": [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, f32, vp, vp, vp, vp, vp],
    "navc_length_head": [vp, i32, i32, i32, vp, vp, vp, vp, i32, vp, vp, vp],
    "navc_embed_ln": [vp, vp, vp, vp, vp, vp, i32, vp, vp, f32, i32, i32, i32, vp, vp, vp, vp],
    "navc_layernorm": [vp, vp, vp, f32, vp, i32, i32, vp, vp, vp, vp],
    "navc_self_attention": [vp, i32, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp],
    "navc_cross_attention": [vp, i32, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp],
    "navc_self_attention_tc": [i32, vp, vp, i32, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp],
    "navc_cross_attention_tc": [i32, vp, vp, i32, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp],
    "navc_pack_rows": [vp, i32, i32, vp, vp, vp],
    "navc_embed_ln_packed": [vp, vp, vp, vp, vp, vp, i32, vp, vp, f32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp],
    "navc_self_attention_tc_packed": [i32, vp, vp, i32, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp],
    "navc_cross_attention_tc_packed": [i32, vp, vp, i32, vp, vp, i32, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp],
    "navc_attention_window": [],
    "navc_pack_tiles": [vp, i32, vp, i32, vp],
    "navc_self_attention_tc_tiles": [i32, vp, vp, i32, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp],
    "navc_cross_attention_tc_tiles": [i32, vp, vp, i32, vp, vp, i32, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp],
    "navc_gather_rows": [vp, vp, i32, vp, vp, i32, vp, vp, vp],
    "navc_compact_rows": [vp, vp, i32, i32, vp, vp, vp, vp],
    "navc_gather_rows2": [vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, i32, vp],
    "navc_refresh_pack": [vp, i32, vp],
    "navc_linear_chain_tc": [i32, vp, vp, i32, vp, vp, i32, C.POINTER(Epilogue), vp, vp, i32, C.POINTER(Epilogue), i32, i32, i32, vp],
    "navc_drop_add_bwd_split": [vp, u64, f32, u64, f32, vp, i32, i32, vp, vp, vp, i32, vp, vp],
    "navc_act_drop_bwd_split": [vp, vp, i32, u64, f32, i32, i32, vp, vp, i32, vp, vp],
    "navc_cross_attention_bwd_tc_split": [i32, vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, vp, vp, vp, i32, vp, vp, i32, vp, vp],
    "navc_linear_tf32": [vp, i32, vp, i32, i32, i32, i32, C.POINTER(Epilogue), vp],
    "navc_vocab_partials_tc_dyn": [i32, vp, vp, i32, vp, vp, i32, vp, i32, i32, i32, vp, vp, vp, vp, vp],
    "navc_length_beam": [vp, i32, i32, i32, i32, vp, vp, vp],
    "navc_init_canvas": [vp, i32, i32, i64, vp, vp, vp, vp],
    "navc_refine_step": [C.POINTER(Step), i32, i32, vp],
    "navc_teacher_probs": [vp, vp, i32, vp, vp, i32, i32, vp, vp],
    "navc_teacher_inputs": [vp, vp, i32, i32, vp, vp, vp],
    "navc_select_best": [vp, vp, vp, i32, i32, i32, f32, vp, vp, vp],
    # training path
    "navc_drop_add": [vp, vp, u64, f32, u64, f32, vp, i32, i32, vp, vp, vp, vp],
    "navc_drop_add_bwd": [vp, u64, f32, u64, f32, vp, i32, i32, vp, vp, vp],
    "navc_act_drop": [vp, i32, u64, f32, i64, vp, vp, vp, vp],
    "navc_act_drop_bwd": [vp, vp, i32, u64, f32, i64, vp, vp],
    "navc_transpose_pack": [vp, i32, i32, i32, vp, vp, i32, vp, vp, vp, i32, vp, vp],
    "navc_highway_fwd_train": [vp, vp, i32, i32, i32, u64, f32, vp, vp],
    "navc_highway_bwd": [vp, vp, vp, i32, i32, i32, u64, f32, vp, vp, vp],
    "navc_bn_stats": [vp, i32, i32, vp, vp, f32, vp, vp, vp],
    "navc_bn_apply_concat": [vp, vp, vp, vp, vp, f32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp],
    "navc_bn_bwd": [vp, vp, vp, vp, vp, vp, f32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp],
    "navc_mean_bwd": [vp, i32, i32, i32, vp, vp],
    "navc_log_softmax_bwd": [vp, vp, i32, i32, i32, vp, i32, vp],
    "navc_layernorm_bwd": [vp, vp, vp, f32, vp, i32, i32, vp, vp, vp, vp],
    "navc_embed_ln_bwd": [vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, f32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp],
    "navc_self_attention_step": [vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp],
    "navc_streamk_error": [],
    "navc_set_streamk": [i32],
    "navc_beam_topk": [vp, i32, i32, i32, i32, vp, vp, i32, i32, i32, vp, vp, vp],
    "navc_beam_advance": [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp],
    "navc_self_attention_tc_rows": [i32, vp, vp, i32, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp],
    "navc_cross_attention_tc_rows": [i32, vp, vp, i32, vp, vp, i32, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp],
    "navc_self_attention_bwd_packed": [vp, i32, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp],
    "navc_cross_attention_bwd_packed": [vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, vp, vp, i32, vp, i32, vp],
    "navc_self_attention_bwd_tc": [i32, vp, i32, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp],
    "navc_cross_attention_bwd_tc": [i32, vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, vp, vp, vp, i32, vp, i32, vp],
    "navc_embed_ln_bwd_packed": [vp, vp, vp, vp, vp, vp, vp, i32, vp, f32, i32, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp],
    "navc_log_softmax_rows": [vp, i32, vp, i32, vp, i32, i32, vp],
    "navc_log_softmax_bwd_rows": [vp, vp, vp, i32, i32, i32, vp, i32, vp],
    "navc_log_softmax_bwd_padrows": [vp, i32, vp, i32, vp, i32, vp, vp],
    "navc_rows_f32": [vp, vp, i32, vp, i32, i32, vp],
    "navc_ce_stats": [vp, vp, vp, i32, vp, vp, i32, vp, vp, vp, vp],
    "navc_ce_grad": [vp, vp, vp, vp, i32, i32, i32, vp],
    "navc_clip_adam": [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, f32, i32, vp],
    "navc_self_attention_bwd": [vp, i32, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp],
    "navc_cross_attention_bwd": [vp, i32, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, i32, vp, i32, vp],
}
EXPORTS = ["navc_version", "navc_last_error", "navc_sm_count"] + list(_PROTOS)

_lib = None
_lock = threading.Lock()
_inited = set()
launches = 0  # number of kernel launches issued through this binding (bench.py reports it)


class NavcError(RuntimeError):
    pass


def load():
    """dlopen libnavc.so and declare prototypes.  No CUDA call is made here."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.isfile(LIB_PATH):
                raise NavcError("libnavc.so not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "or `make -C <pkg>/csrc`. There is no CPU fallback." % LIB_PATH)
            lib = C.CDLL(LIB_PATH)
            lib.navc_version.restype = i32
            lib.navc_last_error.restype = C.c_char_p
            lib.navc_sm_count.restype = i32
            for name, args in _PROTOS.items():
                fn = getattr(lib, name)
                fn.argtypes = args
                fn.restype = i32
            _lib = lib
    return _lib


def ensure_init(device: torch.device):
    """Initialise the library for `device`; raises unless it is an sm_100 CUDA device."""
    lib = load()
    if device.type != "cuda":
        raise NavcError("navc kernels need a CUDA (sm_100a) device, got %s; there is no CPU fallback" % device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _inited:
        with torch.cuda.device(idx):
            rc = lib.navc_init(idx)
        if rc:
            raise NavcError(lib.navc_last_error().decode())
        _inited.add(idx)
    return lib


def call(name, *args):
    global launches
    rc = getattr(_lib, name)(*args)
    if rc:
        raise NavcError("%s: %s" % (name, _lib.navc_last_error().decode()))
    launches += 1


def ptr(t):
    return None if t is None else t.data_ptr()


def stream():
    """Raw handle of torch's current stream on the current device (the fast C entry point:
    torch.cuda.current_stream() costs ~15 us of Python per call, and every launch asks)."""
    return _raw_stream(_current_device())


try:
    _raw_stream = torch._C._cuda_getCurrentRawStream
    _current_device = torch._C._cuda_getDevice
except AttributeError:  # pragma: no cover - older / CPU-only torch builds
    def _raw_stream(_idx):
        return torch.cuda.current_stream().cuda_stream

    def _current_device():
        return 0
