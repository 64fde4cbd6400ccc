"""ctypes binding of libcassnat_b200.so (the C ABI declared in include/cassnat_b200.h).

There is no fallback of any kind: if the shared library is missing, or a call returns a
non-zero cn_status, a RuntimeError is raised.  PyTorch only supplies device memory and the
current CUDA stream.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcassnat_b200.so")

_i64, _i32, _f32, _vp, _int = ctypes.c_int64, ctypes.c_int32, ctypes.c_float, ctypes.c_void_p, ctypes.c_int

# name -> (restype, argtypes); mirrors include/cassnat_b200.h one to one
SIGNATURES = {
    "cn_version": (_int, []),
    "cn_status_string": (ctypes.c_char_p, [_int]),
    "cn_device_check": (_int, []),
    "cn_row_stats": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "cn_bestpath_collapse": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _f32, _int, _vp, _vp, _vp, _vp]),
    "cn_bestpath_workspace_bytes": (_i64, [_i64, _i64]),
    "cn_bestpath_from_logp": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _i64, _i64, _f32, _int,
                                     _vp, _vp, _vp, _vp, _i64, _vp]),
    "cn_viterbi_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "cn_viterbi_align": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _i64, _vp, _i64, _vp, _i64,
                                _vp, _vp, _i64, _vp]),
    "cn_gather_emissions": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _vp, _vp]),
    "cn_viterbi_align_compact": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _i64, _vp, _i64, _vp, _i64,
                                        _vp, _vp, _i64, _vp]),
    "cn_align_to_mask": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _int, _int, _int, _vp, _vp, _vp, _vp]),
    "cn_expand_trigger_mask": (_int, [_vp, _i64, _i64, _i64, _int, _int, _vp, _vp]),
    "cn_tgt_length_mask": (_int, [_vp, _i64, _i64, _vp, _vp, _vp]),
    "cn_gemm_f32": (_int, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i64, _f32, _int, _vp, _i64, _i64, _i64, _i64, _vp]),
    "cn_layernorm": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _f32, _vp, _vp, _i64, _vp]),
    "cn_attention_f32": (_int, [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64,
                                _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _i64, _i64, _vp]),
    "cn_segment_attention_f32": (_int, [_vp, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _int, _int,
                                        _i64, _i64, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _vp]),
    "cn_segment_attention_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "cn_conv_module_workspace_bytes": (_i64, [_i64, _i64]),
    "cn_glu_dwconv_groupnorm_act": (_int, [_vp, _int, _i64, _i64, _vp, _vp, _vp, _vp, _f32, _int, _i64, _i64, _i64, _i64,
                                           _vp, _vp, _vp, _i64, _vp]),
    "cn_gemm_bf16": (_int, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i64, _f32, _int, _vp, _vp, _i64, _i64, _i64, _i64, _vp]),
    "cn_cast_bf16": (_int, [_vp, _vp, _i64, _vp]),
    "cn_gemm_rowpart_count": (_i64, [_i64, _i64]),
    "cn_gemm_bf16_rowpart": (_int, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i64, _f32, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _vp]),
    "cn_gemm_bf16_folded": (_int, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _f32, _f32, _int, _vp, _vp, _i64, _i64, _i64, _i64, _vp]),
    "cn_gemm_bf16_ln": (_int, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i64, _f32, _vp, _i64, _vp, _vp, _f32, _vp, _i64, _i64, _i64, _i64, _vp]),
    "cn_gemm_bf16x3": (_int, [_vp, _i64, _int, _vp, _i64, _vp, _vp, _i64, _i64, _f32, _int, _vp, _vp, _int, _i64, _i64, _i64, _i64, _vp]),
    "cn_cast_bf16_split": (_int, [_vp, _i64, _vp, _i64, _i64, _i64, _vp]),
    "cn_layernorm_bf16x2": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _f32, _vp, _i64, _vp, _i64, _vp]),
    "cn_gemm_bf16x3_rowstats": (_int, [_vp, _i64, _vp, _i64, _vp, _f32, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "cn_gemm_rowstats_workspace_bytes": (_i64, [_i64, _i64]),
    "cn_gemm_bf16_rowstats": (_int, [_vp, _i64, _vp, _i64, _vp, _f32, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "cn_rel_attention_workspace_bytes": (_i64, [_i64, _i64, _i64, _i64]),
    "cn_rel_attention_bf16": (_int, [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64,
                                     _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _vp, _vp]),
    "cn_rel_attention_bf16_table": (_int, [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64,
                                           _vp, _int, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _i64, _i64,
                                           _vp, _vp]),
    "cn_set_sm_limit": (_int, [_int]),
    "cn_rescore_select": (_int, [_vp, _i64, _i64, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "cn_ctc_loss_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "cn_ctc_loss": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _vp, _i64, _f32, _vp, _vp,
                           _vp, _i64, _i64, _vp, _i64, _vp]),
    "cn_attention_bf16": (_int, [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64,
                                 _i64, _i64, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "cn_gemm_bf16_ce_grad": (_int, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _f32, _f32, _vp, _i64, _i64, _i64, _i64, _vp]),
    "cn_attention_bf16x3": (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _i64, _i64, _vp, _vp, _i64, _i64, _vp, _i64, _i64, _i64, _i64,
                                   _i64, _i64, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "cn_conv_frontend_bf16": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "cn_decode_hints": (_int, [_vp, _i64, _vp, _i64, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cn_pack_mask_bits": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp]),
    "cn_trigger_mask_bits": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _int, _int, _vp, _vp]),
    "cn_attention_bf16_bits": (_int, [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _i64, _i64, _i64,
                                      _vp, _i64, _i64, _vp, _vp, _int, _vp]),
    "cn_gemm_f32_strided": (_int, [_vp, _i64, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _i64, _vp, _i64, _i64, _i64,
                                   _i64, _i64, _i64, _i64, _i64, _f32, _f32, _vp]),
    "cn_colsum_workspace_bytes": (_i64, [_i64, _i64]),
    "cn_act_bwd_colsum_bf16": (_int, [_vp, _int, _i64, _vp, _int, _i64, _int, _f32, _vp, _i64, _i64, _i64, _vp, _int, _vp, _i64, _vp]),
    "cn_act_bwd_colsum": (_int, [_vp, _i64, _vp, _i64, _int, _f32, _vp, _i64, _i64, _i64, _vp, _int, _vp, _i64, _vp]),
    "cn_layernorm_bwd_workspace_bytes": (_i64, [_i64, _i64]),
    "cn_layernorm_bwd": (_int, [_vp, _i64, _vp, _i64, _vp, _f32, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _vp]),
    "cn_attention_bwd_small": (_int, [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _int, _vp,
                                      _i64, _i64, _i64, _i64, _i64, _i64, _i64, _f32, _vp, _i64, _i64, _vp, _i64, _i64,
                                      _vp, _i64, _i64, _vp, _i64, _i64, _vp]),
    "cn_attention_fwd_small": (_int, [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _int, _vp, _i64, _i64, _i64, _i64, _i64, _i64,
                                      _i64, _f32, _vp, _i64, _i64, _vp, _i64, _i64, _vp]),
    "cn_attn_softmax_bwd": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _int, _vp]),
    "cn_dropout": (_int, [_vp, _int, _vp, _int, _i64, _f32, _vp, _i64, _i64, _vp]),
    "cn_dropout_add": (_int, [_vp, _int, _vp, _vp, _i64, _f32, _f32, _vp, _i64, _i64, _vp]),
    "cn_label_smoothing": (_int, [_vp, _i64, _i64, _i64, _vp, _i64, _f32, _f32, _vp, _vp, _i64, _vp, _i64, _vp]),
    "cn_gemm_bf16_mn": (_int, [_vp, _i64, _int, _vp, _i64, _int, _vp, _vp, _i64, _i64, _i64, _i64, _vp]),
    "cn_transpose_bf16": (_int, [_vp, _int, _i64, _i64, _i64, _vp, _i64, _vp]),
    "cn_rel_bias_add": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _f32, _vp]),
    "cn_rel_bucket": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _vp]),
    "cn_act_fwd": (_int, [_vp, _int, _vp, _vp, _i64, _vp]),
    "cn_conv_module_bwd_workspace_bytes": (_i64, [_i64, _i64, _i64, _i64]),
    "cn_glu_dwconv_groupnorm_act_bwd": (_int, [_vp, _int, _i64, _i64, _vp, _vp, _vp, _vp, _f32, _int, _i64, _i64, _i64, _i64,
                                               _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
}

_lib = None


def load():
    """Returns the loaded library, raising (never falling back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m cassnat_asr_public_b200.build` "
            "(nvcc, sm_100a). cassnat_asr_public_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def status_string(rc):
    return load().cn_status_string(rc).decode()


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


# kernels launched per C-ABI call (cn_row_stats: +1 when it also writes log-probs)
LAUNCHES = {"cn_row_stats": 1, "cn_bestpath_collapse": 1, "cn_bestpath_from_logp": 2, "cn_viterbi_align": 1,
            "cn_align_to_mask": 1, "cn_expand_trigger_mask": 1, "cn_tgt_length_mask": 1, "cn_gemm_f32": 1,
            "cn_layernorm": 1, "cn_attention_f32": 1, "cn_segment_attention_f32": 3,
            "cn_glu_dwconv_groupnorm_act": 3, "cn_gemm_bf16_rowstats": 2, "cn_gemm_bf16x3_rowstats": 2, "cn_rel_attention_bf16": 2,
            "cn_act_bwd_colsum": 2, "cn_act_bwd_colsum_bf16": 2, "cn_layernorm_bwd": 3, "cn_label_smoothing": 3, "cn_ctc_loss": 4,
            "cn_glu_dwconv_groupnorm_act_bwd": 7, "cn_conv_frontend_bf16": 2}
launch_count = 0          # running total of our kernel launches (bench.py reads and resets it)
_profile = None           # list of (name, start_event, end_event, work) while profiling
PROFILE_SHAPES = False     # tag GEMM records with their shape (diagnostics)


def start_profile():
    global _profile
    _profile = []


def stop_profile():
    """-> {name: dict(calls, ms, flops, bytes)} measured with CUDA events on the launching stream."""
    global _profile
    rec, _profile = _profile, None
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1, work in rec:
        d = out.setdefault(name, dict(calls=0, ms=0.0, flops=0.0, bytes=0.0))
        d["calls"] += 1
        d["ms"] += e0.elapsed_time(e1)
        d["flops"] += work.get("flops", 0.0)
        d["bytes"] += work.get("bytes", 0.0)
    return out


def call(name, *args, work=None, tag=None):
    global launch_count
    fn = getattr(load(), name)
    if _profile is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        _profile.append((tag or name, e0, e1, work or {}))
    else:
        rc = fn(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed: {status_string(rc)} (cn_status {rc})")
    launch_count += LAUNCHES.get(name, 1)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("cassnat_asr_public_b200 ops need CUDA tensors (no CPU fallback)")
