"""ctypes binding of libvbg_sm100a.so (include/vbg.h).  Fails loudly when the
library is absent: there is no eager / CPU fallback for the hot path."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvbg_sm100a.so")

VBG_OK, VBG_EINVAL, VBG_ECUDA, VBG_EUNSUPPORTED, VBG_EWORKSPACE = 0, -1, -2, -3, -4
PREC_FP32, PREC_TF32, PREC_BF16X3 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
RES_NONE, RES_SAME, RES_UP2 = 0, 1, 2
AGG_MEAN, AGG_FIRST = 0, 1
OUT_F32, OUT_SPLIT_BF16 = 0, 1


class Epilogue(C.Structure):
    _fields_ = [("scale", C.c_void_p), ("shift", C.c_void_p), ("residual", C.c_void_p),
                ("res_mode", C.c_int), ("ldr", C.c_int), ("out_h", C.c_int), ("out_w", C.c_int),
                ("act", C.c_int), ("out_mode", C.c_int), ("out_plane", C.c_longlong), ("res_plane", C.c_longlong),
                ("tune", C.c_int)]


class VbgError(RuntimeError):
    pass


_p, _i, _f, _sz, _ll = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_longlong
_EP = C.POINTER(Epilogue)

# name -> argtypes, in include/vbg.h order
SIGNATURES = {
    "vbg_version": [],
    "vbg_last_error": [C.c_char_p, _sz],
    "vbg_tc_available": [],
    "vbg_normalize_resize_pad": [_p, _i, _i, _p, _i, _i, _i, _i, _i, C.POINTER(_f), C.POINTER(_f), _p],
    "vbg_normalize_resize_pad_batch": [_p, _i, _i, _i, _p, _i, _i, _i, _i, _i, C.POINTER(_f), C.POINTER(_f), _p],
    "vbg_normalize_resize_pad_u8": [_p, _i, _i, _i, _p, _i, _i, _i, _i, _i, C.POINTER(_f), C.POINTER(_f), _p],
    "vbg_decode_batch_u8": [_p, _p, _i, _i, _i, _p, _i, _i, C.POINTER(_f), C.POINTER(_f), _p],
    "vbg_resize_coords": [_p, _p, _p, _i, _i, _p, _p],
    "vbg_bert_assemble": [_p, _i, _p, _p, _i, _i, _p, _p, _p],
    "vbg_embed_ln": [_p, _p, _p, _p, _p, _p, _p, _f, _i, _i, _i, _i, _p, _p],
    "vbg_layernorm": [_p, _p, _p, _f, _i, _i, _p, _p],
    "vbg_attention_fwd": [_p, _p, _i, _i, _i, _i, _p, _i, _p],
    "vbg_embed_ln_x": [_p, _p, _p, _p, _p, _p, _p, _f, _i, _i, _i, _i, _p, _ll, _p],
    "vbg_layernorm_x": [_p, _p, _p, _f, _i, _i, _p, _ll, _p],
    "vbg_attention_split_fwd": [_p, _ll, _p, _i, _i, _i, _i, _i, _p, _ll, _p],
    "vbg_segment_starts": [_p, _p, _i, _i, _i, _p, _p, _p],
    "vbg_mask_check": [_p, _i, _i, _p, _p, _p],
    "vbg_segment_reduce": [_p, _p, _p, _i, _i, _i, _p, _p],
    "vbg_box_index_map": [_p, _p, _i, _i, _i, _i, _p, _p],
    "vbg_grid_scatter": [_p, _p, _p, _i, _i, _i, _p, _p],
    "vbg_grid_scatter_x": [_p, _ll, _p, _p, _i, _i, _i, _p, _ll, _p],
    "vbg_label_paint": [_p, _p, _p, _i, _i, _i, _p, _p, _p],
    "vbg_seg_ce_loss": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _sz, _p, _p],
    "vbg_gemm": [_p, _i, _p, _i, _i, _p, _i, _p, _ll, _p, _i, _i, _i, _i, _EP, _i, _p],
    "vbg_gemm_ps": [_p, _ll, _i, _p, _ll, _i, _i, _p, _ll, _i, _p, _i, _i, _i, _i, _EP, _p, _sz, _p],
    "vbg_conv2d_ps": [_p, _ll, _i, _i, _i, _i, _p, _ll, _i, _i, _i, _i, _i, _p, _EP, _p, _sz, _p],
    "vbg_gemm_ps_workspace": [_i, _i, _i, _i],
    "vbg_conv2d_ps_workspace": [_i, _i, _i, _i, _i, _i, _i, _i, _i, _i],
    "vbg_debug_set_timeline": [_p],
    "vbg_merge_bf16": [_p, _p, _ll, _p, _p],
    "vbg_conv2d": [_p, _i, _i, _i, _i, _p, _p, _ll, _i, _i, _i, _i, _i, _p, _EP, _i, _p],
    "vbg_split_bf16": [_p, _ll, _p, _p, _p],
    "vbg_stem_conv": [_p, _i, _i, _i, _p, _p, _ll, _i, _p, _EP, _i, _p],
    "vbg_stem_pack_weights": [_p, _i, _p, _p, _p],
    "vbg_maxpool3x3s2": [_p, _i, _i, _i, _i, _p, _p],
    "vbg_avgpool2x2": [_p, _i, _i, _i, _i, _p, _p],
    "vbg_maxpool3x3s2_x": [_p, _ll, _i, _i, _i, _i, _p, _ll, _p],
    "vbg_avgpool2x2_x": [_p, _ll, _i, _i, _i, _i, _p, _ll, _p],
    "vbg_bn_fold": [_p, _p, _p, _p, _f, _i, _p, _p, _p],
    "vbg_repack_oihw_to_ohwi": [_p, _i, _i, _i, _i, _p, _p],
    "vbg_roi_align_fwd": [_p, _i, _i, _i, _i, _p, _p, _i, _f, _i, _p, _p, _p],
    "vbg_roi_align_x": [_p, _ll, _i, _i, _i, _i, _p, _p, _i, _f, _i, _p, _ll, _p, _p],
    "vbg_roi_align_sel": [_p, _ll, _i, _i, _i, _i, _p, _p, _i, _f, _i, _p, _ll, _p, _i, _p],
    "vbg_transpose_split": [_p, _ll, _i, _i, _p, _ll, _i, _p],
    "vbg_colsum_workspace": [_ll, _i],
    "vbg_colsum": [_p, _ll, _ll, _i, _p, _p, _sz, _p],
    "vbg_linear_wgrad_workspace": [_i, _i, _i],
    "vbg_conv2d_wgrad_workspace": [_i, _i, _i, _i, _i, _i, _i, _i, _i],
    "vbg_conv2d_wgrad": [_p, _ll, _p, _ll, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _sz, _p],
    "vbg_linear_wgrad": [_p, _ll, _p, _ll, _i, _i, _i, _p, _p, _sz, _p],
    "vbg_conv_dgrad_weight": [_p, _i, _i, _i, _i, _p, _ll, _p],
    "vbg_layernorm_bwd": [_p, _p, _p, _f, _i, _i, _p, _p, _p, _p, _sz, _p],
    "vbg_bn_workspace": [_ll, _i],
    "vbg_bn_stats": [_p, _ll, _i, _f, _p, _p, _p, _p, _sz, _p],
    "vbg_bn_apply": [_p, _ll, _i, _p, _p, _p, _p, _p, _i, _p, _p],
    "vbg_bn_bwd": [_p, _p, _p, _ll, _i, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p],
    "vbg_bn_bwd_reduce": [_p, _p, _p, _ll, _i, _p, _p, _p, _p, _p, _sz, _p],
    "vbg_bn_bwd_dx": [_p, _p, _p, _ll, _i, _f, _p, _p, _p, _p, _p, _p, _p, _p],
    "vbg_maxpool3x3s2_bwd": [_p, _p, _i, _i, _i, _i, _p, _p],
    "vbg_maxpool3x3s2_bwd_y": [_p, _p, _p, _i, _i, _i, _i, _p, _p],
    "vbg_sumpool2x2": [_p, _i, _i, _i, _i, _f, _p, _p],
    "vbg_expand2x": [_p, _i, _i, _i, _i, _i, _i, _f, _i, _p, _p],
    "vbg_gelu": [_p, _p, _ll, _p, _p],
    "vbg_gelu_x": [_p, _ll, _p, _ll, _ll, _p, _ll, _p],
    "vbg_dropout": [_p, _ll, _f, C.c_ulonglong, _p, _p],
    "vbg_uniform_keys": [_ll, C.c_ulonglong, _p, _p, _p],
    "vbg_dropout_ds": [_p, _ll, _f, C.c_ulonglong, _p, _p, _p],
    "vbg_grid_scatter_bwd": [_p, _ll, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p],
    "vbg_segment_reduce_bwd": [_p, _p, _p, _i, _i, _i, _p, _p],
    "vbg_embed_bwd": [_p, _p, _p, _i, _i, _p, _p, _p],
    "vbg_roi_align_bwd_workspace": [_i],
    "vbg_roi_align_bwd": [_p, _i, _i, _i, _i, _p, _p, _i, _f, _i, _p, _p, _sz, _p],
    "vbg_seg_ce_bwd": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p],
    "vbg_upsample_split_bwd": [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p],
    "vbg_small_wgrad_workspace": [_ll, _i, _i],
    "vbg_small_wgrad": [_p, _i, _p, _i, _ll, _i, _i, _p, _p, _sz, _p],
    "vbg_stem_wgrad_workspace": [],
    "vbg_stem_wgrad": [_p, _p, _i, _i, _i, _i, _i, _p, _p, _sz, _p],
    "vbg_attention_bwd": [_p, _p, _p, _p, _i, _i, _i, _i, _ll, _p, _p, _sz, _p],
    "vbg_attention_split_train_fwd": [_p, _ll, _p, _i, _i, _i, _i, _i, _p, _ll, _p, _f, C.c_ulonglong, _p, _p],
    "vbg_attention_bwd_tc": [_p, _ll, _p, _ll, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, C.c_ulonglong, _p, _p, _p, _sz, _p],
    "vbg_attention_dropout_mask": [C.c_ulonglong, C.c_ulonglong, _i, _f, _i, _i, _i, _p, C.POINTER(_f), _p],
    "vbg_optim_chunk": [],
    "vbg_sgd_step_mt": [_p, _i, _ll, _f, _f, _f, _i, _f, _p],
    "vbg_adamw_step_mt": [_p, _i, _ll, _f, _f, _f, _f, _f, _f, _f, _f, _p],
    "vbg_softmax_rows": [_p, _i, _i, _p, _p],
    "vbg_full_head_scores": [_p, _p, _i, _i, _p, _p],
    "vbg_upsample_split_nchw": [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p],
    "vbg_nhwc_to_nchw": [_p, _i, _i, _i, _i, _p, _p],
    "vbg_crf_viterbi": [_p, _p, _p, _i, _i, _i, _p, _p, _p, _sz, _p],
    "vbg_crf_nll_fwd": [_p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p],
    "vbg_crf_nll_bwd": [_p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p],
    "vbg_shard_open": [C.c_char_p, C.POINTER(_p)],
    "vbg_shard_close": [_p],
    "vbg_shard_num_docs": [_p],
    "vbg_shard_doc_shape": [_p, _i, _p],
    "vbg_shard_doc_meta": [_p, _i, C.POINTER(_p), C.POINTER(_ll)],
    "vbg_shard_batch_layout": [_p, _p, _i, _p],
    "vbg_shard_collate": [_p, _p, _i, _p, _sz, _i],
}

_lib = None


def load():
    """Load (once) and type the library.  No GPU is needed to load it."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the ViBERTgrid hot path has no fallback. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or python vibertgrid-pytorch_b200/csrc/build.py)")
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == ABI drift, by design
        fn.argtypes = args
        fn.restype = C.c_longlong if name.endswith("_workspace") else C.c_int
    _lib = lib
    return lib


def last_error() -> str:
    buf = C.create_string_buffer(512)
    load().vbg_last_error(buf, 512)
    return buf.value.decode(errors="replace")


launch_count = 0     # kernels enqueued through the C-ABI (bench.py reports it as gpu_launches)


def check(rc: int, what: str = "", launch: bool = True):
    """``launch=False`` for host-only entry points (the shard reader): they enqueue no kernel."""
    global launch_count
    if launch:
        launch_count += 1
    if rc != VBG_OK:
        raise VbgError(f"{what or 'libvbg'} failed (code {rc}): {last_error()}")
