"""ctypes binding of ``libovdet.so`` - the C ABI declared in ``include/ovdet.h``.

The library is the product: there is no CPU or PyTorch fallback.  ``lib()`` raises when the
shared object is missing, and every compute entry returns ``OVDET_ERR_WRONG_ARCH`` on a device
that is not compute capability 10.x; ``check()`` turns any non-zero status into ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libovdet.so")

OVDET_F32, OVDET_BF16 = 0, 1
ACT_NONE, ACT_SIGMOID = 0, 1
MAX_LEVELS = 8

class HeadStepArgs(ctypes.Structure):
    """``ovdet_head_step_args`` of include/ovdet.h, field for field."""
    _fields_ = [
        ("num_levels", c_int32), ("bins", c_int32),
        ("batch", c_int64), ("dim", c_int64), ("classes", c_int64),
        ("obj_embeds", c_void_p * 4), ("box_preds", c_void_p * 4),
        ("heights", c_int32 * 4), ("widths", c_int32 * 4), ("strides", c_int32 * 4),
        ("emb_stride_b", c_int64 * 4), ("emb_stride_d", c_int64 * 4), ("box_stride_b", c_int64 * 4),
        ("text_op", c_void_p), ("text_batched", c_int32), ("activation", c_int32),
        ("class_aware", c_int32), ("topk", c_int32), ("embed_dtype", c_int32),
        ("box_dtype", c_int32),
        ("alpha", c_float), ("beta", c_float), ("conf", c_float), ("iou_thr", c_float),
        ("max_det", c_int64),
        ("scores", c_void_p), ("class_ids", c_void_p), ("inv_norm", c_void_p), ("boxes", c_void_p),
        ("scores_act", c_void_p), ("pass_mask", c_void_p), ("scale", c_void_p), ("clip_wh", c_void_p),
        ("out_boxes", c_void_p), ("out_scores", c_void_p), ("out_classes", c_void_p),
        ("out_anchor", c_void_p), ("out_keep", c_void_p), ("out_count", c_void_p),
        ("out_candidates", c_void_p), ("workspace", c_void_p), ("workspace_bytes", c_size_t),
        ("sim_workspace", c_void_p), ("sim_workspace_bytes", c_size_t),
        ("text_fp16", c_int32), ("reserved", c_int32),
    ]


# name -> (restype, argtypes); mirrors include/ovdet.h one to one
PROTOTYPES = {
    "ovdet_version": (c_int, []),
    "ovdet_strerror": (c_char_p, [c_int]),
    "ovdet_last_cuda_error": (c_int, []),
    "ovdet_check_device": (c_int, []),
    "ovdet_l2norm_regions": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64,
                                     c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "ovdet_l2norm_text": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64,
                                  c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "ovdet_similarity": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64,
                                 c_int, c_int, c_float, c_float, c_void_p, c_int, c_int64,
                                 c_void_p, c_void_p, c_void_p]),
    "ovdet_similarity_fused": (c_int, [POINTER(c_void_p), POINTER(c_int64), POINTER(c_int64),
                                       POINTER(c_int64), c_int, c_int64, c_int64, c_void_p, c_int64,
                                       c_int, c_float, c_float, c_void_p, c_int, c_int64,
                                       c_void_p, c_void_p, c_void_p, c_void_p]),
    "ovdet_similarity_split_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "ovdet_similarity_fused_ws": (c_int, [POINTER(c_void_p), POINTER(c_int64), POINTER(c_int64),
                                          POINTER(c_int64), c_int, c_int64, c_int64, c_void_p, c_int64,
                                          c_int, c_float, c_float, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_size_t, c_int, c_void_p]),
    "ovdet_similarity_fused_bf16in": (c_int, [POINTER(c_void_p), POINTER(c_int64), POINTER(c_int64),
                                              POINTER(c_int64), c_int, c_int64, c_int64, c_void_p, c_int64,
                                              c_int, c_float, c_float, c_void_p, c_int, c_int64,
                                              c_void_p, c_void_p, c_void_p, c_void_p]),
    "ovdet_similarity_fused_fp32": (c_int, [POINTER(c_void_p), POINTER(c_int64), POINTER(c_int64),
                                            POINTER(c_int64), c_int, c_int64, c_int64, c_void_p, c_int64,
                                            c_int, c_float, c_float, c_void_p, c_int, c_int64,
                                            c_void_p, c_void_p, c_void_p, c_void_p]),
    "ovdet_similarity_fused_fp16": (c_int, [POINTER(c_void_p), POINTER(c_int64), POINTER(c_int64),
                                            POINTER(c_int64), c_int, c_int64, c_int64, c_void_p, c_int64,
                                            c_int, c_float, c_float, c_void_p, c_int, c_int64,
                                            c_void_p, c_void_p, c_void_p, c_void_p]),
    "ovdet_similarity_projected": (c_int, [POINTER(c_void_p), POINTER(c_int64), POINTER(c_int64),
                                           POINTER(c_int64), c_int, c_int64, c_int64, POINTER(c_void_p),
                                           c_int64, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p,
                                           c_void_p]),
    "ovdet_rowmax": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "ovdet_repitch_rows": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int, c_void_p]),
    "ovdet_concat_embeddings": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p,
                                        c_int64, c_int64, c_void_p]),
    "ovdet_decode_filter": (c_int, [POINTER(c_void_p), POINTER(c_int32), POINTER(c_int32),
                                    POINTER(c_int32), POINTER(c_int64), c_int, c_int, c_int64,
                                    c_float, c_float, c_void_p, c_float, c_int,
                                    c_void_p, c_void_p, c_void_p, c_void_p]),
    "ovdet_decode_filter_bf16in": (c_int, [POINTER(c_void_p), POINTER(c_int32), POINTER(c_int32),
                                           POINTER(c_int32), POINTER(c_int64), c_int, c_int, c_int64,
                                           c_float, c_float, c_void_p, c_float, c_int,
                                           c_void_p, c_void_p, c_void_p, c_void_p]),
    "ovdet_nms_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "ovdet_nms_batched": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                  c_void_p, c_void_p, c_float, c_int, c_int, c_int64,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_size_t, c_void_p]),
    "ovdet_nms_batched_conf": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int64, c_int64,
                                       c_void_p, c_void_p, c_float, c_int, c_int, c_int64,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_size_t, c_void_p]),
    "ovdet_letterbox_u8": (c_int, [POINTER(c_void_p), POINTER(c_int32), POINTER(c_int32),
                                   POINTER(c_int64), POINTER(c_int32), POINTER(c_int32), c_int,
                                   c_int, c_int, c_void_p, c_void_p]),
    "ovdet_cast_text": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64,
                                c_int, c_void_p]),
    "ovdet_max_sigmoid_attention": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p,
                                            c_int64, c_int, c_int, c_void_p, c_void_p, c_int64, c_int64,
                                            c_void_p]),
    "ovdet_head_step": (c_int, [POINTER(HeadStepArgs), c_void_p]),
    "ovdet_head_step_args_size": (c_size_t, []),
    "ovdet_pack_boxes_i32": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    # vocabulary-parallel exchange (vocab_parallel.cu)
    "ovdet_peer_buffer_create": (c_int, [c_size_t, POINTER(c_void_p), c_void_p]),
    "ovdet_peer_buffer_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "ovdet_peer_buffer_close": (c_int, [c_void_p]),
    "ovdet_peer_buffer_destroy": (c_int, [c_void_p]),
    "ovdet_vp_buffer_bytes": (c_size_t, [c_int64, c_int]),
    "ovdet_vp_buffer_init": (c_int, [c_void_p, c_int64, c_int, c_void_p]),
    "ovdet_similarity_fused_vp": (c_int, [POINTER(c_void_p), POINTER(c_int64), POINTER(c_int64),
                                          POINTER(c_int64), c_int, c_int64, c_int64, c_void_p, c_int64,
                                          c_int, c_float, c_float, c_void_p, c_void_p, c_size_t, c_int,
                                          c_int64, POINTER(c_void_p), c_int, c_int, c_void_p]),
    "ovdet_vp_signal": (c_int, [POINTER(c_void_p), c_int, c_int, c_int64, c_void_p]),
    "ovdet_vp_wait_unpack": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p,
                                     c_int, c_void_p]),
    "ovdet_head_step_vp": (c_int, [POINTER(HeadStepArgs), c_int64, POINTER(c_void_p), c_int, c_int,
                                   c_void_p, c_int, c_void_p]),
    "ovdet_pack_score_keys": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "ovdet_unpack_score_keys": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load (once) and return the library; raise loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) and not os.environ.get("OVDET_LIB_PATH"):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m ovdet.build` (nvcc, sm_100a). "
                "ovdet has no CPU/PyTorch fallback.")
        # OVDET_LIB_PATH: A/B tooling only (tools/ab): load another build of the same library; symbols
        # it does not have yet are skipped
        override = os.environ.get("OVDET_LIB_PATH")
        handle = ctypes.CDLL(override or LIB_PATH)
        for name, (restype, argtypes) in PROTOTYPES.items():
            if override and not hasattr(handle, name):
                continue
            fn = getattr(handle, name)        # AttributeError if the ABI drifted
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


class OvdetError(RuntimeError):
    def __init__(self, status: int, where: str):
        handle = lib()
        msg = handle.ovdet_strerror(status).decode()
        if status == -4:
            msg += f" [cudaError {handle.ovdet_last_cuda_error()}]"
        super().__init__(f"{where}: {msg} (status {status})")
        self.status = status


def check(status: int, where: str) -> None:
    if status != 0:
        raise OvdetError(status, where)
