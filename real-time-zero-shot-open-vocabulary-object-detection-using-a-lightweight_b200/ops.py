"""Tensor-level wrappers over the C ABI (include/ovdet.h): one function per kernel family.

Every function enqueues on torch's current stream of the input's device and returns without a
host synchronisation.  Inputs must live on a CUDA device; there is no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch

from . import _cabi
from ._cabi import check, lib


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def _require_cuda(t: torch.Tensor, name: str, dtype=None):
    if not t.is_cuda:
        raise RuntimeError(f"ovdet: `{name}` must be a CUDA tensor (no CPU fallback exists)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"ovdet: `{name}` must be {dtype}, got {t.dtype}")


# --------------------------------------------------------------------------------------------
# K1: L2 norm + tensor-core operands
# --------------------------------------------------------------------------------------------
def l2norm_regions(obj_embeds: Sequence[torch.Tensor], split: bool = False,
                   operand: Optional[torch.Tensor] = None,
                   inv_norm: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """text_contrastive.py:134,137 for all levels at once.

    obj_embeds: per level fp32 ``[B, D, H, W]`` (NCHW).  Returns the bf16 operand
    ``[B, A, D * (2 if split else 1)]`` (A = sum H*W, levels concatenated P3|P4|P5 as
    yolo_clip.py:205 does) and ``inv_norm [B, A]`` = 1 / max(||x||, 1e-12).
    """
    first = obj_embeds[0]
    _require_cuda(first, "obj_embed", torch.float32)
    batch, dim = first.shape[0], first.shape[1]
    anchors = sum(e.shape[2] * e.shape[3] for e in obj_embeds)
    kop = dim * (2 if split else 1)
    if operand is None:
        operand = torch.empty(batch, anchors, kop, device=first.device, dtype=torch.bfloat16)
    if inv_norm is None:
        inv_norm = torch.empty(batch, anchors, device=first.device, dtype=torch.float32)
    assert operand.shape == (batch, anchors, kop) and operand.is_contiguous()
    assert inv_norm.shape == (batch, anchors) and inv_norm.is_contiguous()
    handle = lib()
    offset = 0
    with torch.cuda.device(first.device):
        for e in obj_embeds:
            _require_cuda(e, "obj_embed", torch.float32)
            b, d, h, w = e.shape
            assert b == batch and d == dim
            if e.stride(3) != 1 or e.stride(2) != w:
                e = e.contiguous()
            check(handle.ovdet_l2norm_regions(e.data_ptr(), b, d, h * w, e.stride(0), e.stride(1),
                                              operand.data_ptr(), anchors, offset, kop, int(split),
                                              inv_norm.data_ptr(), _stream(e)),
                  "ovdet_l2norm_regions")
            offset += h * w
    return operand, inv_norm


def shared_text(text: torch.Tensor) -> bool:
    """True when one vocabulary serves the whole batch: a ``[C, D]`` tensor, batch 1, or the
    stride-0 ``expand`` the offline-vocabulary path builds (model/yolo_clip.py:123)."""
    return text.dim() == 2 or text.shape[0] == 1 or text.stride(0) == 0


def l2norm_text(text: torch.Tensor, split=False,
                operand: Optional[torch.Tensor] = None) -> torch.Tensor:
    """text_contrastive.py:138.  ``text`` is ``[C, D]`` or ``[B, C, D]`` (any batch / row stride,
    as the neck emits it); a shared vocabulary is normalised once.  Returns the normalised
    bf16 operand ``[Bt, C, kop]`` with Bt = 1 for a shared vocabulary.  ``split``: False ``[hi]``,
    True ``[hi | lo]`` (two-kernel fp32 recipe), ``3`` ``[hi | lo | hi]`` (fused fp32-accurate mode),
    ``"fp16"``: ONE float16 segment holding 16 x the unit rows (the fp16 tier of ``similarity_fused``)."""
    _require_cuda(text, "text_embed", torch.float32)
    if text.dim() == 2:
        text = text.unsqueeze(0)
    elif shared_text(text):
        text = text[:1]
    if text.stride(2) != 1:
        text = text.contiguous()
    bt, classes, dim = text.shape
    f16 = isinstance(split, str)
    if f16 and split != "fp16":
        raise ValueError("ovdet: split is False, True, 3 or 'fp16'")
    mode = 3 if f16 else (2 if split == 3 else int(bool(split)))
    kop = dim if f16 else dim * (mode + 1)
    op_dtype = torch.float16 if f16 else torch.bfloat16
    if operand is None:
        operand = torch.empty(bt, classes, kop, device=text.device, dtype=op_dtype)
    assert operand.shape == (bt, classes, kop) and operand.is_contiguous() and operand.dtype == op_dtype
    with torch.cuda.device(text.device):
        check(lib().ovdet_l2norm_text(text.data_ptr(), bt, classes, dim, text.stride(0),
                                      text.stride(1), operand.data_ptr(), kop, mode, None,
                                      _stream(text)), "ovdet_l2norm_text")
    return operand


# --------------------------------------------------------------------------------------------
# K2: similarity GEMM (+ fused class max / argmax)
# --------------------------------------------------------------------------------------------
def similarity(regions_op: torch.Tensor, text_op: torch.Tensor, inv_norm: Optional[torch.Tensor],
               dim: int, alpha: float = 1.0, beta: float = 0.0, split: bool = False,
               logits_dtype: Optional[torch.dtype] = torch.float32, want_max: bool = False,
               logits: Optional[torch.Tensor] = None, row_max: Optional[torch.Tensor] = None,
               row_arg: Optional[torch.Tensor] = None):
    """text_contrastive.py:144,147 (+ yolo_clip.py:198-202 through the fused epilogue).

    Returns ``(logits [B, A, C] or None, row_max [B, A] or None, row_arg [B, A] int32 or None)``.
    """
    _require_cuda(regions_op, "regions_op", torch.bfloat16)
    _require_cuda(text_op, "text_op", torch.bfloat16)
    batch, rows, kop = regions_op.shape
    bt, classes, kop_t = text_op.shape
    assert kop == kop_t == dim * (2 if split else 1)
    assert bt in (1, batch)
    text_batched = int(bt == batch and batch > 1)
    dev = regions_op.device
    if logits is None and logits_dtype is not None:
        logits = torch.empty(batch, rows, classes, device=dev, dtype=logits_dtype)
    if want_max and row_max is None:
        row_max = torch.empty(batch, rows, device=dev, dtype=torch.float32)
    if want_max and row_arg is None:
        row_arg = torch.empty(batch, rows, device=dev, dtype=torch.int32)
    ldt = _cabi.OVDET_F32
    ldc = classes
    if logits is not None:
        assert logits.shape == (batch, rows, classes) and logits.stride(2) == 1
        assert logits.stride(0) == rows * logits.stride(1)
        ldc = logits.stride(1)
        ldt = {torch.float32: _cabi.OVDET_F32, torch.bfloat16: _cabi.OVDET_BF16}[logits.dtype]
    with torch.cuda.device(dev):
        check(lib().ovdet_similarity(regions_op.data_ptr(), text_op.data_ptr(), _ptr(inv_norm),
                                     batch, rows, classes, dim, int(split), text_batched,
                                     float(alpha), float(beta), _ptr(logits), ldt, ldc,
                                     _ptr(row_max), _ptr(row_arg), _stream(regions_op)),
              "ovdet_similarity")
    return logits, row_max, row_arg


def fused_supported(obj_embeds: Sequence[torch.Tensor]) -> bool:
    """Shapes the fused K1+K2 kernel takes (TMA alignment, <= 4 levels, dim <= 512); fp32 levels,
    or bf16 levels (all of them) for the bf16-activation variant."""
    if len(obj_embeds) > 4:
        return False
    dt = obj_embeds[0].dtype
    if dt not in (torch.float32, torch.bfloat16):
        return False
    per16 = 16 // obj_embeds[0].element_size()
    for e in obj_embeds:
        b, d, h, w = e.shape
        if e.dtype != dt or d % 64 or d > 512:
            return False
        if e.stride(3) != 1 or e.stride(2) != w:
            return False
        if e.stride(1) % per16 or e.stride(0) % per16 or e.data_ptr() % 16 or e.stride(1) < h * w:
            return False
    return True


def tma_addressable(obj_embeds: Sequence[torch.Tensor], buffers: Optional[list] = None) -> list:
    """Levels as the fused kernel's TMA loads can address them.  A level whose rows are not 16-byte
    aligned (H*W not a multiple of 4 for fp32 / 8 for bf16: 13x13, 15x15, 19x19 ... at image sizes such
    as 416, 480, 608) or that is not row-contiguous is copied once into a buffer with padded rows
    (``ovdet_repitch_rows``) and handed on as a ``[B, D, H, W]`` view of that buffer; aligned levels pass
    through untouched.  ``buffers``: a list the caller keeps (one slot per level) so that a steady-state
    step allocates nothing."""
    out = []
    for l, e in enumerate(obj_embeds):
        if e.dtype not in (torch.float32, torch.bfloat16) or e.dim() != 4 or not e.is_cuda:
            out.append(e)
            continue
        b, d, h, w = e.shape
        per16 = 16 // e.element_size()
        ok = (e.stride(3) == 1 and e.stride(2) == w and e.stride(1) % per16 == 0 and e.stride(0) % per16 == 0
              and e.data_ptr() % 16 == 0 and e.stride(1) >= h * w)
        if ok or b == 0:
            out.append(e)
            continue
        if e.stride(3) != 1 or e.stride(2) != w or e.stride(0) != d * e.stride(1):
            e = e.contiguous()                     # exotic views: one library copy, then the re-pitch below
        pitch = (h * w + per16 - 1) // per16 * per16
        buf = buffers[l] if buffers is not None and l < len(buffers) else None
        if buf is None or buf.shape != (b, d, pitch) or buf.dtype != e.dtype or buf.device != e.device:
            buf = torch.empty(b, d, pitch, device=e.device, dtype=e.dtype)
            if buffers is not None:
                while len(buffers) <= l:
                    buffers.append(None)
                buffers[l] = buf
        with torch.cuda.device(e.device):
            check(lib().ovdet_repitch_rows(e.data_ptr(), b * d, h * w, e.stride(1), buf.data_ptr(), pitch,
                                           e.element_size(), _stream(e)), "ovdet_repitch_rows")
        out.append(buf.as_strided((b, d, h, w), (d * pitch, pitch, w, 1)))
    return out


def text_operand_fp32(text: torch.Tensor) -> torch.Tensor:
    """Unit-norm text rows as the ``[hi | lo | hi]`` bf16 operand of ``similarity_fused(fp32=True)``
    (one K1b launch)."""
    return l2norm_text(text, split=3)


def fused_fp16_supported(obj_embeds: Sequence[torch.Tensor]) -> bool:
    """Shapes of the fp16 tier: the CTA-pair kernel (dim = 512), fp32 activations, TMA-addressable levels."""
    return (fused_supported(obj_embeds) and obj_embeds[0].shape[1] == 512 and
            obj_embeds[0].dtype == torch.float32)


def fused_fp32_supported(obj_embeds: Sequence[torch.Tensor], classes: int) -> bool:
    """Shapes of the fused fp32-accurate kernel: TMA-addressable levels and a single class tile
    (or an embedding of at most 128 values)."""
    return fused_supported(obj_embeds) and (classes <= 128 or obj_embeds[0].shape[1] <= 128)


def similarity_fused(obj_embeds: Sequence[torch.Tensor], text_op: torch.Tensor, alpha: float = 1.0,
                     beta: float = 0.0, logits_dtype: Optional[torch.dtype] = None,
                     want_max: bool = True, logits: Optional[torch.Tensor] = None,
                     row_max: Optional[torch.Tensor] = None, row_arg: Optional[torch.Tensor] = None,
                     inv_norm: Optional[torch.Tensor] = None, want_arg: bool = True, fp32: bool = False):
    """text_contrastive.py:134-147 for all levels + yolo_clip.py:198-206 in one launch, reading
    the fp32 NCHW ``obj_embeds`` directly.  ``text_op`` comes from ``l2norm_text(split=False)`` - or
    from ``l2norm_text(split="fp16")``: a float16 operand selects the fp16 tensor-core tier (same speed,
    |dlogit| ~ 1e-5 instead of ~4e-3).  Returns ``(logits [B, A, C] or None, row_max, row_arg)``;
    ``want_arg=False`` skips the argmax (scores only)."""
    first = obj_embeds[0]
    _require_cuda(first, "obj_embed")
    f16 = text_op.dtype == torch.float16
    _require_cuda(text_op, "text_op", torch.float16 if f16 else torch.bfloat16)
    if f16 and (fp32 or not fused_fp16_supported(obj_embeds)):
        raise ValueError("ovdet: the fp16 tier takes fp32 activations with dim = 512 (and is not the three-pass mode)")
    in16 = first.dtype == torch.bfloat16
    if in16 and fp32:
        raise ValueError("ovdet: the fp32-accurate mode takes fp32 activations")
    if not fused_supported(obj_embeds):
        raise ValueError("ovdet: shapes/strides not supported by the fused kernel "
                         "(use l2norm_regions + similarity)")
    batch, dim = first.shape[0], first.shape[1]
    n = len(obj_embeds)
    anchors = sum(e.shape[2] * e.shape[3] for e in obj_embeds)
    bt, classes, kop = text_op.shape
    assert kop == dim * (3 if fp32 else 1) and bt in (1, batch)
    if fp32 and not fused_fp32_supported(obj_embeds, classes):
        raise ValueError("ovdet: the fused fp32-accurate kernel takes at most 128 classes (or dim <= 128)")
    dev = first.device
    if logits is None and logits_dtype is not None:
        logits = torch.empty(batch, anchors, classes, device=dev, dtype=logits_dtype)
    if want_max and row_max is None:
        row_max = torch.empty(batch, anchors, device=dev, dtype=torch.float32)
    if want_max and want_arg and row_arg is None:
        row_arg = torch.empty(batch, anchors, device=dev, dtype=torch.int32)
    ldt, ldc = _cabi.OVDET_F32, classes
    if logits is not None:
        assert logits.shape == (batch, anchors, classes) and logits.stride(2) == 1
        assert logits.stride(0) == anchors * logits.stride(1)
        ldc = logits.stride(1)
        ldt = {torch.float32: _cabi.OVDET_F32, torch.bfloat16: _cabi.OVDET_BF16}[logits.dtype]
    ptrs = (ctypes.c_void_p * n)(*[e.data_ptr() for e in obj_embeds])
    hw = (ctypes.c_int64 * n)(*[e.shape[2] * e.shape[3] for e in obj_embeds])
    sb = (ctypes.c_int64 * n)(*[e.stride(0) for e in obj_embeds])
    sd = (ctypes.c_int64 * n)(*[e.stride(1) for e in obj_embeds])
    entry = (lib().ovdet_similarity_fused_fp32 if fp32 else lib().ovdet_similarity_fused_fp16 if f16 else
             lib().ovdet_similarity_fused_bf16in if in16 else lib().ovdet_similarity_fused)
    with torch.cuda.device(dev):
        check(entry(ptrs, hw, sb, sd, n, batch, dim, text_op.data_ptr(),
                    classes, int(bt == batch and batch > 1), float(alpha),
                    float(beta), _ptr(logits), ldt, ldc, _ptr(row_max),
                    _ptr(row_arg), _ptr(inv_norm), _stream(first)),
              "ovdet_similarity_fused_fp32" if fp32 else "ovdet_similarity_fused")
    return logits, row_max, row_arg


def rowmax(logits: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """yolo_clip.py:198-202 on materialised ``[..., C]`` logits (class axis contiguous).
    Ties resolve to the lowest class index."""
    _require_cuda(logits, "logits")
    assert logits.stride(-1) == 1
    flat = logits.reshape(-1, logits.shape[-1])
    rows, classes = flat.shape
    ldt = {torch.float32: _cabi.OVDET_F32, torch.bfloat16: _cabi.OVDET_BF16}[flat.dtype]
    out_max = torch.empty(rows, device=flat.device, dtype=torch.float32)
    out_arg = torch.empty(rows, device=flat.device, dtype=torch.int32)
    with torch.cuda.device(flat.device):
        check(lib().ovdet_rowmax(flat.data_ptr(), ldt, rows, classes, flat.stride(0),
                                 out_max.data_ptr(), out_arg.data_ptr(), _stream(flat)), "ovdet_rowmax")
    return out_max.reshape(logits.shape[:-1]), out_arg.reshape(logits.shape[:-1])


def concat_embeddings(obj_embeds: Sequence[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """yolo_clip.py:208-214: per level fp32 ``[B, D, H, W]`` -> ``[B, sum HW, D]`` (levels
    concatenated P3|P4|P5, anchor index ``i * W + j`` inside a level), one launch per level."""
    first = obj_embeds[0]
    _require_cuda(first, "obj_embed", torch.float32)
    batch, dim = first.shape[0], first.shape[1]
    anchors = sum(e.shape[2] * e.shape[3] for e in obj_embeds)
    if out is None:
        out = torch.empty(batch, anchors, dim, device=first.device, dtype=torch.float32)
    assert out.shape == (batch, anchors, dim) and out.is_contiguous() and out.dtype == torch.float32
    offset = 0
    with torch.cuda.device(first.device):
        for e in obj_embeds:
            _require_cuda(e, "obj_embed", torch.float32)
            b, d, h, w = e.shape
            assert b == batch and d == dim
            if e.stride(3) != 1 or e.stride(2) != w:
                e = e.contiguous()
            check(lib().ovdet_concat_embeddings(e.data_ptr(), b, d, h * w, e.stride(0), e.stride(1),
                                                out.data_ptr(), anchors, offset, _stream(e)),
                  "ovdet_concat_embeddings")
            offset += h * w
    return out


# --------------------------------------------------------------------------------------------
# K3: DFL decode + activation + confidence threshold
# --------------------------------------------------------------------------------------------
class LevelTable:
    """Host-side arrays describing the pyramid levels for ovdet_decode_filter."""

    def __init__(self, box_preds: Sequence[torch.Tensor], strides: Sequence[int]):
        n = len(box_preds)
        if n > _cabi.MAX_LEVELS or n != len(strides):
            raise ValueError("ovdet: unsupported number of levels")
        self.keep = []
        self.dtype = box_preds[0].dtype
        if self.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("ovdet: box_preds must be float32 or bfloat16")
        for p in box_preds:
            _require_cuda(p, "box_preds", self.dtype)
            b, ch, h, w = p.shape
            if p.stride(3) != 1 or p.stride(2) != w or p.stride(1) != h * w:
                p = p.contiguous()
            self.keep.append(p)
        self.n = n
        self.batch = self.keep[0].shape[0]
        self.channels = self.keep[0].shape[1]
        self.anchors = sum(p.shape[2] * p.shape[3] for p in self.keep)
        self.ptrs = (ctypes.c_void_p * n)(*[p.data_ptr() for p in self.keep])
        self.heights = (ctypes.c_int32 * n)(*[p.shape[2] for p in self.keep])
        self.widths = (ctypes.c_int32 * n)(*[p.shape[3] for p in self.keep])
        self.strides = (ctypes.c_int32 * n)(*[int(s) for s in strides])
        self.bstrides = (ctypes.c_int64 * n)(*[p.stride(0) for p in self.keep])


def decode_filter(box_preds: Sequence[torch.Tensor], strides: Sequence[int],
                  scores: Optional[torch.Tensor] = None, conf: float = 0.25,
                  activation: str = "none", width_scale: float = 1.0, height_scale: float = 1.0,
                  boxes: Optional[torch.Tensor] = None, scores_act: Optional[torch.Tensor] = None,
                  pass_mask: Optional[torch.Tensor] = None, want_boxes: bool = True):
    """box_head.py:150-218 (+ detector.py:184 when ``scores`` is given).

    Returns ``(boxes [B, A, 4] fp32, scores_act or None, pass_mask [B, ceil(A/32)] int32 or None)``.
    """
    table = LevelTable(box_preds, strides)
    dev = table.keep[0].device
    batch, anchors = table.batch, table.anchors
    if table.channels % 4:
        raise ValueError("ovdet: box_preds channels must be 4 * (reg_max + 1)")
    bins = table.channels // 4
    act = {"none": _cabi.ACT_NONE, "sigmoid": _cabi.ACT_SIGMOID}[activation]
    if boxes is None and want_boxes:
        boxes = torch.empty(batch, anchors, 4, device=dev, dtype=torch.float32)
    if scores is not None:
        _require_cuda(scores, "scores", torch.float32)
        assert scores.shape == (batch, anchors) and scores.is_contiguous()
        if pass_mask is None:
            pass_mask = torch.empty(batch, (anchors + 31) // 32, device=dev, dtype=torch.int32)
        if act == _cabi.ACT_SIGMOID and scores_act is None:
            scores_act = torch.empty_like(scores)
    entry = lib().ovdet_decode_filter_bf16in if table.dtype == torch.bfloat16 else lib().ovdet_decode_filter
    with torch.cuda.device(dev):
        check(entry(table.ptrs, table.heights, table.widths, table.strides,
                                        table.bstrides, table.n, bins, batch, float(width_scale),
                                        float(height_scale), _ptr(scores), float(conf), act,
                                        _ptr(boxes), _ptr(scores_act), _ptr(pass_mask),
                                        torch.cuda.current_stream(dev).cuda_stream),
              "ovdet_decode_filter")
    return boxes, scores_act, pass_mask


# --------------------------------------------------------------------------------------------
# K4: gather + rescale/clip + sort + top-k + NMS
# --------------------------------------------------------------------------------------------
class NmsResult:
    __slots__ = ("boxes", "scores", "classes", "anchor", "keep", "count", "candidates")

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)


def nms_workspace_bytes(batch: int, anchors: int) -> int:
    return int(lib().ovdet_nms_workspace_bytes(batch, anchors))


def nms_batched(boxes: torch.Tensor, scores: torch.Tensor, classes: Optional[torch.Tensor] = None,
                pass_mask: Optional[torch.Tensor] = None, scale: Optional[torch.Tensor] = None,
                clip_wh: Optional[torch.Tensor] = None, iou_thr: float = 0.45,
                class_aware: bool = False, topk: int = 0, max_det: Optional[int] = None,
                out: Optional[NmsResult] = None, workspace: Optional[torch.Tensor] = None,
                conf: Optional[float] = None) -> NmsResult:
    """detector.py:185-208 + _nms :225-256 + _compute_iou :258-287, for every image.

    boxes ``[B, A, 4]`` fp32, scores ``[B, A]`` fp32, classes ``[B, A]`` int32 (optional),
    pass_mask ``[B, ceil(A/32)]`` int32 bit mask (optional: all anchors), scale ``[B]`` fp32,
    clip_wh ``[B, 2]`` fp32 (w, h).  Result rows are in kept order (score desc).  ``conf`` (instead
    of ``pass_mask``): the kernel thresholds ``scores > conf`` itself (detector.py:184).
    """
    _require_cuda(boxes, "boxes", torch.float32)
    _require_cuda(scores, "scores", torch.float32)
    batch, anchors = scores.shape
    assert boxes.shape == (batch, anchors, 4) and boxes.is_contiguous() and scores.is_contiguous()
    if classes is not None:
        _require_cuda(classes, "classes", torch.int32)
        assert classes.shape == (batch, anchors) and classes.is_contiguous()
    dev = boxes.device
    if max_det is None:
        max_det = max(anchors, 1)
    if out is None:
        out = NmsResult(
            boxes=torch.zeros(batch, max_det, 4, device=dev, dtype=torch.float32),
            scores=torch.zeros(batch, max_det, device=dev, dtype=torch.float32),
            classes=torch.zeros(batch, max_det, device=dev, dtype=torch.int32),
            anchor=torch.full((batch, max_det), -1, device=dev, dtype=torch.int32),
            keep=torch.full((batch, max_det), -1, device=dev, dtype=torch.int32),
            count=torch.zeros(batch, device=dev, dtype=torch.int32),
            candidates=torch.zeros(batch, device=dev, dtype=torch.int32))
    need = nms_workspace_bytes(batch, anchors)
    if workspace is None:
        workspace = torch.empty(max(need, 16), device=dev, dtype=torch.uint8)
    if conf is not None:
        if pass_mask is not None:
            raise ValueError("ovdet: give either pass_mask or conf")
        with torch.cuda.device(dev):
            check(lib().ovdet_nms_batched_conf(boxes.data_ptr(), scores.data_ptr(), _ptr(classes), float(conf),
                                               batch, anchors, _ptr(scale), _ptr(clip_wh), float(iou_thr),
                                               int(class_aware), int(topk), int(max_det),
                                               out.boxes.data_ptr(), out.scores.data_ptr(),
                                               out.classes.data_ptr(), out.anchor.data_ptr(),
                                               out.keep.data_ptr(), out.count.data_ptr(),
                                               out.candidates.data_ptr(), workspace.data_ptr(),
                                               workspace.numel(), _stream(boxes)), "ovdet_nms_batched_conf")
        return out
    with torch.cuda.device(dev):
        check(lib().ovdet_nms_batched(boxes.data_ptr(), scores.data_ptr(), _ptr(classes),
                                      _ptr(pass_mask), batch, anchors, _ptr(scale), _ptr(clip_wh),
                                      float(iou_thr), int(class_aware), int(topk), int(max_det),
                                      out.boxes.data_ptr(), out.scores.data_ptr(),
                                      out.classes.data_ptr(), out.anchor.data_ptr(),
                                      out.keep.data_ptr(), out.count.data_ptr(),
                                      out.candidates.data_ptr(), workspace.data_ptr(),
                                      workspace.numel(), _stream(boxes)), "ovdet_nms_batched")
    return out


# --------------------------------------------------------------------------------------------
# P1 / P2: letterbox pre-processing and int-truncated box records ("next" rows)
# --------------------------------------------------------------------------------------------
def letterbox_geometry(orig_h: int, orig_w: int, image_size: Tuple[int, int]) -> Tuple[float, int, int]:
    """inference/detector.py:139-142 (python-float arithmetic, as the reference does it)."""
    input_h, input_w = image_size
    scale_factor = min(input_h / orig_h, input_w / orig_w)
    return scale_factor, int(orig_h * scale_factor), int(orig_w * scale_factor)


def letterbox(images: Sequence[torch.Tensor], image_size: Tuple[int, int] = (640, 640),
              out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, List[float]]:
    """inference/detector.py:139-156 for a list of uint8 ``[H, W, 3]`` RGB CUDA tensors (any
    sizes): bilinear resize exactly as cv2.resize does it, top-left paste on a zero canvas,
    /255, HWC -> CHW.  Returns the ``[N, 3, H, W]`` fp32 batch and the per-image scale factors."""
    n = len(images)
    out_h, out_w = image_size
    dev = images[0].device
    keep, scales, rhs, rws = [], [], [], []
    for im in images:
        _require_cuda(im, "image", torch.uint8)
        if im.dim() != 3 or im.shape[2] != 3:
            raise ValueError("ovdet: images must be uint8 [H, W, 3]")
        if im.stride(2) != 1 or im.stride(1) != 3:
            im = im.contiguous()
        keep.append(im)
        s, rh, rw = letterbox_geometry(im.shape[0], im.shape[1], image_size)
        if rh < 1 or rw < 1:
            raise ValueError("ovdet: image collapses to zero size on the canvas")
        scales.append(s)
        rhs.append(rh)
        rws.append(rw)
    if out is None:
        out = torch.empty(n, 3, out_h, out_w, device=dev, dtype=torch.float32)
    assert out.shape == (n, 3, out_h, out_w) and out.is_contiguous() and out.dtype == torch.float32
    ptrs = (ctypes.c_void_p * n)(*[im.data_ptr() for im in keep])
    hs = (ctypes.c_int32 * n)(*[im.shape[0] for im in keep])
    ws = (ctypes.c_int32 * n)(*[im.shape[1] for im in keep])
    st = (ctypes.c_int64 * n)(*[im.stride(0) for im in keep])
    with torch.cuda.device(dev):
        check(lib().ovdet_letterbox_u8(ptrs, hs, ws, st, (ctypes.c_int32 * n)(*rhs),
                                       (ctypes.c_int32 * n)(*rws), n, out_h, out_w, out.data_ptr(),
                                       torch.cuda.current_stream(dev).cuda_stream), "ovdet_letterbox_u8")
    return out, scales


def pack_boxes_i32(boxes: torch.Tensor, count: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """detector.py:216 ``boxes[i].astype(int)`` for the kept rows of an ``NmsResult``:
    ``boxes [B, max_det, 4]`` fp32 -> int32 (toward zero); rows past ``count[b]`` are zero."""
    _require_cuda(boxes, "boxes", torch.float32)
    _require_cuda(count, "count", torch.int32)
    batch, max_det, _ = boxes.shape
    assert boxes.is_contiguous() and count.shape == (batch,)
    if out is None:
        out = torch.empty(batch, max_det, 4, device=boxes.device, dtype=torch.int32)
    with torch.cuda.device(boxes.device):
        check(lib().ovdet_pack_boxes_i32(boxes.data_ptr(), count.data_ptr(), batch, max_det,
                                         out.data_ptr(), _stream(boxes)), "ovdet_pack_boxes_i32")
    return out


# --------------------------------------------------------------------------------------------
# N1: max-sigmoid text attention of the neck ("next" row)
# --------------------------------------------------------------------------------------------
def cast_text(text: torch.Tensor, split3: bool = False) -> torch.Tensor:
    """Raw (un-normalised) bf16 operand of projected text ``[C, c]`` or ``[B, C, c]``:
    ``[Bt, C, kop]`` with the columns zero padded to a multiple of 64 (x3 for ``split3``)."""
    _require_cuda(text, "text", torch.float32)
    if text.dim() == 2:
        text = text.unsqueeze(0)
    elif shared_text(text):
        text = text[:1]
    if text.stride(2) != 1:
        text = text.contiguous()
    bt, classes, dim = text.shape
    kop = (dim + 63) // 64 * 64 * (3 if split3 else 1)
    operand = torch.empty(bt, classes, kop, device=text.device, dtype=torch.bfloat16)
    with torch.cuda.device(text.device):
        check(lib().ovdet_cast_text(text.data_ptr(), bt, classes, dim, text.stride(0), text.stride(1),
                                    operand.data_ptr(), kop, int(split3), _stream(text)), "ovdet_cast_text")
    return operand


def max_sigmoid_attention(y: torch.Tensor, projected_text: torch.Tensor, precise: bool = True,
                          out: Optional[torch.Tensor] = None, return_scores: bool = False):
    """model/repvl_pan.py:80-95: ``y [B, c, H, W]`` fp32 (NCHW), ``projected_text [B, C, c]`` (or a
    shared ``[C, c]``) -> ``y * sigmoid(max_C(y^T t'))``, same shape and layout as ``y``."""
    _require_cuda(y, "y", torch.float32)
    b, c, h, w = y.shape
    if y.stride(3) != 1 or y.stride(2) != w:
        y = y.contiguous()
    if (h * w) % 4 or y.stride(0) % 4 or y.stride(1) % 4 or y.data_ptr() % 16:
        raise ValueError("ovdet: max_sigmoid_attention needs H*W and the strides to be multiples of 4")
    text_op = cast_text(projected_text, split3=precise)
    bt = text_op.shape[0]
    assert bt in (1, b) and projected_text.shape[-1] == c
    if out is None:
        out = torch.empty_like(y, memory_format=torch.contiguous_format)
    assert out.shape == y.shape and out.stride(3) == 1 and out.stride(2) == w
    row_max = torch.empty(b, h * w, device=y.device, dtype=torch.float32)
    with torch.cuda.device(y.device):
        check(lib().ovdet_max_sigmoid_attention(y.data_ptr(), b, c, h * w, y.stride(0), y.stride(1),
                                                text_op.data_ptr(), text_op.shape[1],
                                                int(bt == b and b > 1), int(precise), row_max.data_ptr(),
                                                out.data_ptr(), out.stride(0), out.stride(1), _stream(y)),
              "ovdet_max_sigmoid_attention")
    return (out, row_max) if return_scores else out


# --------------------------------------------------------------------------------------------
# f-2: the head's 1x1 projection folded into the similarity ("next" row)
# --------------------------------------------------------------------------------------------
def project_vocabulary(text: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """Fold ``nn.Conv2d(hidden, embed, 1)`` (text_contrastive.py:67) into the vocabulary.

    With ``E = W x + b`` (the conv) and unit-norm text rows ``t``: ``<E, t> = <x, W^T t> + <b, t>``
    and ``||E||^2 = x'^T G' x'`` for the augmented ``x' = [x, 1]`` and
    ``G' = [[W^T W, W^T b], [b^T W, b^T b]]``.  So the kernel multiplies the HIDDEN features
    (K = hidden + 1 instead of embed) against one operand holding the projected classes
    ``[W^T t_c | <b, t_c>]`` followed by the rows of ``G'``; the 512-wide embedding is never formed.

    text ``[C, D]`` (or ``[B, C, D]`` per image) fp32, weight ``[D, K]`` or ``[D, K, 1, 1]``, bias
    ``[D]`` or None.  Returns the bf16 operand ``[Bt, Cpad + kop, kop]`` with
    ``kop = ceil(K / 64) * 64 + 16`` (class rows padded to a multiple of 128, then ``kop`` rows of
    ``G'``; the constant 1 of ``x'`` sits at column ``ceil(K / 64) * 64``).  Built once per
    (vocabulary, weights) with plain fp32 torch GEMMs (a one-off ``[C, D] x [D, K]`` product)."""
    _require_cuda(text, "text", torch.float32)
    w = weight.reshape(weight.shape[0], -1).to(torch.float32)
    d, k = w.shape
    b = bias.to(torch.float32) if bias is not None else torch.zeros(d, device=w.device)
    if text.dim() == 2:
        text = text.unsqueeze(0)
    elif shared_text(text):
        text = text[:1]
    bt, classes, dt = text.shape
    assert dt == d
    kpad = (k + 63) // 64 * 64
    kop = kpad + 16
    cpad = (classes + 127) // 128 * 128
    that = torch.nn.functional.normalize(text, p=2, dim=-1)              # text_contrastive.py:138
    v = torch.zeros(bt, cpad + kop, kop, device=text.device, dtype=torch.float32)
    v[:, :classes, :k] = that @ w                                         # W^T t_c
    v[:, :classes, kpad] = that @ b                                       # <b, t_c>
    g = v[:, cpad:, :]
    g[:, :k, :k] = w.t() @ w
    g[:, :k, kpad] = w.t() @ b
    g[:, kpad, :k] = w.t() @ b
    g[:, kpad, kpad] = b @ b
    return v.to(torch.bfloat16).contiguous()


def similarity_projected(hidden: Sequence[torch.Tensor], level_ops: Sequence[torch.Tensor], classes: int,
                         alpha: float = 1.0, beta: float = 0.0, row_max: Optional[torch.Tensor] = None,
                         row_arg: Optional[torch.Tensor] = None, inv_norm: Optional[torch.Tensor] = None,
                         want_arg: bool = True):
    """text_contrastive.py:112 (the 1x1 projection) + :134-147 + yolo_clip.py:198-206 for all levels in
    one launch, from the HIDDEN features ``hidden[l] [B, K, H, W]`` and the per-level operands of
    ``project_vocabulary``.  Returns ``(row_max [B, A], row_arg [B, A] int32 or None)``."""
    first = hidden[0]
    _require_cuda(first, "hidden", torch.float32)
    if not fused_supported_strides(hidden):
        raise ValueError("ovdet: hidden feature strides not addressable by TMA (H*W and strides must be multiples of 4)")
    batch, k = first.shape[0], first.shape[1]
    n = len(hidden)
    assert len(level_ops) == n
    anchors = sum(h.shape[2] * h.shape[3] for h in hidden)
    kop = (k + 63) // 64 * 64 + 16
    cpad = (classes + 127) // 128 * 128
    bt = level_ops[0].shape[0]
    for op in level_ops:
        _require_cuda(op, "level_op", torch.bfloat16)
        assert op.shape == (bt, cpad + kop, kop) and op.is_contiguous()
    dev = first.device
    if row_max is None:
        row_max = torch.empty(batch, anchors, device=dev, dtype=torch.float32)
    if want_arg and row_arg is None:
        row_arg = torch.empty(batch, anchors, device=dev, dtype=torch.int32)
    ptrs = (ctypes.c_void_p * n)(*[h.data_ptr() for h in hidden])
    hw = (ctypes.c_int64 * n)(*[h.shape[2] * h.shape[3] for h in hidden])
    sb = (ctypes.c_int64 * n)(*[h.stride(0) for h in hidden])
    sd = (ctypes.c_int64 * n)(*[h.stride(1) for h in hidden])
    ops_ptr = (ctypes.c_void_p * n)(*[op.data_ptr() for op in level_ops])
    with torch.cuda.device(dev):
        check(lib().ovdet_similarity_projected(ptrs, hw, sb, sd, n, batch, k, ops_ptr, classes,
                                               int(bt == batch and batch > 1), float(alpha), float(beta),
                                               _ptr(row_max), _ptr(row_arg), _ptr(inv_norm), _stream(first)),
              "ovdet_similarity_projected")
    return row_max, row_arg


def fused_supported_strides(levels: Sequence[torch.Tensor]) -> bool:
    """TMA addressability of per-level fp32 ``[B, K, H, W]`` tensors (any K)."""
    if len(levels) > 4:
        return False
    for e in levels:
        b, d, h, w = e.shape
        if e.dtype != torch.float32 or e.stride(3) != 1 or e.stride(2) != w:
            return False
        if e.stride(1) % 4 or e.stride(0) % 4 or e.data_ptr() % 16 or e.stride(1) < h * w:
            return False
    return True
