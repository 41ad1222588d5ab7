"""Drop-in head modules: the reference's constructor arguments, method signatures, return
shapes / strides / dtypes and state-dict keys, with the arithmetic of the hot path running in
``libovdet.so``.

* ``TextContrastiveHead``  <- model/heads/text_contrastive.py:32-222
* ``BoxHead``              <- model/heads/box_head.py:31-218
* ``head_tail``            <- the tail of YOLOCLIP.forward, model/yolo_clip.py:173-223, from the conv outputs
* ``forward_tail``         <- the same tail from the neck's outputs, all six dict keys
* ``patch_yolo_clip``      <- rebinds an existing YOLOCLIP.forward to ``forward_tail``

The convolution stacks stay in PyTorch/cuDNN (out of
scope); no parameter or buffer is added, so reference checkpoints load unchanged.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops

PRECISIONS = ("fp32", "bf16", "fp16", "auto")


class ConvBlock(nn.Module):
    """conv -> BN -> SiLU; submodule names ``conv`` / ``bn`` / ``act`` match the reference's
    state-dict keys (text_contrastive.py:11-29)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 3, stride: int = 1,
                 padding: Optional[int] = None):
        super().__init__()
        pad = kernel_size // 2 if padding is None else padding
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, pad, bias=False)
        self.bn = nn.BatchNorm2d(out_channels)
        self.act = nn.SiLU(inplace=True)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.act(self.bn(self.conv(x)))


def _branch(cin: int, hidden: int, cout: int) -> nn.Sequential:
    return nn.Sequential(ConvBlock(cin, hidden, 3), ConvBlock(hidden, hidden, 3),
                         nn.Conv2d(hidden, cout, kernel_size=1))


def _init_like_reference(module: nn.Module) -> None:
    # text_contrastive.py:90-99 / box_head.py:71-81
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)


class TextContrastiveHead(nn.Module):
    """Region/text contrastive head.  ``precision`` selects the tensor-core recipe of
    ``compute_similarity``: ``"fp32"`` (default; three bf16 passes over hi/lo operand halves,
    |dlogit| ~ 1e-5), ``"fp16"`` (one pass with fp16 operands, |dlogit| <~ 1e-4, embed_dim 512) or
    ``"bf16"`` (one pass, |dlogit| <~ 8e-3).  It is a plain attribute, not a parameter or buffer."""

    def __init__(self, in_channels: int, embed_dim: int = 512, hidden_dim: int = 256,
                 reg_max: int = 16, cls_alpha: float = 1.0, cls_beta: float = 0.0,
                 width_scale: float = 1.0, height_scale: float = 1.0, precision: str = "fp32"):
        super().__init__()
        self.obj_embed_conv = _branch(in_channels, hidden_dim, embed_dim)
        self.box_conv = _branch(in_channels, hidden_dim, 4 * (reg_max + 1))
        self.embed_dim = embed_dim
        self.reg_max = reg_max
        self.cls_alpha = cls_alpha
        self.cls_beta = cls_beta
        self.width_scale = width_scale
        self.height_scale = height_scale
        assert precision in PRECISIONS
        self.precision = precision
        _init_like_reference(self)

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """text_contrastive.py:101-117: ``(obj_embed [B,D,H,W], box_preds [B,4R,H,W])``."""
        return self.obj_embed_conv(x), self.box_conv(x)

    def forward_hidden(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """``(hidden [B,hidden_dim,H,W], box_preds)``: ``forward`` without the last layer of
        ``obj_embed_conv`` - the 1x1 projection that ``ops.similarity_projected`` folds into the
        similarity (``projection()`` hands out its weight and bias)."""
        return self.obj_embed_conv[1](self.obj_embed_conv[0](x)), self.box_conv(x)

    def projection(self) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """Weight ``[embed_dim, hidden_dim, 1, 1]`` and bias of text_contrastive.py:67."""
        conv = self.obj_embed_conv[2]
        return conv.weight, conv.bias

    def compute_similarity(self, obj_embed: torch.Tensor, text_embed: torch.Tensor) -> torch.Tensor:
        """text_contrastive.py:119-153: L2-normalise both sides, contract over D, apply
        ``cls_alpha * s + cls_beta``.  Returns logical ``[B,C,H,W]`` whose memory is ``[B,HW,C]``
        (the same transposed view, with the same strides, the reference returns)."""
        b, d, h, w = obj_embed.shape
        c = text_embed.shape[-2]
        level = ops.tma_addressable([obj_embed.float()])
        if self.precision == "fp16" and ops.fused_fp16_supported(level):
            text_op = ops.l2norm_text(text_embed.float(), split="fp16")
            logits, _, _ = ops.similarity_fused(level, text_op, self.cls_alpha, self.cls_beta,
                                                logits_dtype=torch.float32, want_max=False)
            return logits.transpose(1, 2).reshape(b, c, h, w)
        split = self.precision != "bf16"        # "fp16" outside its shapes: the three-pass recipe
        # one fused launch (L2 norm + similarity, fp32 NCHW read in place) when the shape allows:
        # bf16 always, fp32-accurate for a single class tile; else K1 -> K2
        if d % 64 == 0 and d <= 512 and (ops.fused_fp32_supported(level, c) if split else ops.fused_supported(level)):
            text_op = ops.text_operand_fp32(text_embed.float()) if split else ops.l2norm_text(text_embed.float())
            logits, _, _ = ops.similarity_fused(level, text_op, self.cls_alpha, self.cls_beta,
                                                logits_dtype=torch.float32, want_max=False, fp32=split)
            return logits.transpose(1, 2).reshape(b, c, h, w)
        regions_op, inv_norm = ops.l2norm_regions(level, split=split)
        text_op = ops.l2norm_text(text_embed.float(), split=split)
        logits, _, _ = ops.similarity(regions_op, text_op, inv_norm, d, self.cls_alpha,
                                      self.cls_beta, split=split, logits_dtype=torch.float32)
        return logits.transpose(1, 2).reshape(b, c, h, w)

    def decode_boxes(self, box_preds: torch.Tensor, grid_sizes: List[Tuple[int, int]],
                     strides: List[int]) -> torch.Tensor:
        """text_contrastive.py:155-222 (never called by the reference model): the same
        centre/size decode with ``width_scale`` / ``height_scale``; every level is cropped out
        of the one ``box_preds`` tensor at ``[:height, :width]``."""
        levels = [box_preds[:, :, :gh, :gw] for gh, gw in grid_sizes]
        boxes, _, _ = ops.decode_filter(levels, strides, width_scale=self.width_scale,
                                        height_scale=self.height_scale)
        return boxes


class BoxHead(nn.Module):
    def __init__(self, in_channels: List[int], hidden_dim: int = 256, reg_max: int = 16,
                 strides: List[int] = [8, 16, 32]):
        super().__init__()
        self.box_convs = nn.ModuleList(_branch(c, hidden_dim, 4 * (reg_max + 1)) for c in in_channels)
        self.reg_max = reg_max
        self.strides = strides
        _init_like_reference(self)

    def forward(self, features: List[torch.Tensor]) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
        """box_head.py:83-113: per-level predictions and the int64 ``(x, y, stride)`` grids.
        The grids are returned for API compatibility only; ``decode_boxes`` derives the cell
        coordinates from the anchor index."""
        assert len(features) == len(self.box_convs), \
            f"Expected {len(self.box_convs)} feature maps, got {len(features)}"
        preds, grids = [], []
        for feat, conv, stride in zip(features, self.box_convs, self.strides):
            preds.append(conv(feat))
            grids.append(self._create_grid(feat.shape[0], feat.shape[2], feat.shape[3], stride, feat.device))
        return preds, grids

    def _create_grid(self, batch_size: int, height: int, width: int, stride: int,
                     device: torch.device) -> torch.Tensor:
        """box_head.py:115-148: int64 ``[B,H,W,3]``, last axis (column, row, stride)."""
        ys = torch.arange(height, device=device).view(height, 1).expand(height, width)
        xs = torch.arange(width, device=device).view(1, width).expand(height, width)
        cell = torch.stack([xs, ys, torch.full_like(xs, stride)], dim=-1)
        return cell.unsqueeze(0).expand(batch_size, -1, -1, -1)

    def decode_boxes(self, box_preds: List[torch.Tensor], grids: List[torch.Tensor] = None) -> torch.Tensor:
        """box_head.py:150-218 -> ``[B, sum HW, 4]`` xyxy."""
        boxes, _, _ = ops.decode_filter(box_preds, self.strides)
        return boxes


def _similarity_levels(obj_embeds: Sequence[torch.Tensor], text_embeddings: torch.Tensor,
                       cls_alpha: float, cls_beta: float, precision: str, want_logits: bool):
    """Similarity + class max / argmax of all levels: ``(logits or None, scores, class_ids int32,
    path)``.  ONE fused launch straight from the fp32 NCHW conv outputs whenever the kernel takes
    the shape (always for bf16 with TMA-addressable levels; fp32-accurate for any class count at
    dim <= 512); K1 -> K2 otherwise."""
    classes = text_embeddings.shape[-2]
    levels = [e if e.dtype in (torch.float32, torch.bfloat16) else e.float() for e in obj_embeds]
    levels = ops.tma_addressable(levels)          # odd H*W (13x13, 19x19 ...): rows re-pitched to 16 bytes
    if precision == "auto":             # inside the fp32 bar at the best speed the shape allows
        # (the same rule as pipeline.resolve_precision: the fp16 tier - one tensor-core pass, |dlogit| <= 1e-4 -
        # wherever its kernel takes the shape; the three-pass recipe otherwise and for fp32 logits)
        precision = "fp16" if (ops.fused_fp16_supported(levels) and not want_logits) else "fp32"
    if precision == "fp16" and not ops.fused_fp16_supported(levels):
        precision = "fp32"              # shapes outside the fp16 tier: the (more accurate) three-pass recipe
    split = precision == "fp32"
    fused = ops.fused_fp32_supported(levels, classes) if split else ops.fused_supported(levels)
    logits_dtype = torch.float32 if want_logits else None
    if fused:
        text_op = (ops.text_operand_fp32(text_embeddings.float()) if split else
                   ops.l2norm_text(text_embeddings.float(), split="fp16" if precision == "fp16" else False))
        logits, scores, class_ids = ops.similarity_fused(levels, text_op, cls_alpha, cls_beta,
                                                         logits_dtype=logits_dtype, want_max=True, fp32=split)
        return logits, scores, class_ids, "fused_fp32" if split else ("fused_fp16" if precision == "fp16" else "fused")
    levels = [e.float() for e in levels]
    regions_op, inv_norm = ops.l2norm_regions(levels, split=split)
    text_op = ops.l2norm_text(text_embeddings.float(), split=split)
    logits, scores, class_ids = ops.similarity(regions_op, text_op, inv_norm, levels[0].shape[1], cls_alpha,
                                               cls_beta, split=split, logits_dtype=logits_dtype, want_max=True)
    return logits, scores, class_ids, "split"


def head_tail(obj_embeds: Sequence[torch.Tensor], text_embeddings: torch.Tensor,
              box_preds: Sequence[torch.Tensor], strides: Sequence[int] = (8, 16, 32),
              cls_alpha: float = 1.0, cls_beta: float = 0.0, precision: str = "auto",
              return_logits: bool = False) -> Dict[str, torch.Tensor]:
    """The tail of ``YOLOCLIP.forward`` (model/yolo_clip.py:173-223) from the convolution outputs
    on: similarity for every level, class max / argmax, level concat, box decode - ONE fused
    similarity launch (L2 norm + tcgen05 GEMM + class max / argmax reading the NCHW conv outputs
    in place) and ONE decode launch.  Returns the reference's dict keys ``boxes`` / ``scores`` /
    ``class_ids`` (int64), plus ``logits [B, A, C]`` when asked."""
    logits, scores, class_ids, _ = _similarity_levels(obj_embeds, text_embeddings, cls_alpha, cls_beta,
                                                      precision, return_logits)
    boxes, _, _ = ops.decode_filter(box_preds, strides)
    out = {"boxes": boxes, "scores": scores, "class_ids": class_ids.long()}
    if return_logits:
        out["logits"] = logits
    return out


class TailOutputs(dict):
    """The forward dict of ``YOLOCLIP.forward`` (model/yolo_clip.py:216-223) with values that
    may be produced on first access.  ``obj_embeddings [B, A, D]`` is a 17 MB-per-image transposed
    copy of the conv outputs that only the trainer reads (train/trainer.py:144-182); inference
    (inference/detector.py:179-181) reads ``boxes`` / ``scores`` / ``class_ids`` only, so the copy
    is made when - and if - the key is looked up.  Behaves like the plain dict otherwise (``keys``,
    ``in``, ``items``, ``**`` unpacking all see six entries)."""

    def __init__(self, eager: Dict, lazy: Dict):
        super().__init__(eager)
        self._lazy = dict(lazy)
        for key in self._lazy:
            super().__setitem__(key, None)

    def _resolve(self, key):
        make = self._lazy.pop(key, None)
        if make is not None:
            super().__setitem__(key, make())

    def __getitem__(self, key):
        self._resolve(key)
        return super().__getitem__(key)

    def get(self, key, default=None):
        return self[key] if key in self else default

    def __iter__(self):            # a python-level __iter__ keeps dict(x) / {**x} off the C fast path
        return super().__iter__()

    def items(self):
        for key in list(self._lazy):
            self._resolve(key)
        return super().items()

    def values(self):
        for key in list(self._lazy):
            self._resolve(key)
        return super().values()

    def pop(self, key, *default):
        self._resolve(key)
        return super().pop(key, *default)

    def copy(self):
        return dict(self.items())


def forward_tail(pan_features: Sequence[torch.Tensor], text_embeddings: torch.Tensor,
                 contrastive_heads: Sequence[nn.Module], box_head: nn.Module,
                 precision: str = "auto") -> Dict[str, torch.Tensor]:
    """Drop-in for model/yolo_clip.py:173-223 - everything ``YOLOCLIP.forward`` does after the
    neck: ``pan_features`` (the neck's per-level maps) and the neck's text embeddings ``[B, C, D]``
    (any strides: the neck emits ``(D, B*D, 1)``, the offline vocabulary a stride-0 expand) go
    through the heads' convolution stacks (cuDNN, unchanged), ONE fused similarity launch for all
    levels (normalise + contraction + class max / argmax) and ONE decode launch.

    ``contrastive_heads`` / ``box_head`` are the model's own modules - the reference's classes or
    the drop-ins of this file; only their convolution stacks and plain attributes are used.
    Returns the reference's six keys with its shapes and dtypes: ``boxes [B, A, 4]`` fp32,
    ``scores [B, A]`` fp32, ``class_ids [B, A]`` int64, ``obj_embeddings [B, A, D]`` fp32,
    ``text_embeddings`` (as handed in), ``box_preds`` (list of ``[B, 4R, H, W]``).

    ``precision``: ``"auto"`` (default) keeps every score within 1e-4 of the reference's fp32 arithmetic
    at the best speed the shape allows - the fp16 tensor-core tier (one pass, bf16 speed) at embed_dim 512
    with TMA-addressable levels, the three-pass recipe otherwise; ``"fp32"`` / ``"fp16"`` / ``"bf16"`` force
    a recipe."""
    assert precision in PRECISIONS
    assert len(pan_features) == len(contrastive_heads)
    # yolo_clip.py:177-186 calls head(feat) and drops the second output (the head's own box
    # branch, dead in the reference): only the embedding branch is evaluated here
    obj_embeds = [head.obj_embed_conv(feat) if hasattr(head, "obj_embed_conv") else head(feat)[0]
                  for feat, head in zip(pan_features, contrastive_heads)]
    # box_head.forward also builds int64 grids that only decode_boxes consumed (box_head.py:115-148)
    if hasattr(box_head, "box_convs"):
        box_preds = [conv(feat) for conv, feat in zip(box_head.box_convs, pan_features)]
    else:
        box_preds = box_head(list(pan_features))[0]
    affine = {(float(h.cls_alpha), float(h.cls_beta)) for h in contrastive_heads}
    if len(affine) == 1:
        (alpha, beta), = affine
        _, scores, class_ids, _ = _similarity_levels(obj_embeds, text_embeddings, alpha, beta, precision, False)
    else:       # per-level affine parameters: one launch per level, concatenated like yolo_clip.py:205-206
        parts = [_similarity_levels([e], text_embeddings, float(h.cls_alpha), float(h.cls_beta), precision, False)
                 for e, h in zip(obj_embeds, contrastive_heads)]
        scores = torch.cat([p[1] for p in parts], dim=1)
        class_ids = torch.cat([p[2] for p in parts], dim=1)
    boxes, _, _ = ops.decode_filter(box_preds, box_head.strides)
    eager = {"boxes": boxes, "scores": scores, "class_ids": class_ids.long()}
    lazy = {"obj_embeddings": lambda: ops.concat_embeddings([e.float() for e in obj_embeds])}
    out = TailOutputs(eager, lazy)
    out["text_embeddings"] = text_embeddings
    out["box_preds"] = box_preds
    return out


def prompt_embeddings(model: nn.Module, batch: int, text_prompts=None, class_names=None) -> torch.Tensor:
    """The ``[B, C, D]`` text embeddings ``YOLOCLIP.forward`` builds before the backbone
    (model/yolo_clip.py:121-165): the offline vocabulary as a stride-0 expand, or the encoder's
    output for a shared prompt list / one prompt list per image (short lists are zero padded to
    the longest, a short outer list repeats its last entry)."""
    if getattr(model, "offline_mode", False):
        if model.offline_vocabulary is None:
            if class_names is None:
                raise ValueError("In offline mode, either offline_vocabulary or class_names must be provided")
            model.offline_vocabulary = model.vocab_builder.build_online_vocabulary(class_names)
        return model.offline_vocabulary.unsqueeze(0).expand(batch, -1, -1)
    if text_prompts is None:
        raise ValueError("In online mode, text_prompts must be provided")
    if len(text_prompts) > 0 and isinstance(text_prompts[0], (list, tuple)):
        if len(text_prompts) == 1 or batch == 1:        # one list serves every image: encode once, stride-0 batch
            return model.text_encoder(list(text_prompts[0])).unsqueeze(0).expand(batch, -1, -1)
        per_image = [model.text_encoder(list(text_prompts[min(i, len(text_prompts) - 1)])) for i in range(batch)]
        rows = max(e.shape[0] for e in per_image)
        padded = [e if e.shape[0] == rows else torch.cat([e, e.new_zeros(rows - e.shape[0], e.shape[1])])
                  for e in per_image]
        return torch.stack(padded)
    return model.text_encoder(text_prompts).unsqueeze(0).expand(batch, -1, -1)


def patch_yolo_clip(model: nn.Module, precision: str = "auto") -> nn.Module:
    """Point an existing ``YOLOCLIP`` instance (model/yolo_clip.py:16-263) at the CUDA tail: its
    ``forward`` keeps the text handling, backbone and neck it has and hands the neck's outputs to
    ``forward_tail``.  No parameter, buffer or submodule changes, so checkpoints load as before::

        model = YOLOCLIP(...); ovdet.heads.patch_yolo_clip(model)
    """
    assert precision in PRECISIONS

    def forward(images, text_prompts=None, class_names=None):
        text = prompt_embeddings(model, images.shape[0], text_prompts, class_names)
        features = model.backbone(images)
        pan_features, updated_text = model.neck(features, text)
        return forward_tail(pan_features, updated_text, model.contrastive_heads, model.box_head, precision)

    model.forward = forward
    model.ovdet_precision = precision
    return model


def head_tail_projected(hidden: Sequence[torch.Tensor], projections: Sequence[Tuple[torch.Tensor, Optional[torch.Tensor]]],
                        text_embeddings: torch.Tensor, box_preds: Sequence[torch.Tensor],
                        strides: Sequence[int] = (8, 16, 32), cls_alpha: float = 1.0,
                        cls_beta: float = 0.0) -> Dict[str, torch.Tensor]:
    """``head_tail`` from the HIDDEN features of every level's ``obj_embed_conv`` (the input of its
    last, 1x1 layer) with that layer folded into the similarity ("next" row f-2): one launch for
    projection + normalise + similarity + class max of all levels, one decode launch.  The
    512-wide ``obj_embeddings`` are never formed, so the dict carries only ``boxes`` / ``scores`` /
    ``class_ids`` - what ``YOLOCLIPDetector.postprocess_detections`` consumes."""
    classes = text_embeddings.shape[-2]
    level_ops = [ops.project_vocabulary(text_embeddings, w, b) for w, b in projections]
    scores, class_ids = ops.similarity_projected(hidden, level_ops, classes, cls_alpha, cls_beta)
    boxes, _, _ = ops.decode_filter(box_preds, strides)
    return {"boxes": boxes, "scores": scores, "class_ids": class_ids.long()}
