"""Drop-in head modules: the reference's constructor arguments, method signatures, return
shapes / strides / dtypes and state-dict keys, with the arithmetic of the hot path running in
``libovdet.so``.

* ``TextContrastiveHead``  <- model/heads/text_contrastive.py:32-222
* ``BoxHead``              <- model/heads/box_head.py:31-218
* ``head_tail``            <- the tail of YOLOCLIP.forward, model/yolo_clip.py:173-223

The convolution stacks stay in PyTorch/cuDNN (out of
scope); no parameter or buffer is added, so reference checkpoints load unchanged.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops

PRECISIONS = ("fp32", "bf16")


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
    |dlogit| ~ 1e-5) or ``"bf16"`` (one pass, |dlogit| <~ 8e-3).  It is a plain attribute, not
    a parameter or buffer."""

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
        split = self.precision == "fp32"
        level = [obj_embed.float()]
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


def head_tail(obj_embeds: Sequence[torch.Tensor], text_embeddings: torch.Tensor,
              box_preds: Sequence[torch.Tensor], strides: Sequence[int] = (8, 16, 32),
              cls_alpha: float = 1.0, cls_beta: float = 0.0, precision: str = "bf16",
              return_logits: bool = False) -> Dict[str, torch.Tensor]:
    """The tail of ``YOLOCLIP.forward`` (model/yolo_clip.py:173-223) from the convolution outputs
    on: similarity for every level, class max / argmax, level concat, box decode.  All levels
    go through ONE normalise launch per level, ONE GEMM with the max/argmax fused in the
    epilogue and ONE decode launch.  Returns the reference's dict keys ``boxes`` / ``scores`` /
    ``class_ids`` (int64), plus ``logits [B, A, C]`` when asked."""
    split = precision == "fp32"
    dim = obj_embeds[0].shape[1]
    regions_op, inv_norm = ops.l2norm_regions(obj_embeds, split=split)
    text_op = ops.l2norm_text(text_embeddings, split=split)
    logits, scores, class_ids = ops.similarity(
        regions_op, text_op, inv_norm, dim, cls_alpha, cls_beta, split=split,
        logits_dtype=torch.float32 if return_logits else None, want_max=True)
    boxes, _, _ = ops.decode_filter(box_preds, strides)
    out = {"boxes": boxes, "scores": scores, "class_ids": class_ids.long()}
    if return_logits:
        out["logits"] = logits
    return out


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
