"""Synthetic inputs for the head + post-processing path.

Random-init weights give degenerate post-processing (exp(E[w]) boxes of +-5e4 px, cosine
scores < 0.25), so parity and throughput runs feed the hot path with tensors of the shapes
the reference's convolutions emit, with controlled distributions:

* ``obj_embeds[l]``  ~ N(0,1) fp32 ``[B, D, S/s, S/s]`` (NCHW, as ``obj_embed_conv`` emits,
  model/heads/text_contrastive.py:64-68,112), with a planted signal on ~2 % of the anchors
  (3x3 clumps of neighbouring cells sharing one class) so that a realistic number of
  candidates pass ``conf`` and NMS has overlapping boxes to resolve;
* ``text`` ~ N(0,1) fp32 ``[C, D]`` (same on every rank);
* ``box_preds[l]`` ~ N(0,1) fp32 ``[B, 4*(reg_max+1), S/s, S/s]`` with one bin per coordinate
  boosted so that boxes are 1..20 strides wide and centred within a cell of their anchor.

Pure torch, device agnostic (generate on CPU for oracle-sized cases, on the GPU for the
benchmark sizes).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import torch


@dataclass
class HeadInputs:
    obj_embeds: List[torch.Tensor]     # per level [B, D, H, W] fp32
    box_preds: List[torch.Tensor]      # per level [B, 4*(reg_max+1), H, W] fp32
    text: torch.Tensor                 # [C, D] fp32
    strides: Sequence[int]
    image_size: int

    @property
    def batch(self) -> int:
        return self.obj_embeds[0].shape[0]

    @property
    def num_anchors(self) -> int:
        return sum(e.shape[2] * e.shape[3] for e in self.obj_embeds)

    def text_batched(self) -> torch.Tensor:
        """[B, C, D] stride-0 view, as the offline vocabulary path builds it
        (model/yolo_clip.py:123)."""
        return self.text.unsqueeze(0).expand(self.batch, -1, -1)


def make_text(num_classes: int, embed_dim: int = 512, device="cpu", seed: int = 4321) -> torch.Tensor:
    g = torch.Generator(device=device).manual_seed(seed)
    return torch.randn(num_classes, embed_dim, generator=g, device=device, dtype=torch.float32)


def make_inputs(batch: int, image_size: int = 640, num_classes: int = 1203, embed_dim: int = 512,
                strides: Sequence[int] = (8, 16, 32), reg_max: int = 16, device="cpu",
                seed: int = 1234, text_seed: int = 4321, plant_frac: float = 0.02,
                plant_gain: float = 1.5, box_boost: float = 12.0) -> HeadInputs:
    g = torch.Generator(device=device).manual_seed(seed)
    text = make_text(num_classes, embed_dim, device, text_seed)
    bins = reg_max + 1
    obj_embeds, box_preds = [], []
    for s in strides:
        h = w = image_size // s
        hw = h * w
        emb = torch.randn(batch, embed_dim, hw, generator=g, device=device, dtype=torch.float32)
        # --- planted clumps: centres on a random subset of cells, 3x3 neighbourhood each
        n_clumps = max(1, int(round(plant_frac * hw / 9.0)))
        cy = torch.randint(1, max(2, h - 1), (batch, n_clumps), generator=g, device=device)
        cx = torch.randint(1, max(2, w - 1), (batch, n_clumps), generator=g, device=device)
        cls = torch.randint(0, num_classes, (batch, n_clumps), generator=g, device=device)
        dy = torch.tensor([-1, -1, -1, 0, 0, 0, 1, 1, 1], device=device)
        dx = torch.tensor([-1, 0, 1, -1, 0, 1, -1, 0, 1], device=device)
        yy = (cy.unsqueeze(-1) + dy).clamp_(0, h - 1)
        xx = (cx.unsqueeze(-1) + dx).clamp_(0, w - 1)
        cell = (yy * w + xx).reshape(batch, -1)                               # [B, n*9]
        ccls = cls.unsqueeze(-1).expand(-1, -1, 9).reshape(batch, -1)
        # deduplicate overlapping clumps: last writer wins through a dense class map
        cmap = torch.full((batch, hw), -1, device=device, dtype=torch.long)
        cmap.scatter_(1, cell, ccls)
        bidx, aidx = torch.nonzero(cmap >= 0, as_tuple=True)
        planted = text[cmap[bidx, aidx]] * plant_gain                          # [n, D]
        jitter = 1.0 + 0.1 * torch.randn(planted.shape[0], 1, generator=g, device=device)
        emb.permute(0, 2, 1)[bidx, aidx] += planted * jitter
        obj_embeds.append(emb.reshape(batch, embed_dim, h, w))
        # --- box logits with one boosted bin per coordinate
        pred = torch.randn(batch, 4, bins, hw, generator=g, device=device, dtype=torch.float32)
        k_xy = torch.randint(0, 2, (batch, 2, 1, hw), generator=g, device=device)
        k_wh = torch.randint(0, 4, (batch, 2, 1, hw), generator=g, device=device)
        kstar = torch.cat([k_xy, k_wh], dim=1)
        pred.scatter_add_(2, kstar, torch.full_like(kstar, box_boost, dtype=torch.float32))
        box_preds.append(pred.reshape(batch, 4 * bins, h, w))
    return HeadInputs(obj_embeds, box_preds, text, tuple(strides), image_size)


@dataclass
class ProjectedInputs:
    hidden: List[torch.Tensor]          # per level [B, K, H, W] fp32: input of the head's last 1x1 conv
    weights: List[torch.Tensor]         # per level [D, K, 1, 1]
    biases: List[torch.Tensor]          # per level [D]
    box_preds: List[torch.Tensor]
    text: torch.Tensor
    strides: Sequence[int]
    image_size: int

    @property
    def batch(self) -> int:
        return self.hidden[0].shape[0]

    def text_batched(self) -> torch.Tensor:
        return self.text.unsqueeze(0).expand(self.batch, -1, -1)

    def projections(self):
        return list(zip(self.weights, self.biases))


def make_projected_inputs(batch: int, hidden_dim: int = 256, bias_scale: float = 0.05, **kw) -> ProjectedInputs:
    """Inputs for the projected path ("next" row f-2): per-level 1x1 convolutions ``(W, b)`` with the
    reference's initialisation scale (text_contrastive.py:90-99, kaiming fan_out) and HIDDEN
    features ``x = W^+ (e - b)`` for the embeddings ``e`` of ``make_inputs``, so that ``W x + b`` is
    the orthogonal projection of ``e`` onto the range of ``W``: the planted anchors keep cosine
    ~0.6 to their class (> conf), the background stays ~0, and the candidate statistics match
    the un-projected workload."""
    base = make_inputs(batch=batch, **kw)
    dev = base.text.device
    d = base.text.shape[1]
    g = torch.Generator(device=dev).manual_seed(kw.get("seed", 1234) + 99)
    hidden, weights, biases = [], [], []
    for e in base.obj_embeds:
        w = torch.randn(d, hidden_dim, generator=g, device=dev) * (2.0 / d) ** 0.5
        b = torch.randn(d, generator=g, device=dev) * bias_scale
        pinv = torch.linalg.pinv(w)                                           # [K, D]
        bsz, _, h, wd = e.shape
        x = torch.matmul(pinv, e.reshape(bsz, d, h * wd) - b.view(1, d, 1))   # [B, K, HW]
        hidden.append(x.reshape(bsz, hidden_dim, h, wd).contiguous())
        weights.append(w.reshape(d, hidden_dim, 1, 1))
        biases.append(b)
    return ProjectedInputs(hidden, weights, biases, base.box_preds, base.text, base.strides, base.image_size)
