"""Drop-in for the neck's text-guided CSP layer ("next" row f-1 of SURVEY.md section 8).

``TextGuidedCSPLayer`` <- model/repvl_pan.py:33-101: same constructor arguments, submodule names
(``cv1``/``cv2``/``cv3``/``bottlenecks``/``text_proj``: reference checkpoints load unchanged) and
forward signature.  The convolutions and the text projection stay in PyTorch/cuDNN; the
max-sigmoid attention of :77-95 (permute, matmul against the projected text, max over classes,
sigmoid, scale, permute back) runs in ``libovdet.so`` (``ops.max_sigmoid_attention``).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .heads import ConvBlock


class DarkBottleneck(nn.Module):
    """model/repvl_pan.py:104-116."""

    def __init__(self, in_channels: int, out_channels: int, shortcut: bool = True):
        super().__init__()
        self.cv1 = ConvBlock(in_channels, out_channels // 2, kernel_size=1)
        self.cv2 = ConvBlock(out_channels // 2, out_channels, kernel_size=3)
        self.shortcut = shortcut and in_channels == out_channels

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x + self.cv2(self.cv1(x)) if self.shortcut else self.cv2(self.cv1(x))


class TextGuidedCSPLayer(nn.Module):
    """``precise`` (plain attribute, not a parameter): True = three-pass bf16 product that matches
    the reference's fp32 matmul to ~1e-6 (needs hidden channels <= 128, true for every layer of
    the reference's neck); False = one bf16 pass."""

    def __init__(self, in_channels: int, out_channels: int, text_dim: int, n_bottlenecks: int = 1,
                 precise: bool = True):
        super().__init__()
        c_ = out_channels // 2
        self.cv1 = ConvBlock(in_channels, c_, kernel_size=1)
        self.cv2 = ConvBlock(in_channels, c_, kernel_size=1)
        self.cv3 = ConvBlock(2 * c_, out_channels, kernel_size=1)
        self.bottlenecks = nn.ModuleList([DarkBottleneck(c_, c_, shortcut=True) for _ in range(n_bottlenecks)])
        self.text_proj = nn.Linear(text_dim, c_)
        self.precise = precise

    def forward(self, x: torch.Tensor, text_embeddings: torch.Tensor) -> torch.Tensor:
        y1 = self.cv1(x)
        # a shared vocabulary (the stride-0 expand of model/yolo_clip.py:123) is projected once
        text = text_embeddings[0] if ops.shared_text(text_embeddings) and text_embeddings.dim() == 3 \
            else text_embeddings
        for bottleneck in self.bottlenecks:
            y1_temp = bottleneck(y1)
            projected = self.text_proj(text)                                  # :77
            y1 = ops.max_sigmoid_attention(y1_temp, projected, precise=self.precise)   # :80-95
        y2 = self.cv2(x)
        return self.cv3(torch.cat((y1, y2), dim=1))
