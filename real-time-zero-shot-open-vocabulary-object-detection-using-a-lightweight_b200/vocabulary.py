"""Offline vocabulary: the reference's JSON wire format and a device-side operand cache.

Wire format (clip/vocab_builder.py:90-104, read back at :110-130 and stacked in
model/yolo_clip.py:244-263): one JSON object ``{class_name: [D floats]}``; the class order is
the key order of the file.  ``Vocabulary`` keeps the fp32 matrix and hands out the normalised
bf16 tensor-core operands (``ops.l2norm_text``) once per (device, recipe) instead of once per
level per forward as the reference does (model/heads/text_contrastive.py:138).
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Sequence, Tuple

import torch


def load_offline_vocabulary(path: str) -> Tuple[List[str], torch.Tensor]:
    """Class names in file order and the ``[C, D]`` float32 embedding matrix."""
    with open(path, "r") as f:
        table = json.load(f)
    if not isinstance(table, dict) or not table:
        raise ValueError(f"ovdet: {path} is not a non-empty {{class_name: embedding}} object")
    names = list(table.keys())
    dim = len(table[names[0]])
    for n in names:
        if len(table[n]) != dim:
            raise ValueError(f"ovdet: embedding of {n!r} has {len(table[n])} values, expected {dim}")
    return names, torch.tensor([table[n] for n in names], dtype=torch.float32)


def save_offline_vocabulary(path: str, class_names: Sequence[str], embeddings: torch.Tensor) -> None:
    """Write the reference's format: python floats of the float32 values, insertion order."""
    if embeddings.dim() != 2 or embeddings.shape[0] != len(class_names):
        raise ValueError("ovdet: embeddings must be [len(class_names), D]")
    folder = os.path.dirname(path)
    if folder:
        os.makedirs(folder, exist_ok=True)
    rows = embeddings.detach().to("cpu", torch.float32).numpy()
    with open(path, "w") as f:
        json.dump({n: rows[i].tolist() for i, n in enumerate(class_names)}, f)


class Vocabulary:
    """Class names + fp32 embeddings + cached device operands."""

    def __init__(self, class_names: Sequence[str], embeddings: torch.Tensor):
        if embeddings.dim() != 2 or embeddings.shape[0] != len(class_names):
            raise ValueError("ovdet: embeddings must be [len(class_names), D]")
        self.class_names = list(class_names)
        self.embeddings = embeddings.detach().to(torch.float32).contiguous()
        self._operands: Dict[tuple, torch.Tensor] = {}

    @classmethod
    def load(cls, path: str) -> "Vocabulary":
        return cls(*load_offline_vocabulary(path))

    def save(self, path: str) -> None:
        save_offline_vocabulary(path, self.class_names, self.embeddings)

    def __len__(self) -> int:
        return len(self.class_names)

    @property
    def dim(self) -> int:
        return self.embeddings.shape[1]

    def matrix(self, device) -> torch.Tensor:
        """The fp32 ``[C, D]`` matrix on ``device`` (what ``offline_vocabulary`` holds in the
        reference, model/yolo_clip.py:260)."""
        return self.embeddings.to(device)

    def operand(self, device, split: bool = False) -> torch.Tensor:
        """Normalised bf16 operand ``[1, C, D * (2 if split else 1)]`` on ``device``; built by the
        K1 text kernel on first use and cached."""
        from . import ops
        key = (str(torch.device(device)), bool(split))
        if key not in self._operands:
            self._operands[key] = ops.l2norm_text(self.matrix(device), split=split)
        return self._operands[key]
