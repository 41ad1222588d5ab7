"""A stand-in for OpenAI ``clip`` so that the reference modules import in the build container.

TEST INFRASTRUCTURE.  ``clip`` is not vendored in the reference, not listed in its
requirements.txt, not installed here and there is no network.  It is
off the hot path: it only produces the *input* text embeddings, which every parity case
supplies directly.  Used by ``oracle/make_golden.py`` only.
"""
import sys
import types

import torch


class _FakeClipModel(torch.nn.Module):
    def __init__(self, dim: int = 512):
        super().__init__()
        self.dim = dim
        self.anchor = torch.nn.Parameter(torch.zeros(1), requires_grad=False)

    def encode_text(self, tokens: torch.Tensor) -> torch.Tensor:
        # deterministic pseudo-embedding per token row
        out = []
        for row in tokens:
            g = torch.Generator().manual_seed(int(row.sum().item()) % (2 ** 31))
            out.append(torch.randn(self.dim, generator=g))
        return torch.stack(out).to(tokens.device)


def install() -> None:
    if "clip" in sys.modules:
        return
    mod = types.ModuleType("clip")

    def load(name, device="cpu", jit=False):
        return _FakeClipModel().to(device), (lambda img: img)

    def tokenize(texts, context_length: int = 77, truncate: bool = False):
        if isinstance(texts, str):
            texts = [texts]
        rows = []
        for t in texts:
            ids = [ord(ch) % 255 + 1 for ch in t][:context_length]
            rows.append(ids + [0] * (context_length - len(ids)))
        return torch.tensor(rows, dtype=torch.long)

    mod.load = load
    mod.tokenize = tokenize
    sys.modules["clip"] = mod
