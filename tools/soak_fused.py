"""Soak of the fused similarity kernel's producer / converter / MMA / epilogue protocol: random batch sizes,
class counts and pyramid shapes (aligned and odd levels, one to many N tiles, with and without the
eight-warp converter), every launch compared with the independent two-kernel path (K1 -> K2) on the same
inputs.  A protocol race (a block published before its slot was released, a stale poll ...) shows up as a wrong
score, not as a hang, because every wait is bounded.

    python tools/soak_fused.py [iterations] [seed]

Prints one JSON line.  A test tool; nothing of the product imports it.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ovdet import ops


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(seed)
    worst = {"bf16": 0.0, "fp16": 0.0}
    arg_mismatch = 0
    rows_total = 0
    for it in range(iters):
        batch = int(torch.randint(1, 7, (1,), generator=g))
        classes = int(torch.randint(1, 700, (1,), generator=g))
        n_levels = int(torch.randint(1, 4, (1,), generator=g))
        shapes = []
        for _ in range(n_levels):
            h = int(torch.randint(1, 41, (1,), generator=g))
            w = int(torch.randint(1, 41, (1,), generator=g))
            shapes.append((h, w))
        per_image = bool(torch.randint(0, 2, (1,), generator=g))
        embs = [torch.randn(batch, 512, h, w, device=dev) * (0.3 + l) for l, (h, w) in enumerate(shapes)]
        text = torch.randn(batch, classes, 512, device=dev) if per_image else torch.randn(classes, 512, device=dev)
        levels = ops.tma_addressable(embs)
        # reference: K1 -> K2 with the same bf16 operands
        rop, inv = ops.l2norm_regions(embs)
        top = ops.l2norm_text(text)
        ref_logits, ref_max, ref_arg = ops.similarity(rop, top, inv, 512, 1.0, 0.0, logits_dtype=torch.float32,
                                                      want_max=True)
        _, m, a = ops.similarity_fused(levels, top, 1.0, 0.0, logits_dtype=None, want_max=True)
        torch.cuda.synchronize()
        err = (m - ref_max).abs().max().item()
        worst["bf16"] = max(worst["bf16"], err)
        assert err <= 2e-5, (it, batch, classes, shapes, per_image, err)
        bad = (a != ref_arg)
        # an argmax may differ only between (near-)equal logits
        if bad.any():
            idx = bad.nonzero()
            gap = (ref_logits[idx[:, 0], idx[:, 1], a[bad].long()] - ref_max[bad]).abs().max().item()
            assert gap <= 4e-5, (it, gap)
            arg_mismatch += int(bad.sum())
        rows_total += m.numel()
        # fp16 tier against the fp32 arithmetic of torch on the same inputs
        if ops.fused_fp16_supported(levels):
            top16 = ops.l2norm_text(text, split="fp16")
            _, m16, _ = ops.similarity_fused(levels, top16, 1.0, 0.0, logits_dtype=None, want_max=True)
            flat = torch.cat([e.flatten(2).transpose(1, 2) for e in embs], dim=1)
            tn = torch.nn.functional.normalize(text if per_image else text.unsqueeze(0).expand(batch, -1, -1), dim=-1)
            an = torch.nn.functional.normalize(flat, dim=-1)
            exact = torch.bmm(an.double(), tn.double().transpose(1, 2)).max(dim=-1).values.float()
            torch.cuda.synchronize()
            err16 = (m16 - exact).abs().max().item()
            worst["fp16"] = max(worst["fp16"], err16)
            assert err16 <= 1e-4, (it, batch, classes, shapes, per_image, err16)
    print(json.dumps({"iterations": iters, "seed": seed, "rows_checked": rows_total,
                      "max_abs_diff_vs_two_kernel_path_bf16": worst["bf16"],
                      "max_abs_err_fp16_tier_vs_fp64": worst["fp16"], "argmax_near_ties": arg_mismatch, "ok": True}))


if __name__ == "__main__":
    main()
