"""Time the similarity GEMM alone (CUDA events, L2-exceeding operands) for quick kernel iteration."""
import argparse, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ovdet import ops

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--anchors", type=int, default=8400)
ap.add_argument("--classes", type=int, default=1203)
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--logits", default="none")
ap.add_argument("--split", action="store_true")
ap.add_argument("--fused", action="store_true")
ap.add_argument("--projected", action="store_true", help="f-2: hidden features (K = 256) in, 1x1 projection folded")
a = ap.parse_args()
dev = torch.device("cuda:0")
D = 512
kop = D * (2 if a.split else 1)
rop = torch.randn(a.batch, a.anchors, kop, device=dev).to(torch.bfloat16)
top = torch.nn.functional.normalize(torch.randn(1, a.classes, kop, device=dev), dim=-1).to(torch.bfloat16)
inv = torch.rand(a.batch, a.anchors, device=dev) + 0.5
ld = {"none": None, "bf16": torch.bfloat16, "fp32": torch.float32}[a.logits]
logits = torch.empty(a.batch, a.anchors, a.classes, device=dev, dtype=ld) if ld else None
rmax = torch.empty(a.batch, a.anchors, device=dev)
rarg = torch.empty(a.batch, a.anchors, device=dev, dtype=torch.int32)
embs = None
if a.fused:
    assert a.anchors == 8400
    embs = [torch.randn(a.batch, D, s, s, device=dev) for s in (80, 40, 20)]
    top1 = top[:, :, :D].contiguous()
if a.projected:
    hid = [torch.randn(a.batch, 256, s, s, device=dev) for s in (80, 40, 20)]
    text = torch.randn(a.classes, D, device=dev)
    lops = [ops.project_vocabulary(text, torch.randn(D, 256, device=dev) * 0.06, torch.randn(D, device=dev) * 0.1)
            for _ in hid]
def run():
    if a.projected:
        ops.similarity_projected(hid, lops, a.classes, row_max=rmax, row_arg=rarg)
        return
    if a.fused:
        ops.similarity_fused(embs, top1, logits_dtype=None, logits=logits, want_max=True, row_max=rmax, row_arg=rarg)
        return
    ops.similarity(rop, top, inv, D, split=a.split, logits_dtype=None, logits=logits, want_max=True, row_max=rmax, row_arg=rarg)
for _ in range(5): run()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(a.iters): run()
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / a.iters
fl = 2.0 * a.batch * a.anchors * a.classes * D * (3 if a.split else 1)
if a.projected:
    mma = 2.0 * a.batch * a.anchors * ((a.classes + 127) // 128 * 128 + 272) * 272
    print(f"projected: {ms:.3f} ms; equivalent D=512 similarity rate {fl/ms/1e9:.1f} TFLOP/s; MMA work issued {mma/ms/1e9:.1f} TFLOP/s")
print(f"gemm B={a.batch} A={a.anchors} C={a.classes} logits={a.logits} split={a.split} fused={a.fused}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s (bf16 MMA flops)")
