"""Stress of the K3 parity test body (tests/test_gpu_kernels.py::test_decode_filter_vs_oracle): many
score sets, every one checked against the oracle; prints what differs, if anything."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from oracle import ref_port
from ovdet import ops, synth

dev = torch.device("cuda:0")
inp = synth.make_inputs(batch=2, image_size=320, num_classes=10, embed_dim=64, seed=3)
grids = [ref_port.create_grid(2, p.shape[2], p.shape[3], s) for p, s in zip(inp.box_preds, inp.strides)]
ref = ref_port.decode_boxes(inp.box_preds, grids)
A = ref.shape[1]

def unpack(mask):
    m = mask.to(torch.int64) & 0xffffffff
    bits = (m.unsqueeze(-1) >> torch.arange(32)) & 1
    return bits.reshape(mask.shape[0], -1)[:, :A].bool()

bad = 0
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 300):
    torch.manual_seed(1000 + it)
    scores = torch.rand(2, A) - 0.3
    scores[0, 5] = float("nan")
    for act in ("none", "sigmoid"):
        boxes, sact, mask = ops.decode_filter([p.to(dev) for p in inp.box_preds], inp.strides,
                                              scores=scores.to(dev), conf=0.25, activation=act)
        db = (boxes.cpu() - ref).abs()
        ok_boxes = bool((db <= 1e-3 + 1e-4 * ref.abs()).all())
        s = torch.sigmoid(scores) if act == "sigmoid" else scores
        want = s > 0.25
        bits = unpack(mask.cpu())
        near = (s - 0.25).abs() < 1e-6 if act == "sigmoid" else torch.zeros_like(want)
        ok_mask = torch.equal(bits[~near], want[~near])
        ok_act = True
        if act == "sigmoid":
            d = (sact.cpu() - s).abs()
            d[torch.isnan(s)] = 0
            ok_act = bool((d <= 1e-7 + 1e-6 * s.abs().nan_to_num()).all())
        if not (ok_boxes and ok_mask and ok_act):
            bad += 1
            print("iter", it, act, "boxes", ok_boxes, float(db.max()), "mask", ok_mask,
                  int((bits[~near] != want[~near]).sum()), "act", ok_act,
                  float(d.max()) if act == "sigmoid" else None, flush=True)
print("failures", bad)
