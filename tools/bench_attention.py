"""N1 max-sigmoid attention row alone (the three T-CSP shapes of the `n` neck at 640^2, C = 1203):
CUDA-event time per call.  Used for same-box A/B of kernel switches, e.g.
    for v in 5 13; do OVDET_EPI2=$v python tools/bench_attention.py; done"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ovdet import ops

dev = torch.device("cuda:0")
out_line = {"OVDET_EPI2": os.environ.get("OVDET_EPI2")}
for (c, side), n in (((32, 80), 64), ((64, 40), 64), ((128, 20), 64)):
    y = torch.randn(n, c, side, side, device=dev)
    t = torch.randn(1203, c, device=dev)
    out = torch.empty_like(y)
    for precise in (True, False):
        fn = lambda: ops.max_sigmoid_attention(y, t, precise=precise, out=out)
        for _ in range(5): fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(50): fn()
        e.record(); torch.cuda.synchronize()
        out_line[f"c{c}_{'fp32' if precise else 'bf16'}_ms"] = round(s.elapsed_time(e) / 50, 4)
print(json.dumps(out_line))
