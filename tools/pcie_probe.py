"""Pinned host -> device copy bandwidth of this box (the ceiling of bench.py's e2e figure)."""
import json, torch
dev = torch.device("cuda:0")
out = {}
for mb in (64, 624, 2048):
    h = torch.empty(mb * 1024 * 1024, dtype=torch.uint8).pin_memory()
    d = torch.empty_like(h, device=dev)
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5): d.copy_(h, non_blocking=True)
    e.record(); torch.cuda.synchronize()
    out[f"h2d_{mb}MiB_GBps"] = 5 * h.numel() / (s.elapsed_time(e) * 1e-3) / 1e9
print(json.dumps(out))
