"""Micro-benchmarks of the kernels outside the K1..K4 step (CUDA events, inputs > L2 where the
row streams): P1 letterbox.  Prints one JSON line per row."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ovdet import ops

dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}


def timeit(fn, iters=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


for (h, w), n in (((1080, 1920), 64), ((480, 640), 64), ((1280, 1280), 64), ((480, 640), 1)):
    imgs = [torch.randint(0, 256, (h, w, 3), device=dev, dtype=torch.uint8) for _ in range(n)]
    out = torch.empty(n, 3, 640, 640, device=dev)
    ms = timeit(lambda: ops.letterbox(imgs, (640, 640), out=out))
    _, rh, rw = ops.letterbox_geometry(h, w, (640, 640))
    # algorithmic bytes: canvas written once (fp32 CHW) + every source byte the taps touch, once
    taps = min(h, 2 * rh) * min(w, 2 * rw) * 3
    bytes_ = n * (3 * 640 * 640 * 4 + taps)
    print(json.dumps({"row": "P1 letterbox", "images": n, "source": [h, w], "ms": ms,
                      "images_per_s": n / ms * 1e3, "GBps": bytes_ / ms / 1e6,
                      "hbm_frac": bytes_ / ms / 1e6 / peaks["hbm_gbs"]}))
