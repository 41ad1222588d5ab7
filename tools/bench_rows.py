"""Micro-benchmarks of the kernels outside the K1..K4 step (CUDA events, inputs > L2 where the
row streams): P1 letterbox.  Prints one JSON line per row."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ovdet import ops

dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}


def timeit(fn, iters=50, warm=5, graph=False):
    """CUDA-event time per call; graph=True replays the call from a CUDA graph so that the figure
    is the device time even when the Python wrapper (ctypes tables per image) is slower."""
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    if graph:
        g = torch.cuda.CUDAGraph()
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            with torch.cuda.graph(g, stream=st):
                fn()
        torch.cuda.synchronize()
        fn = g.replay
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


for (h, w), n in (((1080, 1920), 64), ((480, 640), 64), ((1280, 1280), 64), ((480, 640), 1)):
    imgs = [torch.randint(0, 256, (h, w, 3), device=dev, dtype=torch.uint8) for _ in range(n)]
    out = torch.empty(n, 3, 640, 640, device=dev)
    ms_host = timeit(lambda: ops.letterbox(imgs, (640, 640), out=out))
    ms = timeit(lambda: ops.letterbox(imgs, (640, 640), out=out), graph=True)
    _, rh, rw = ops.letterbox_geometry(h, w, (640, 640))
    # algorithmic bytes: canvas written once (fp32 CHW) + every source byte the taps touch, once
    taps = min(h, 2 * rh) * min(w, 2 * rw) * 3
    bytes_ = n * (3 * 640 * 640 * 4 + taps)
    print(json.dumps({"row": "P1 letterbox", "images": n, "source": [h, w], "ms": ms, "ms_python_call": ms_host,
                      "images_per_s": n / ms * 1e3, "GBps": bytes_ / ms / 1e6,
                      "hbm_frac": bytes_ / ms / 1e6 / peaks["hbm_gbs"]}))

# N1 max-sigmoid attention: the three T-CSP layer shapes of the `n` neck at 640^2, C = 1203
for (c, hw_side), n in (((32, 80), 64), ((64, 40), 64), ((128, 20), 64)):
    y = torch.randn(n, c, hw_side, hw_side, device=dev)
    t = torch.randn(1203, c, device=dev)
    out = torch.empty_like(y)
    for precise in (True, False):
        ms = timeit(lambda: ops.max_sigmoid_attention(y, t, precise=precise, out=out), iters=20)
        hw = hw_side * hw_side
        flops = 2.0 * n * hw * 1203 * c
        bytes_ = n * hw * c * 4 * 3 + n * hw * 8          # y read by the GEMM and by the scale, out written
        print(json.dumps({"row": "N1 max-sigmoid attention", "images": n, "channels": c, "hw": hw, "classes": 1203,
                          "precise": precise, "ms": ms, "TFLOPs": flops / ms / 1e9,
                          "GBps": bytes_ / ms / 1e6, "hbm_frac": bytes_ / ms / 1e6 / peaks["hbm_gbs"]}))
    ref = lambda: torch.sigmoid(torch.matmul(y.flatten(2).transpose(1, 2), t.t()).max(dim=-1, keepdim=True)[0])
    ms_t = timeit(ref, iters=5)
    print(json.dumps({"row": "N1 torch (matmul+max+sigmoid only, same GPU)", "channels": c, "ms": ms_t}))
