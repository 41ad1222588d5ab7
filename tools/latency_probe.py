"""Batch-1 latency breakdown of the K1..K4 step: per-stage GPU time (CUDA events), host time per
call, and the same step replayed from a CUDA graph."""
import os, sys, time, statistics, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ovdet import synth
from ovdet.pipeline import HeadConfig, HeadPipeline

dev = torch.device("cuda:0")
shapes = [(80, 80), (40, 40), (20, 20)]
inp = synth.make_inputs(batch=1, image_size=640, num_classes=1203, device=dev, seed=77)
pipe = HeadPipeline(1, shapes, 1203, HeadConfig(precision="bf16", max_det=300), device=dev)
pipe.set_vocabulary(inp.text)
for _ in range(20):
    pipe.run(inp.obj_embeds, inp.box_preds)
torch.cuda.synchronize()
stages = {k: [] for k in ("l2norm", "similarity", "decode", "nms")}
total, host = [], []
for _ in range(200):
    ev = {}
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    pipe.run(inp.obj_embeds, inp.box_preds, events=ev)
    b.record()
    host.append((time.perf_counter() - t0) * 1e3)
    b.synchronize()
    total.append(a.elapsed_time(b))
    for k in stages:
        stages[k].append(ev[k][0].elapsed_time(ev[k][1]))
plain, host1 = [], []
for _ in range(200):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    t0 = time.perf_counter()
    pipe.run(inp.obj_embeds, inp.box_preds)
    host1.append((time.perf_counter() - t0) * 1e3)
    b.record(); b.synchronize()
    plain.append(a.elapsed_time(b))
out = {"p50_ms_with_stage_events": statistics.median(total), "p50_ms": statistics.median(plain),
       "host_ms_per_call_p50": statistics.median(host), "host_ms_single_call_p50": statistics.median(host1),
       "stage_p50_ms": {k: statistics.median(v) for k, v in stages.items()}}
# CUDA graph replay of the same step
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    pipe.run(inp.obj_embeds, inp.box_preds)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        pipe.run(inp.obj_embeds, inp.box_preds)
torch.cuda.synchronize()
lat = []
for _ in range(200):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); b.synchronize()
    lat.append(a.elapsed_time(b))
out["p50_ms_cuda_graph"] = statistics.median(lat)
print(json.dumps(out))
