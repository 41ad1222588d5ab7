import os, sys, statistics, json
sys.path.insert(0, "/root/repo")
import torch
from ovdet import synth, ops
from ovdet.pipeline import HeadConfig, HeadPipeline
dev = torch.device("cuda:0")
shapes = [(80, 80), (40, 40), (20, 20)]
inp = synth.make_inputs(batch=1, image_size=640, num_classes=1203, device=dev, seed=77)
pipe = HeadPipeline(1, shapes, 1203, HeadConfig(precision="bf16", max_det=300), device=dev)
pipe.set_vocabulary(inp.text)
for _ in range(10): pipe.run(inp.obj_embeds, inp.box_preds)
torch.cuda.synchronize()
def p50_graph(fn, n=300):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)
res = {}
res["full_step"] = p50_graph(lambda: pipe.run(inp.obj_embeds, inp.box_preds))
ws = pipe._sim_ws if hasattr(pipe, "_sim_ws") else None
res["similarity_only"] = p50_graph(lambda: ops.similarity_fused(inp.obj_embeds, pipe.text_op, 1.0, 0.0, logits_dtype=None, want_max=True, row_max=pipe.scores, row_arg=pipe.class_ids, inv_norm=pipe.inv_norm))
res["decode_only"] = p50_graph(lambda: ops.decode_filter(inp.box_preds, (8, 16, 32), scores=pipe.scores, conf=0.25, boxes=pipe.boxes, pass_mask=pipe.pass_mask))
res["nms_only"] = p50_graph(lambda: ops.nms_batched(pipe.boxes, pipe.scores, pipe.class_ids, pipe.pass_mask, iou_thr=0.45, max_det=300, out=pipe.result, workspace=pipe.workspace))
res["empty_graph_kernel"] = p50_graph(lambda: pipe.scores.add_(0.0))
print(json.dumps(res))
if os.environ.get("OVDET_LIB_PATH", "").endswith("nmstr.so"):
    for _ in range(5):
        ops.nms_batched(pipe.boxes, pipe.scores, pipe.class_ids, pipe.pass_mask, iou_thr=0.45, max_det=300, out=pipe.result, workspace=pipe.workspace)
    torch.cuda.synchronize()
    t = pipe.result.keep[0, 280:289].cpu().numpy().astype("int64") & 0xffffffff
    d = [(int(t[i + 1]) - int(t[i])) & 0xffffffff for i in range(8)]
    print(json.dumps({"nms_phase_cycles": dict(zip(["pdl_wait", "prefix_scan", "keys", "sort", "load_boxes", "mask", "resolve", "emit"], d)),
                      "candidates": int(pipe.result.candidates[0]), "kept": int(pipe.result.count[0])}))
