"""Where do class shards and one full-vocabulary launch disagree?  (virtual ranks, one GPU)"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from ovdet import synth, ops
from ovdet import vocab_parallel as vp
from ovdet.pipeline import HeadConfig, HeadPipeline

dev = torch.device("cuda:0")
B, C, W = 16, 1203, 2
shapes = [(80, 80), (40, 40), (20, 20)]
cfg = HeadConfig(precision="bf16", max_det=300)
x = synth.make_inputs(batch=B, image_size=640, num_classes=C, device=dev, seed=77)
full = HeadPipeline(B, shapes, C, cfg, device=dev)
full.set_vocabulary(x.text)
r = full.run(x.obj_embeds, x.box_preds, events={})
torch.cuda.synchronize()
ws, wc, wcount = full.scores.clone(), full.class_ids.clone(), r.count.clone()
heads = [vp.VocabParallelHead(B, shapes, C, cfg, device=dev, rank=k, world=W) for k in range(W)]
for h in heads:
    h.connect([g.buffer.ptr for g in heads]); h.set_vocabulary(x.text)
for h in heads: h.similarity(x.obj_embeds)
for h in heads: h.signal()
heads[0].merge(); res = heads[0].finish(x.box_preds)
torch.cuda.synchronize()
ds = (heads[0].scores != ws); dc = (heads[0].class_ids != wc)
print("score mismatches", int(ds.sum()), "class mismatches", int(dc.sum()), "count equal", torch.equal(res.count, wcount))
# fp32 logits of the full vocabulary for the disagreeing rows
logits, _, _ = ops.similarity_fused(x.obj_embeds, full.text_op, 1.0, 0.0, logits_dtype=torch.float32, want_max=False)
torch.cuda.synchronize()
idx = dc.nonzero()
for b, a in idx[:10].tolist():
    row = logits[b, a]
    c_full, c_vp = int(wc[b, a]), int(heads[0].class_ids[b, a])
    print((b, a), "full", c_full, float(row[c_full]), "vp", c_vp, float(row[c_vp]), "score", float(ws[b, a]), float(heads[0].scores[b, a]),
          "top2", row.topk(2).values.tolist())
idx = ds.nonzero()
for b, a in idx[:10].tolist():
    print("score", (b, a), float(ws[b, a]), float(heads[0].scores[b, a]))
for h in heads: h.close()
