"""Does the HBM-bound decode kernel hide under the tensor-bound fused similarity kernel when the
two run on different streams?  Sequential vs forked timing at the benchmark shape."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ovdet import ops, synth

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
inp = synth.make_inputs(batch=B, image_size=640, num_classes=1203, device=dev, seed=1)
top = ops.l2norm_text(inp.text)
A = inp.num_anchors
rmax = torch.empty(B, A, device=dev); rarg = torch.empty(B, A, device=dev, dtype=torch.int32)
boxes = torch.empty(B, A, 4, device=dev)
side = torch.cuda.Stream()
main = torch.cuda.current_stream()

def seq():
    ops.similarity_fused(inp.obj_embeds, top, logits_dtype=None, want_max=True, row_max=rmax, row_arg=rarg)
    ops.decode_filter(inp.box_preds, inp.strides, boxes=boxes)

def fork(decode_first):
    ev = torch.cuda.Event(); ev.record(main)
    if not decode_first:
        ops.similarity_fused(inp.obj_embeds, top, logits_dtype=None, want_max=True, row_max=rmax, row_arg=rarg)
    with torch.cuda.stream(side):
        side.wait_event(ev)
        ops.decode_filter(inp.box_preds, inp.strides, boxes=boxes)
        done = torch.cuda.Event(); done.record(side)
    if decode_first:
        ops.similarity_fused(inp.obj_embeds, top, logits_dtype=None, want_max=True, row_max=rmax, row_arg=rarg)
    main.wait_event(done)

def timeit(fn, iters=60):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters

for _ in range(2):
    print(json.dumps({"batch": B, "sequential_ms": timeit(seq), "fork_fused_first_ms": timeit(lambda: fork(False)),
                      "fork_decode_first_ms": timeit(lambda: fork(True))}))
