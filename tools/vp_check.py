"""Vocabulary-parallel head across real GPUs (one process per GPU, CUDA IPC peer mappings over
NVLink): parity against one GPU holding the whole vocabulary, then the time of a step with the
in-kernel key exchange ("fused") against the same step with a library all-reduce ("allreduce").

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/vp_check.py [--classes 4800] [--batch 1] [--steps 200]

Rank 0 prints one JSON line per configuration.  Times are CUDA events on the launching stream,
max over ranks."""
import argparse
import json
import os
import statistics
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ovdet import shard, synth  # noqa: E402
from ovdet import vocab_parallel as vp  # noqa: E402
from ovdet.pipeline import HeadConfig, HeadPipeline  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--classes", type=int, nargs="+", default=[1203, 4800, 19200])
    ap.add_argument("--batch", type=int, nargs="+", default=[1, 16])
    ap.add_argument("--image-size", type=int, default=640)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--assert-parity", action="store_true", help="exit 1 on any mismatch or timeout")
    args = ap.parse_args()
    failed = False
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    s = args.image_size
    shapes = [(s // 8, s // 8), (s // 16, s // 16), (s // 32, s // 32)]
    cfg = HeadConfig(precision="bf16", max_det=300)
    for classes in args.classes:
        for batch in args.batch:
            # the same inputs on every rank (the batch is replicated, the vocabulary sharded)
            inp = synth.make_inputs(batch=batch, image_size=s, num_classes=classes, device=dev, seed=77)
            for t in inp.obj_embeds + inp.box_preds + [inp.text]:
                dist.broadcast(t, 0)                            # byte-identical on every rank
            full = HeadPipeline(batch, shapes, classes, cfg, device=dev)
            full.set_vocabulary(inp.text)
            r = full.run(inp.obj_embeds, inp.box_preds, events={})
            torch.cuda.synchronize()
            want = [t.clone() for t in (full.scores, full.class_ids, r.count, r.anchor)]
            line = {"classes": classes, "batch": batch, "image_size": s, "world": world}
            for mode in ("fused", "allreduce"):
                head = vp.VocabParallelHead(batch, shapes, classes, cfg, device=dev, exchange=mode)
                head.set_vocabulary(inp.text)
                ok = True
                bad_scores = 0
                for _ in range(3):                              # parity over both buffer parities
                    res = head.run(inp.obj_embeds, inp.box_preds)
                    torch.cuda.synchronize()
                    ok &= torch.equal(head.scores, want[0]) and torch.equal(res.count, want[2])
                    same_cls = float((head.class_ids == want[1]).float().mean())
                    bad_scores += int((head.scores != want[0]).sum())
                    for b, k in enumerate(want[2].tolist()):
                        ok &= torch.equal(res.anchor[b, :k], want[3][b, :k])
                for _ in range(args.warmup):
                    head.run(inp.obj_embeds, inp.box_preds)
                dist.barrier()
                torch.cuda.synchronize()
                ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                      for _ in range(args.steps)]
                for a, b in ev:
                    a.record()
                    head.run(inp.obj_embeds, inp.box_preds)
                    b.record()
                torch.cuda.synchronize()
                p50 = statistics.median(a.elapsed_time(b) for a, b in ev)
                total = ev[0][0].elapsed_time(ev[-1][1]) / args.steps
                timed_out = head.timed_out() if mode == "fused" else False
                line[mode] = {"parity": bool(ok), "class_agreement": same_cls, "score_mismatches": bad_scores,
                              "p50_ms": shard.max_over_ranks(p50, dev),
                              "ms_per_step": shard.max_over_ranks(total, dev), "timed_out": timed_out}
                if mode == "fused":
                    # the same step replayed from a CUDA graph (device-resident step counters)
                    head.capture(inp.obj_embeds, inp.box_preds)
                    for _ in range(args.warmup):
                        head.replay()
                    dist.barrier()
                    torch.cuda.synchronize()
                    for a, b in ev:
                        a.record()
                        res = head.replay()
                        b.record()
                    torch.cuda.synchronize()
                    g_ok = torch.equal(head.scores, want[0]) and torch.equal(res.count, want[2])
                    line[mode]["graph_p50_ms"] = shard.max_over_ranks(
                        statistics.median(a.elapsed_time(b) for a, b in ev), dev)
                    line[mode]["graph_parity"] = bool(g_ok) and not head.timed_out()
                    failed |= not line[mode]["graph_parity"]
                failed |= (not ok) or timed_out or same_cls < 1.0
                head.close()
            # one GPU, whole vocabulary, same step
            for _ in range(args.warmup):
                full.run(inp.obj_embeds, inp.box_preds)
            torch.cuda.synchronize()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                  for _ in range(args.steps)]
            for a, b in ev:
                a.record()
                full.run(inp.obj_embeds, inp.box_preds)
                b.record()
            torch.cuda.synchronize()
            line["one_gpu_full_vocabulary"] = {
                "p50_ms": statistics.median(a.elapsed_time(b) for a, b in ev),
                "ms_per_step": ev[0][0].elapsed_time(ev[-1][1]) / args.steps}
            full.capture(inp.obj_embeds, inp.box_preds)
            for _ in range(args.warmup):
                full.replay()
            torch.cuda.synchronize()
            for a, b in ev:
                a.record()
                full.replay()
                b.record()
            torch.cuda.synchronize()
            line["one_gpu_full_vocabulary"]["graph_p50_ms"] = statistics.median(a.elapsed_time(b) for a, b in ev)
            if rank == 0:
                print(json.dumps(line), flush=True)
            del full
            torch.cuda.empty_cache()
    flag = torch.tensor([int(failed)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    dist.destroy_process_group()
    if args.assert_parity and int(flag.item()):
        sys.exit(1)


if __name__ == "__main__":
    main()
