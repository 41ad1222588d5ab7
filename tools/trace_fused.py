"""Clock-stamp timeline of ONE CTA pair of the fused similarity kernel (an OVDET_TRACE build).

    bash tools/build_variant.sh libovdet_trace.so -DOVDET_TRACE
    python tools/trace_fused.py [--logits bf16] [--out gpurun_out/trace.json]

The kernel (csrc/sim_fused_sm100.cu, OVDET_TR) stamps clock64() at the waits of the MMA warp, of
converter warp 4, of the first warp of each epilogue group and of the activation producer of
blockIdx.x == 0.  This script runs the bench workload (batch 256, 640x640, 1203 prompts), reads the
stamps of the last launch and prints, per anchor tile, where the MMA warp's time went: waiting for a
free accumulator (t_empty), for converted activation blocks (a_ready), for text stages (b_full), and
issuing.  A measurement tool; nothing of the product imports it.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logits", default="none")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--classes", type=int, default=1203)
    ap.add_argument("--lib", default="libovdet_trace.so")
    ap.add_argument("--out", default=None)
    ap.add_argument("--single-call", action="store_true", help="trace the product call (ovdet_head_step: class split at batch 1)")
    args = ap.parse_args()
    import torch
    import ovdet
    os.environ["OVDET_LIB_PATH"] = os.path.join(ovdet.PKG_DIR, args.lib)
    dev = torch.device("cuda", 0)
    roles, cap = 5, 32768
    trace = torch.zeros(roles * cap, dtype=torch.int64, device=dev)
    os.environ["OVDET_TRACE_PTR"] = hex(trace.data_ptr())
    from ovdet import synth
    from ovdet.pipeline import HeadConfig, HeadPipeline
    shapes = [(640 // s, 640 // s) for s in (8, 16, 32)]
    inp = synth.make_inputs(batch=args.batch, image_size=640, num_classes=args.classes, device=dev, seed=1234)
    cfg = HeadConfig(precision=args.precision, max_det=300,
                     logits_dtype=None if args.logits == "none" else args.logits)
    pipe = HeadPipeline(args.batch, shapes, args.classes, cfg, device=dev)
    pipe.set_vocabulary(inp.text)
    for _ in range(5):
        pipe.run(inp.obj_embeds, inp.box_preds)
    torch.cuda.synchronize()
    trace.zero_()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev = {}
    if args.single_call:
        ev0.record()
        pipe.run(inp.obj_embeds, inp.box_preds)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
    else:
        pipe.run(inp.obj_embeds, inp.box_preds, events=ev)
        torch.cuda.synchronize()
        ms = ev["similarity"][0].elapsed_time(ev["similarity"][1])
    t = trace.cpu().numpy().astype("uint64").reshape(roles, cap)
    out = {"similarity_ms": ms, "roles": {}}
    mask = (1 << 48) - 1
    for r in range(roles):
        row = t[r]
        n = int((row != 0).sum())
        tags = (row[:n] >> 48).astype("int64")
        clk = (row[:n] & mask).astype("int64")
        out["roles"][r] = (tags, clk)
    if args.out:
        dump = {"similarity_ms": ms,
                "roles": {str(r): {"tags": out["roles"][r][0][:6000].tolist(), "clk": out["roles"][r][1][:6000].tolist()}
                          for r in range(roles)}}
        with open(args.out, "w") as f:
            json.dump(dump, f)
    # ---- MMA warp: split the launch into N tiles (tag 1 ... tag 5) ---------------------------------
    tags, clk = out["roles"][0]
    t_start, t_end = int(clk[0]), int(clk[-1])
    total = t_end - t_start
    wait_tempty = wait_aready = wait_b = issue = 0
    i = 0
    tiles = 0
    first_tiles = []
    per_ntile = []
    while i < len(tags):
        assert tags[i] == 1, (i, tags[i])
        t1 = clk[i]
        t2 = clk[i + 1]
        wait_tempty += t2 - t1
        j = i + 2
        prev = t2
        wa = wb = 0
        had3 = False
        while tags[j] != 5:
            if tags[j] == 3:
                wa += clk[j] - prev
                had3 = True
            elif tags[j] == 4:
                wb += clk[j] - prev
            prev = clk[j]
            j += 1
        wait_aready += wa
        wait_b += wb
        dur = clk[j] - t1
        per_ntile.append(dur)
        if had3:
            first_tiles.append((int(t2 - t1), int(wa), int(wb), int(dur)))
        tiles += 1
        i = j + 1
    n_anchor = len(first_tiles)
    print(f"similarity {ms:.3f} ms; MMA warp of CTA 0: {tiles} N tiles over {n_anchor} anchor-tile pairs, "
          f"{total} cycles ({total / max(1, n_anchor):.0f} per anchor tile)")
    print(f"  waiting t_empty {wait_tempty / total:6.1%}   a_ready {wait_aready / total:6.1%}   "
          f"b_full {wait_b / total:6.1%}   rest (issue + commit) {1 - (wait_tempty + wait_aready + wait_b) / total:6.1%}")
    import numpy as np
    ft = np.array(first_tiles[2:-1] if len(first_tiles) > 4 else first_tiles)
    print("  first N tile of an anchor tile: median wait t_empty / a_ready / b_full / duration =",
          np.median(ft, axis=0).tolist())
    pn = np.array(per_ntile)
    print("  N tile duration: median", float(np.median(pn)), "p90", float(np.percentile(pn, 90)),
          "mean", float(pn.mean()))
    # ---- converter warp 4: per block as_full wait, convert, a_free wait, publish -------------------
    tags, clk = out["roles"][1]
    # transitions between consecutive stamps: 10 top of a block iteration, 11 fp32 block landed (conversion
    # starts), 12 publish begins, 13 A slot released by the MMAs, 14 published
    names = {(10, 11): "poll+wait as_full", (11, 10): "convert", (11, 12): "convert", (10, 12): "poll barriers",
             (12, 13): "wait a_free", (13, 14): "tcgen05.st+arrive", (14, 11): "wait as_full (rest)",
             (14, 12): "between publishes", (14, 10): "norm / loop"}
    acc = {}
    for k in range(len(tags) - 1):
        key = (int(tags[k]), int(tags[k + 1]))
        a = acc.setdefault(key, [0, 0])
        a[0] += int(clk[k + 1] - clk[k])
        a[1] += 1
    span = int(clk[-1] - clk[0]) if len(clk) else 1
    print(f"  converter warp 4 ({span} cycles):")
    for key, (tot, n) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
        print(f"    {names.get(key, str(key)):24s} {tot / span:6.1%}  mean {tot / n:7.0f} cycles x {n}")
    # ---- epilogue groups ---------------------------------------------------------------------------
    for r in (2, 3):
        tags, clk = out["roles"][r]
        if len(tags) < 4:
            continue
        w = dr = tail = 0
        for k in range(len(tags) - 1):
            a, b = tags[k], tags[k + 1]
            dt = int(clk[k + 1] - clk[k])
            if (a, b) == (20, 21):
                w += dt
            elif (a, b) == (21, 22):
                dr += dt
            elif (a, b) == (22, 23):
                tail += dt
        fine = {}
        for k in range(len(tags) - 1):
            key = (int(tags[k]), int(tags[k + 1]))
            if key[0] >= 22 or key[1] >= 24:
                a = fine.setdefault(key, [0, 0])
                a[0] += int(clk[k + 1] - clk[k])
                a[1] += 1
        if any(k[1] >= 24 for k in fine):
            print(f"  epilogue group {r - 2} store phase (22 release, 24 buffer free, 25 staged, 26 fenced, 23 done):",
                  {f"{a}->{b}": round(t / n) for (a, b), (t, n) in sorted(fine.items())})
        span = int(clk[-1] - clk[0])
        n = int((tags == 21).sum())
        print(f"  epilogue group {r - 2}: wait t_full {w / span:.1%}, drain-to-release {dr / span:.1%} "
              f"({dr / max(1, n):.0f} cycles per tile), after release {tail / span:.1%} ({tail / max(1, n):.0f}); {n} tiles")
    # ---- activation producer -----------------------------------------------------------------------
    tags, clk = out["roles"][4]
    if len(clk) > 16:
        gaps = np.diff(clk)
        print("  activation producer: issue gaps median", float(np.median(gaps)), "p90", float(np.percentile(gaps, 90)))
    if args.out:
        dump = {"similarity_ms": ms,
                "roles": {str(r): {"tags": out["roles"][r][0][:6000].tolist(), "clk": out["roles"][r][1][:6000].tolist()}
                          for r in range(roles)}}
        with open(args.out, "w") as f:
            json.dump(dump, f)


if __name__ == "__main__":
    main()
