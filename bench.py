#!/usr/bin/env python
"""Benchmark of the open-vocabulary head + post-processing hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W]              # B200 arm (libovdet.so)
    python bench.py --impl reference [--steps K] [--warmup W]        # reference CPU arm

Workload (BASELINE.json configs[2]): batch 256 @ 640x640 (8400 anchors: 80^2 | 40^2 | 20^2),
1203 prompts, D = 512, bf16 operands / fp32 accumulate, class max fused in the GEMM epilogue,
conf 0.25, IoU 0.45; synthetic conv outputs (ovdet.synth) and a synthetic
shared vocabulary.  One process per GPU, the batch is the sharded unit, no collective on the
data path; per-GPU work is fixed as N grows ("weak").

One step = K1b (L2 norm of the text rows) -> K1+K2 fused (L2 norm + tcgen05 similarity GEMM +
class max/argmax straight from the fp32 NCHW conv outputs, 1 launch) -> K3 (DFL decode + threshold) -> K4 (gather / sort / NMS) over
one batch (`--no-fused`: K1 as 3 launches writing a bf16 operand, then the K2 GEMM).  `value` times the steps
through the product call (HeadPipeline.run -> ONE C call, ovdet_head_step, K3 / K4 under programmatic
dependent launch) with the inputs resident in HBM; the per-kernel times behind `roofline` / `stages_ms`
come from a second pass of the same number of steps, launched stage by stage with CUDA events between
the kernels; `e2e` times Detector.predict_host on pinned HOST buffers with the H2D copies and the D2H of
the detections inside the timed region.  `config` holds only what defines the workload and is identical
in both arms; what a run observed (candidates, kept, path taken, index mismatches against the oracle)
is under `observed` / `parity`.

    python bench.py --per-image-text          # text [B, C, 512] with the neck's strides (repvl_pan.py:173-182)
    python bench.py --mode vocab-parallel     # N > 1: prompts sharded over the GPUs, in-kernel exchange
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMAGE_SIZE = 640
NUM_CLASSES = 1203
EMBED_DIM = 512
BATCH_PER_GPU = 256
MAX_DET = 300
STRIDES = (8, 16, 32)
METRIC = "images/sec (640x640, 1203 prompts), head + post-process"
UNIT = "images/s"


PROJECTED_NOTE = " + the head's 1x1 projection (step starts at the hidden features, SURVEY 8f-2)"


def workload_name(batch):
    which = {(640, 80): "configs[1] shapes", (640, 1203): "configs[2]", (1280, 1203): "configs[3] shapes",
             (640, 4800): "configs[4] shapes"}.get((IMAGE_SIZE, NUM_CLASSES), "custom")
    return (f"batch {batch}/GPU @ {IMAGE_SIZE}x{IMAGE_SIZE}, {NUM_CLASSES} prompts, bf16 similarity GEMM "
            f"(BASELINE.json {which})")


def static_config(args, n_gpus):
    """What defines the workload - the same dict in both arms (the driver compares them)."""
    batch = args.batch
    shapes = [(IMAGE_SIZE // s, IMAGE_SIZE // s) for s in STRIDES]
    anchors = sum(h * w for h, w in shapes)
    in_ch = 256 if args.projected else EMBED_DIM
    esz = 2 if args.input_dtype == "bf16" else 4
    input_bytes = batch * anchors * (in_ch + 68) * esz
    return {"workload": workload_name(batch) + (PROJECTED_NOTE if args.projected else ""),
            "global_batch": batch * n_gpus, "anchors": anchors, "classes": NUM_CLASSES, "embed_dim": EMBED_DIM,
            "precision": ("bf16 operands, fp32 accumulate, fused class max/argmax" if args.precision == "bf16"
                          else "fp16 operands (per-row power-of-two scaling), fp32 accumulate, |dlogit| <~ 1e-4"
                          if args.precision == "fp16"
                          else "three bf16 passes over hi/lo operand halves (|dlogit| ~ 1e-5), fp32 accumulate"),
            "text": ("per-image [B, C, 512], strides (512, B*512, 1) as the neck emits it, normalised every step"
                     if args.per_image_text else "shared vocabulary [C, 512], re-normalised every step"),
            "materialised_logits": args.logits, "region_embedding_dtype": args.input_dtype,
            "conf": 0.25, "iou": 0.45, "max_det": args.max_det,
            "parallelism": f"batch-sharded x{n_gpus}, vocabulary replicated, no collective",
            "l2": f"inputs are {input_bytes / 1e9:.2f} GB per step per GPU (> 126 MB L2), no flush needed"}


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap,enforced.power.limit,clocks_event_reasons.active")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def mark(self):
        """Host time stamp; rows are kept only between the two marks given to stop()."""
        return time.perf_counter()

    def in_window(self, t0, t1=None):
        return sum(1 for t, _ in self.rows if t >= t0 and (t1 is None or t <= t1))

    def stop(self, t0=None, t1=None, window="timed region"):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        power, limit, masks = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, row in self.rows:
            if (t0 is not None and t < t0) or (t1 is not None and t > t1):
                continue
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
            try:
                power.append(float(parts[2]))
                if len(parts) > 7:
                    limit = float(parts[7])
            except ValueError:
                pass
            if len(parts) > 8:
                masks.add(parts[8])
        # the per-reason flags are sampled states and can read "Not Active" between two cap events; the
        # board power beside its enforced limit and the raw reasons bitmask are reported as well
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "samples": len(sm), "window": window, "reasons": sorted(reasons),
                "power_w": statistics.median(power) if power else None, "power_limit_w": limit,
                "reasons_bitmask": sorted(masks)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tflops": p["bf16_tflops_sustained"], "tflops_burst": p["bf16_tflops"],
                "hbm_gbs": p["hbm_gbs"], "source": "measured"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


NCU_CAPTURE = "r2_ncu_full_main_b256.json"     # ncu --set full of `bench.py --profile` at the current build


def ncu_traffic(kernel_substr: str, batch: int):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel
    from the committed `ncu --set full` capture of this same workload (profiles/, batch 256);
    None when the run's shape differs from the captured one."""
    path = os.path.join(ROOT, "profiles", NCU_CAPTURE)
    if batch != 256 or (IMAGE_SIZE, NUM_CLASSES) != (640, 1203) or not os.path.exists(path):
        return None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    with open(path) as f:
        for rec in json.load(f)["launches"]:
            if kernel_substr in rec["kernel"]:
                rd, wr = rec["dram__bytes_read.sum"], rec["dram__bytes_write.sum"]
                return rd["value"] * scale[rd["unit"]] + wr["value"] * scale[wr["unit"]]
    return None


# ----------------------------------------------------------------------------------------------
# reference CPU arm / cpu_baseline leg (the only places that execute oracle/)
# ----------------------------------------------------------------------------------------------
def cpu_reference_sample(sample_images: int, seed: int = 1234, projected: bool = False,
                         per_image_text: bool = False, host_inputs=None):
    """Build a bounded sample of the workload on the host and return a callable running the
    reference algorithm (oracle port of yolo_clip.py:173-214 + detector.py:163-223) over it;
    ``projected``: the step starts one layer earlier, at the input of the head's 1x1 projection
    (text_contrastive.py:67,112), which the reference computes as a convolution.  ``host_inputs``
    (obj_embeds, box_preds, text): run on these tensors (the first images of the device batch)
    instead of generating a sample."""
    import torch
    from oracle import ref_port
    from ovdet import synth
    sizes = [(IMAGE_SIZE, IMAGE_SIZE)] * sample_images
    scales = [1.0] * sample_images
    if projected:
        pin = synth.make_projected_inputs(batch=sample_images, image_size=IMAGE_SIZE, num_classes=NUM_CLASSES,
                                          embed_dim=EMBED_DIM, device="cpu", seed=seed)
        ptext = pin.text_batched()

        def pstep():
            with torch.no_grad():
                embeds = [torch.nn.functional.conv2d(h, w, b) for h, w, b in zip(pin.hidden, pin.weights, pin.biases)]
                tail = ref_port.head_tail(embeds, ptext, pin.box_preds, STRIDES)
                return ref_port.postprocess_batch(tail, sizes, scales)
        return pstep
    if host_inputs is not None:
        obj_embeds, box_preds, text = host_inputs
    else:
        inp = synth.make_inputs(batch=sample_images, image_size=IMAGE_SIZE, num_classes=NUM_CLASSES,
                                embed_dim=EMBED_DIM, device="cpu", seed=seed)
        obj_embeds, box_preds = inp.obj_embeds, inp.box_preds
        text = per_image_text_like(inp.text, sample_images) if per_image_text else inp.text_batched()
    if text.dim() == 2:
        text = text.unsqueeze(0).expand(sample_images, -1, -1)

    def step():
        with torch.no_grad():
            tail = ref_port.head_tail(obj_embeds, text, box_preds, STRIDES)
            return ref_port.postprocess_batch(tail, sizes, scales)
    return step


def per_image_text_like(text, batch, seed: int = 777):
    """Per-image text embeddings the way the neck hands them over (repvl_pan.py:173-182:
    ``text + attention(text, image patches)``): every image's rows differ a little from the shared
    vocabulary, memory is [C, B, D] viewed as [B, C, D] - strides (D, B*D, 1)."""
    import torch
    g = torch.Generator(device=text.device).manual_seed(seed)
    c, d = text.shape
    out = torch.empty(c, batch, d, device=text.device, dtype=text.dtype)
    out.normal_(generator=g).mul_(0.05).add_(text.unsqueeze(1))
    return out.transpose(0, 1)


REF_SAMPLE = 16          # images per step of the reference arm / of the cpu_baseline leg


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = REF_SAMPLE
    step = cpu_reference_sample(sample, projected=args.projected, per_image_text=args.per_image_text)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    sample_desc = (f"{sample} images/step of the same synthetic workload (the metric is per image; a step of "
                   f"{args.batch} images would take {args.batch / value:.1f} s), reference algorithm "
                   f"(torch CPU similarity+max+decode, numpy NMS, fp32, no detection cap) via oracle/ref_port.py")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": static_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from ovdet import shard, synth
    from ovdet.detector import Detector
    from ovdet.pipeline import HeadConfig, HeadPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # one rank per GPU, each on its own slice of the host cores: the pinned staging buffers of the e2e
    # leg are then first touched (and later read by the DMA engines) from distinct cores / memory
    if world > 1 and hasattr(os, "sched_setaffinity"):
        try:
            cpus = sorted(os.sched_getaffinity(0))
            per = max(1, len(cpus) // world)
            os.sched_setaffinity(0, set(cpus[local * per:(local + 1) * per]) or set(cpus))
        except OSError:
            pass
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world
    batch = args.batch
    shapes = [(IMAGE_SIZE // s, IMAGE_SIZE // s) for s in STRIDES]
    anchors = sum(h * w for h, w in shapes)

    cfg = HeadConfig(precision=args.precision, max_det=args.max_det, fused=not args.no_fused,
                     logits_dtype=None if args.logits == "none" else args.logits)
    projections = None
    if args.projected:
        pin = synth.make_projected_inputs(batch=batch, image_size=IMAGE_SIZE, num_classes=NUM_CLASSES,
                                          embed_dim=EMBED_DIM, device=dev, seed=1234 + rank)
        projections = pin.projections()
        # same field names as HeadInputs; obj_embeds = the hidden features
        inp = SimpleNamespace(obj_embeds=pin.hidden, box_preds=pin.box_preds, text=pin.text)
    else:
        inp = synth.make_inputs(batch=batch, image_size=IMAGE_SIZE, num_classes=NUM_CLASSES,
                                embed_dim=EMBED_DIM, device=dev, seed=1234 + rank)
    if args.input_dtype == "bf16":
        # the head convolutions ran under autocast: bf16 region embeddings and box logits
        inp = SimpleNamespace(obj_embeds=[e.to(torch.bfloat16) for e in inp.obj_embeds],
                              box_preds=[b.to(torch.bfloat16) for b in inp.box_preds], text=inp.text)
    pipe = HeadPipeline(batch, shapes, NUM_CLASSES, cfg, device=dev, projections=projections,
                        per_image_text=args.per_image_text)
    # the vocabulary is replicated: rank 0's copy goes to every GPU once, outside the timed region
    vocab = shard.broadcast_vocabulary(inp.text if rank == 0 else None, NUM_CLASSES, EMBED_DIM, dev)
    # The text rows are re-normalised inside every timed step (K1b, text_contrastive.py:138), as the
    # reference does on every forward, although the vocabulary is constant.  Projected mode: the
    # projected operand (text x the 1x1 conv weights) is a per-vocabulary precompute, like the
    # reference's offline vocabulary.  --per-image-text: every image has its own [C, 512] rows, laid
    # out with the neck's strides; K1b then runs over batch * C rows per step.
    if args.per_image_text:
        step_text = per_image_text_like(vocab, batch, seed=777 + rank)
        assert step_text.stride() == (EMBED_DIM, batch * EMBED_DIM, 1)
    else:
        pipe.set_vocabulary(vocab)
        step_text = None if args.projected else vocab
    # detector.py:193-202 rescales and clips every candidate box to the original image: the step does too
    sizes, scales = [(IMAGE_SIZE, IMAGE_SIZE)] * batch, [1.0] * batch
    pipe.set_geometry(sizes, scales)
    input_bytes = sum(t.numel() * t.element_size() for t in inp.obj_embeds + inp.box_preds)
    if args.per_image_text:
        input_bytes += step_text.numel() * step_text.element_size()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value"): the product call, nothing between the kernels ------
    # nvidia-smi needs ~0.1-0.3 s to print its first row: start it before the warm-up, keep only
    # the rows that arrive between the marks around the timed region
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        pipe.run(inp.obj_embeds, inp.box_preds, text=step_text)
    barrier()
    res = pipe.result
    kept = res.count.float().mean().item()
    cand = res.candidates.float().mean().item()
    overflow = int((res.count >= args.max_det).sum().item())

    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_mark0 = sampler.mark()
    start.record()
    for _ in range(args.steps):
        pipe.run(inp.obj_embeds, inp.box_preds, text=step_text)
    stop.record()
    barrier()
    elapsed_ms = start.elapsed_time(stop)
    timed_path = pipe.last_path
    single_call = pipe.last_single_call

    # ---- the same steps again, stage by stage, with CUDA events between the kernels ---------------
    stage_events = []
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(args.steps):
        ev = {}
        pipe.run(inp.obj_embeds, inp.box_preds, text=step_text, events=ev)
        stage_events.append(ev)
    s1.record()
    barrier()
    staged_ms = s0.elapsed_time(s1)
    clocks = None
    if rank == 0:
        t_mark1 = sampler.mark()
        window = "timed region + the per-stage pass of the same steps"
        if sampler.proc is not None and sampler.in_window(t_mark0, t_mark1) < 5:
            # a timed region shorter than a few sampling periods: keep the same steps running
            # (untimed) until there are 5 rows under this load, at most 3 s
            window += " + untimed continuation of the same steps"
            t_end = time.perf_counter() + 3.0
            while sampler.in_window(t_mark0) < 5 and time.perf_counter() < t_end:
                pipe.run(inp.obj_embeds, inp.box_preds, text=step_text)
                torch.cuda.synchronize()
            t_mark1 = sampler.mark()
        clocks = sampler.stop(t_mark0, t_mark1, window)
    elapsed_ms = shard.max_over_ranks(elapsed_ms, dev)
    value = n_gpus * batch * args.steps / (elapsed_ms / 1e3)

    def stage_ms(name):
        return statistics.mean(e[name][0].elapsed_time(e[name][1]) for e in stage_events)
    stages = {k: stage_ms(k) for k in ("l2norm", "similarity", "decode", "nms")}

    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_only": True, "value": value, "stages_ms": stages}), flush=True)
        return

    # ---- end to end through the public API with HOST buffers ---------------------------------
    chunk = min(args.e2e_chunk, batch)
    assert batch % chunk == 0
    host_obj = [torch.empty((batch,) + tuple(t.shape[1:]), dtype=t.dtype).pin_memory() for t in inp.obj_embeds]
    host_box = [torch.empty((batch,) + tuple(t.shape[1:]), dtype=t.dtype).pin_memory() for t in inp.box_preds]
    for h, d in zip(host_obj + host_box, inp.obj_embeds + inp.box_preds):
        h.copy_(d)
    torch.cuda.synchronize()
    det = Detector(device=str(dev), config=cfg)
    det.set_vocabulary(vocab)
    out_host = None

    def e2e_step():
        nonlocal out_host
        out_host = det.predict_host(host_obj, host_box, chunk=chunk, projections=projections,
                                    orig_sizes=sizes, scale_factors=scales)

    e2e = None
    if not args.per_image_text:           # predict_host serves a shared vocabulary
        for _ in range(max(1, min(args.warmup, 3))):
            e2e_step()
        barrier()
        e2e_steps = max(1, args.e2e_steps)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(e2e_steps + 1)]
        marks[0].record()
        for i in range(e2e_steps):
            e2e_step()                    # returns after the step's detections are in host memory
            marks[i + 1].record()
        barrier()
        e2e_ms = shard.max_over_ranks(marks[0].elapsed_time(marks[-1]), dev)
        e2e_value = n_gpus * batch * e2e_steps / (e2e_ms / 1e3)
        d2h_bytes = sum(v.numel() * v.element_size() for v in out_host.values())
        step_gbs = sorted(input_bytes / (marks[i].elapsed_time(marks[i + 1]) * 1e-3) / 1e9 for i in range(e2e_steps))
        gbs_min = -shard.max_over_ranks(-step_gbs[0], dev)                 # slowest step of the slowest rank
        gbs_med = -shard.max_over_ranks(-statistics.median(step_gbs), dev)  # median step of the slowest rank
        e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": input_bytes,
               "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps, "chunk_images": chunk,
               "h2d_GBps_achieved_per_gpu": e2e_value / n_gpus * (input_bytes / batch) / 1e9,
               "h2d_GBps_per_gpu_min_over_steps_slowest_rank": gbs_min,
               "h2d_GBps_per_gpu_median_over_steps_slowest_rank": gbs_med,
               "h2d_note": "PCIe-bound: pinned host->device copies of one B200 peak at ~55.6 GB/s (Gen5 x16, "
                           "tools/pcie_probe.py); copies and kernels overlap on two streams; with several ranks "
                           "the host's memory system is shared (tools/pcie_probe.py --all-gpus)",
               "api": "ovdet.detector.Detector.predict_host (pinned host buffers in, detections out)"}

    # ---- batch-1 latency (the second half of BASELINE.json's metric) -------------------------
    p50 = p50_graph = None
    if rank == 0 and not args.per_image_text:
        if args.projected:
            pone = synth.make_projected_inputs(batch=1, image_size=IMAGE_SIZE, num_classes=NUM_CLASSES,
                                               embed_dim=EMBED_DIM, device=dev, seed=1234 + rank)
            one = SimpleNamespace(obj_embeds=pone.hidden, box_preds=pone.box_preds)
        else:
            one = synth.make_inputs(batch=1, image_size=IMAGE_SIZE, num_classes=NUM_CLASSES,
                                    embed_dim=EMBED_DIM, device=dev, seed=77)
            if args.input_dtype == "bf16":
                one = SimpleNamespace(obj_embeds=[e.to(torch.bfloat16) for e in one.obj_embeds],
                                      box_preds=[b.to(torch.bfloat16) for b in one.box_preds])
        pipe1 = HeadPipeline(1, shapes, NUM_CLASSES, cfg, device=dev, projections=projections)
        pipe1.set_vocabulary(vocab)
        pipe1.set_geometry(sizes[:1], scales[:1])
        lat = []
        for i in range(60):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            pipe1.run(one.obj_embeds, one.box_preds)
            b.record()
            b.synchronize()
            if i >= 10:
                lat.append(a.elapsed_time(b))
        p50 = statistics.median(lat)
        # the same step replayed from a CUDA graph (the serving configuration: fixed input buffers)
        pipe1.capture(one.obj_embeds, one.box_preds)
        torch.cuda.synchronize()
        lat = []
        for i in range(110):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            pipe1.replay()
            b.record()
            b.synchronize()
            if i >= 10:
                lat.append(a.elapsed_time(b))
        p50_graph = statistics.median(lat)

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only); the same pass is the checker of
    # the device results at bench size: the first REF_SAMPLE images of the device batch go through the
    # oracle and the kept anchor sets are compared ----------------------------------------------------
    cpu = parity = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        import numpy as np
        import torch as _t
        from oracle import ref_port
        cores = os.cpu_count() or 1
        _t.set_num_threads(cores)
        sample = min(REF_SAMPLE, batch)
        if args.projected:
            step = cpu_reference_sample(sample, projected=True)
        else:
            host_in = ([e[:sample].float().cpu() for e in inp.obj_embeds], [b[:sample].float().cpu() for b in inp.box_preds],
                       step_text[:sample].cpu().contiguous() if args.per_image_text else vocab.cpu())
            step = cpu_reference_sample(sample, host_inputs=host_in)
            pipe.run(inp.obj_embeds, inp.box_preds, text=step_text)
            torch.cuda.synchronize()
            got = pipe.result
            want = step()
            mism = [len(set(got.anchor[i, :int(got.count[i])].tolist()) ^ set(int(a) for a in want[i]["anchor_idx"]))
                    for i in range(sample)]
            kept_ref = [len(want[i]["anchor_idx"]) for i in range(sample)]
            # the oracle fed with the device's own scores / boxes: post-processing must be bit-exact
            fed = {"boxes": pipe.boxes[:sample].cpu(), "scores": pipe.scores[:sample].cpu(),
                   "class_ids": pipe.class_ids[:sample].cpu().long()}
            exact = ref_port.postprocess_batch(fed, [(IMAGE_SIZE, IMAGE_SIZE)] * sample, [1.0] * sample)
            bit_exact = all(np.array_equal(got.anchor[i, :int(got.count[i])].cpu().numpy(),
                                           exact[i]["anchor_idx"][:args.max_det]) for i in range(sample))
            parity = {"images_checked": sample, "e2e_index_mismatches": int(sum(mism)),
                      "e2e_index_mismatches_per_image_max": int(max(mism)), "oracle_kept_total": int(sum(kept_ref)),
                      "postprocess_bit_exact_on_device_scores": bool(bit_exact),
                      "note": "kept anchor sets: device path vs oracle scores -> oracle NMS on the first images of "
                              "this run's batch (symmetric difference); differences need two overlapping "
                              "candidates whose scores are closer than the similarity error "
                              f"({{'fp32': '~1e-5', 'fp16': '<= 1e-4, fp16', 'bf16': '<= 4e-3, bf16'}}[args.precision])"}
        step()
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 12.0:
            step()
            reps += 1
        dt = time.perf_counter() - t0
        cpu = {"value": sample * reps / dt, "unit": UNIT, "cores": _t.get_num_threads(), "kind": "port",
               "sample": f"the first {sample} images of this run's batch x {reps} passes through "
                         f"oracle/ref_port.py (torch CPU similarity/max/decode + numpy NMS, fp32, no detection cap)"}

    if rank == 0:
        peaks = measured_peaks()
        flops = 2.0 * batch * anchors * NUM_CLASSES * EMBED_DIM * (3 if args.precision == "fp32" else 1)
        achieved = flops / (stages["similarity"] * 1e-3) / 1e12
        proj = timed_path == "projected"
        fused = timed_path in ("fused", "fused_fp32")
        k1b = 0 if proj or (step_text is None) else 1                          # text rows, every step
        launches = (3 if (fused or proj) else len(shapes) + 3) + k1b
        if args.precision == "fp16" and fused:
            timed_path = "fused_fp16"
        kernel = ("sim_fused_kernel, projected mode (1x1 projection folded: hidden fp32 NCHW in, quadratic-form "
                  "norm, tcgen05 GEMM K = 272, class max/argmax)" if proj else
                  "sim_fused_kernel, fp32-accurate three-pass mode" if timed_path == "fused_fp32" else
                  "sim_fused_kernel, fp16 operand tier (K1+K2: fp32 NCHW in, row scaling, L2 norm, tcgen05 GEMM, class "
                  "max/argmax)" if timed_path == "fused_fp16" else
                  "sim_fused_kernel (K1+K2: fp32 NCHW in, L2 norm, tcgen05 GEMM, class max/argmax)" if fused
                  else "sim_gemm_kernel (K2)")
        if proj:
            # MMA work the kernel issues per launch: [A x 272] x [272 x (classes padded to 128 + 272 rows of G')]
            flops = 2.0 * batch * anchors * ((NUM_CLASSES + 127) // 128 * 128 + 272) * 272
            achieved = flops / (stages["similarity"] * 1e-3) / 1e12
        # dominant kernel's algorithmic bytes: the fused kernel reads the fp32 activations once; the
        # two-kernel path's GEMM reads the bf16 operand (hi|lo halves for the fp32 recipe)
        kop_bytes = EMBED_DIM * 2 * (2 if args.precision == "fp32" else 1)
        in_esz = 2 if args.input_dtype == "bf16" else 4
        text_sets = batch if args.per_image_text else 1
        alg_bytes = (batch * anchors * (EMBED_DIM * in_esz + 12) if fused else batch * anchors * (kop_bytes + 12)) \
            + text_sets * NUM_CLASSES * kop_bytes
        if args.logits != "none":
            alg_bytes += batch * anchors * NUM_CLASSES * (2 if args.logits == "bf16" else 4)
        if proj:
            alg_bytes = batch * anchors * (256 * 4 + 12) + 3 * (NUM_CLASSES + 272) * 272 * 2
        # which measured peak: a timed region shorter than a second is a burst (the clock has not yet
        # settled at the power cap), a long one is sustained; both fractions are reported
        # (and only if the clock actually stayed up: a power-capped board is in its sustained regime)
        clock_up = bool(clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz")
                        and clocks["sm_mhz"] >= 0.95 * clocks["sm_max_mhz"])
        burst = (elapsed_ms + staged_ms) < 1000.0 and clock_up
        peak_tf = peaks["tflops_burst"] if burst else peaks["tflops"]
        t_tensor = flops / (peaks["tflops"] * 1e12)
        t_hbm = alg_bytes / (peaks["hbm_gbs"] * 1e9)
        if t_tensor >= t_hbm:
            roofline = {"bound": "tensor", "kernel": kernel, "achieved": achieved, "peak": peak_tf,
                        "unit": "TFLOP/s", "frac": achieved / peak_tf,
                        "frac_of_burst": achieved / peaks["tflops_burst"],
                        "frac_of_sustained": achieved / peaks["tflops"],
                        "peak_source": peaks["source"] + (" (bf16_tflops: timed region < 1 s at >= 95 % of the maximum "
                                                          "SM clock, burst)" if burst else
                                                          " (bf16_tflops_sustained: timed region >= 1 s, or the SM clock "
                                                          "under load below 95 % of its maximum - power-capped)")}
        else:       # few classes or materialised logits: the kernel is HBM-bound (C = 80 is 35 FLOP/B)
            gbs = alg_bytes / (stages["similarity"] * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": kernel, "achieved": gbs, "peak": peaks["hbm_gbs"],
                        "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "peak_source": peaks["source"] + " (hbm_gbs)",
                        "tensor_TFLOPs": achieved, "tensor_frac_of_sustained": achieved / peaks["tflops"]}
        # north_star quotes nominal figures too: 2.25 PFLOP/s dense bf16, ~8 TB/s HBM3e
        roofline["frac_of_nominal"] = roofline["achieved"] / (2250.0 if roofline["bound"] == "tensor" else 8000.0)
        plain = not proj and not args.per_image_text and args.logits == "none"
        roofline.update({"traffic": ncu_traffic("sim_fused" if fused else "sim_gemm", batch) if plain else None,
                         "traffic_source": f"profiles/{NCU_CAPTURE} (ncu --set full, bytes per launch)",
                         "algorithmic_bytes": alg_bytes, "algorithmic_flops": flops,
                         "ms_per_launch": stages["similarity"],
                         "ms_per_launch_source": "CUDA events around the kernel in the per-stage pass (same steps, "
                                                 "run right after the timed region)"})
        k1_bytes = batch * anchors * (EMBED_DIM * 4 + EMBED_DIM * 2 + 4)
        k3_bytes = batch * anchors * (68 * 4 + 4 + 16) + batch * ((anchors + 31) // 32) * 4
        # K4: pass-mask words + per candidate (score, box, class) in + per kept row (box, score, class, anchor, keep) out
        k4_bytes = batch * (((anchors + 31) // 32) * 4 + cand * 24 + kept * 32 + 8)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16": "bf16", "fp16": "fp16", "fp32": "bf16x3 (fp32-accurate hi/lo split)"}[args.precision],
            "data": "synthetic",
            "config": static_config(args, n_gpus),
            "observed": {"path": timed_path, "timed_through": "ovdet_head_step (one C call, PDL)" if single_call
                         else "per-stage C calls", "mean_candidates_per_image": cand, "mean_kept_per_image": kept,
                         "images_at_max_det": overflow,
                         "ms_per_step_per_stage_pass": staged_ms / args.steps},
            "parity": parity,
            "e2e": e2e,
            "gpu_launches": launches * args.steps,
            "roofline": roofline,
            "stages_ms": stages,
            "stage_rooflines": {
                "l2norm_hbm_frac": (None if (fused or proj) else
                                    k1_bytes / (stages["l2norm"] * 1e-3) / 1e9 / peaks["hbm_gbs"]),
                "similarity_hbm_frac_fp32_input": batch * anchors * EMBED_DIM * 4 / (stages["similarity"] * 1e-3)
                                                  / 1e9 / peaks["hbm_gbs"],
                "decode_hbm_frac": k3_bytes / (stages["decode"] * 1e-3) / 1e9 / peaks["hbm_gbs"],
                "decode_hbm_frac_of_nominal_8TBps": k3_bytes / (stages["decode"] * 1e-3) / 1e9 / 8000.0,
                # K4 is latency/compute-shaped (one CTA per image, serial greedy resolve): its HBM fraction is
                # tiny by construction; IoU pairs resolved per second (candidates^2 / 2 per image) is the honest figure
                "nms_hbm_frac": k4_bytes / (stages["nms"] * 1e-3) / 1e9 / peaks["hbm_gbs"],
                "nms_iou_pairs_per_s": batch * cand * cand / 2.0 / (stages["nms"] * 1e-3),
                "nms_images_per_s": batch / (stages["nms"] * 1e-3)},
            "latency_ms_p50_batch1": p50_graph,
            "latency_ms_p50_batch1_eager_python": p50,
            "latency_note": "batch-1 K1..K4 step, CUDA events; graph = HeadPipeline.replay() of the captured "
                            "step, eager = HeadPipeline.run (one C call per step)",
            "clocks": clocks,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_vocab_parallel(args):
    """`--mode vocab-parallel` under torchrun (N >= 2): every rank sees the whole batch and owns a class
    shard; the per-anchor (max, argmax) is reduced inside the similarity kernel over NVLink peer
    mappings (ovdet.vocab_parallel).  Rank 0 prints one JSON line: parity against
    one GPU holding the whole vocabulary (scores / classes / kept anchors byte-equal), ms per step of
    the sharded step (eager single C call and CUDA-graph replay), of the same step with an NCCL
    all-reduce instead of the in-kernel exchange, and of one GPU with the whole vocabulary."""
    import torch
    import torch.distributed as dist
    from ovdet import shard, synth
    from ovdet import vocab_parallel as vp
    from ovdet.pipeline import HeadConfig, HeadPipeline
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world < 2:
        raise SystemExit("bench.py --mode vocab-parallel needs torchrun with at least 2 ranks")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    batch, classes = args.batch, NUM_CLASSES
    shapes = [(IMAGE_SIZE // s, IMAGE_SIZE // s) for s in STRIDES]
    cfg = HeadConfig(precision="bf16", max_det=args.max_det)
    inp = synth.make_inputs(batch=batch, image_size=IMAGE_SIZE, num_classes=classes, embed_dim=EMBED_DIM,
                            device=dev, seed=77)
    for t in inp.obj_embeds + inp.box_preds + [inp.text]:
        dist.broadcast(t, 0)                                    # the batch is replicated, byte-identical
    full = HeadPipeline(batch, shapes, classes, cfg, device=dev)
    full.set_vocabulary(inp.text)
    r = full.run(inp.obj_embeds, inp.box_preds)
    torch.cuda.synchronize()
    want = [t.clone() for t in (full.scores, full.class_ids, r.count, r.anchor)]

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            fn()
        b.record()
        dist.barrier()
        torch.cuda.synchronize()
        return shard.max_over_ranks(a.elapsed_time(b) / args.steps, dev)

    out = {}
    ok_all = True
    for mode in ("fused", "allreduce"):
        head = vp.VocabParallelHead(batch, shapes, classes, cfg, device=dev, exchange=mode)
        head.set_vocabulary(inp.text)
        ok = True
        for _ in range(3):                                      # both key-array parities
            res = head.run(inp.obj_embeds, inp.box_preds)
            torch.cuda.synchronize()
            ok &= torch.equal(head.scores, want[0]) and torch.equal(head.class_ids, want[1])
            ok &= torch.equal(res.count, want[2])
            for b_, k in enumerate(want[2].tolist()):
                ok &= torch.equal(res.anchor[b_, :k], want[3][b_, :k])
        ms = timed(lambda: head.run(inp.obj_embeds, inp.box_preds))
        entry = {"parity": bool(ok), "ms_per_step": ms}
        if mode == "fused":
            head.capture(inp.obj_embeds, inp.box_preds)
            entry["graph_ms_per_step"] = timed(head.replay)
            torch.cuda.synchronize()
            entry["graph_parity"] = bool(torch.equal(head.scores, want[0]) and torch.equal(head.result.count, want[2]))
            entry["timed_out"] = head.timed_out()
            ok &= entry["graph_parity"] and not entry["timed_out"]
        flag = torch.tensor([int(not ok)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        entry["parity_all_ranks"] = not bool(flag.item())
        ok_all &= entry["parity_all_ranks"]
        out[mode] = entry
        head.close()
    one = timed(lambda: full.run(inp.obj_embeds, inp.box_preds))
    if rank == 0:
        ms = out["fused"]["ms_per_step"]
        line = {"metric": "images/sec, vocabulary-parallel head + post-process (prompts sharded over the GPUs)",
                "value": batch / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": f"batch {batch} (replicated) @ {IMAGE_SIZE}x{IMAGE_SIZE}, {classes} prompts sharded "
                                       f"over {world} GPUs, in-kernel max/argmax exchange over NVLink peer memory",
                           "mode": "vocab-parallel", "classes": classes, "batch": batch, "max_det": args.max_det},
                "parity": {"vs": "one GPU holding the whole vocabulary: scores, class ids, kept anchors byte-equal, "
                                 "3 steps + graph replay, all ranks", "ok": bool(ok_all)},
                "exchange": out, "one_gpu_full_vocabulary_ms_per_step": one,
                "speedup_vs_one_gpu": one / ms, "gpu_launches": 5 * args.steps}
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()
    if not ok_all:
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="images per GPU per step (default 256; 16 in "
                    "--mode vocab-parallel, where the batch is replicated)")
    ap.add_argument("--e2e-chunk", type=int, default=32)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--per-image-text", action="store_true",
                    help="text embeddings [B, C, 512] with the neck's strides (the reference's real forward, "
                         "repvl_pan.py:173-182) instead of one shared vocabulary; K1b runs over B*C rows per step")
    ap.add_argument("--mode", default="batch", choices=["batch", "vocab-parallel"],
                    help="vocab-parallel (N > 1): the prompts sharded over the GPUs with the max/argmax exchange fused "
                         "into the similarity kernel; prints parity against one GPU holding the whole vocabulary + ms")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fused", action="store_true", help="two-kernel K1 -> K2 path instead of the fused kernel")
    ap.add_argument("--profile", action="store_true",
                    help="device-resident loop only (for ncu): no e2e, latency or CPU legs")
    ap.add_argument("--input-dtype", default="fp32", choices=["fp32", "bf16"],
                    help="dtype of the region embeddings handed to the path: fp32 (the reference's convolutions; "
                         "the metric's configuration) or bf16 (the head ran under autocast)")
    ap.add_argument("--max-det", type=int, default=MAX_DET,
                    help="rows of the per-image output (the reference has no cap; the line reports how many "
                         "images reached it - 0 at the default configuration)")
    ap.add_argument("--projected", action="store_true",
                    help="SURVEY 8f-2: start the step at the hidden features and fold the head's 1x1 projection "
                         "into the similarity (both arms)")
    ap.add_argument("--logits", default="none", choices=["none", "bf16", "fp32"],
                    help="also materialise the [B, A, C] logits (the reference's compute_similarity output); "
                         "default: class max/argmax fused in the GEMM epilogue only")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16", "fp32"],
                    help="bf16 = the metric's configuration; fp16 = one pass with fp16 operands (|dlogit| <~ 1e-4, same "
                         "tensor rate); fp32 = three-pass hi/lo recipe (BASELINE configs[1])")
    ap.add_argument("--image-size", type=int, default=IMAGE_SIZE,
                    help="default 640 (the metric's configuration); 1280 = BASELINE configs[3]")
    ap.add_argument("--classes", type=int, default=NUM_CLASSES,
                    help="default 1203 (the metric's configuration); 4800 = BASELINE configs[4]")
    args = ap.parse_args()
    globals().update(IMAGE_SIZE=args.image_size, NUM_CLASSES=args.classes)
    if args.batch is None:
        args.batch = 16 if args.mode == "vocab-parallel" else BATCH_PER_GPU
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "vocab-parallel":
        run_vocab_parallel(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
