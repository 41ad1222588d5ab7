"""The head + post-processing hot path as one pre-planned sequence of launches.

``HeadPipeline`` owns every intermediate buffer (bf16 operands, inverse norms, scores, class
ids, boxes, pass mask, NMS workspace and outputs) for a fixed problem shape, so a step performs
no allocation and no host synchronisation.  bf16 precision: K1+K2 fused (one launch reads the
fp32 NCHW conv outputs of every level, normalises, multiplies, class max/argmax) -> K3 (decode +
threshold) -> K4 (gather/sort/top-k/NMS): 3 launches.  fp32 precision (3-pass hi/lo split) or
inputs the fused kernel's TMA cannot address: K1 (one launch per level) -> K2 -> K3 -> K4.  This
is what ``Detector.predict`` and ``bench.py`` run.  Reference call sequence it replaces:
model/yolo_clip.py:173-214 followed by inference/detector.py:163-223 for every image.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import torch

import ctypes
import os

from . import _cabi, ops

# NVTX ranges ("tracing / profiling"): one range per step and, on the per-stage
# launch path, one per kernel stage - what nsys / ncu --nvtx group by.  OVDET_NVTX=0 removes them.
_NVTX = os.environ.get("OVDET_NVTX", "1") != "0"
_STAGE_NAMES = {"l2norm": "ovdet.K1 l2norm", "similarity": "ovdet.K2 similarity", "decode": "ovdet.K3 decode",
                "nms": "ovdet.K4 nms"}


def _nvtx_push(name: str) -> None:
    if _NVTX:
        torch.cuda.nvtx.range_push(name)


def _nvtx_pop() -> None:
    if _NVTX:
        torch.cuda.nvtx.range_pop()


@dataclass(frozen=True)
class HeadConfig:
    """Defaults are the reference's inference defaults (config/default_config.py:86-95,
    model/yolo_clip.py:39, model/heads/text_contrastive.py:44-45); the last five fields are
    extensions that default to the reference behaviour."""
    embed_dim: int = 512
    reg_max: int = 16
    strides: Tuple[int, ...] = (8, 16, 32)
    cls_alpha: float = 1.0
    cls_beta: float = 0.0
    conf_threshold: float = 0.25
    iou_threshold: float = 0.45
    # "auto" (default): inside the fp32 bar of the reference's arithmetic (|dlogit| <= 1e-4) at the best speed
    #         the shape allows - "fp16" for embed_dim 512 with TMA-addressable levels (one tensor-core pass;
    #         at 80 prompts 0.22 ms against 0.35 ms for the three-pass mode, batch 64), else "fp32" (the fused
    #         streaming three-pass mode up to 128 classes, the two-kernel path above);
    # "bf16": one tensor-core pass, |dlogit| <~ 8e-3 (the BASELINE metric's configuration; opt-in);
    # "fp16": one pass at the same rate with fp16 operands and per-row power-of-two scaling, |dlogit| <~ 1e-4
    #         (embed_dim 512, fp32 activations, TMA-addressable levels);
    # "fp32": three bf16 passes over hi/lo operand halves, |dlogit| <~ 3e-5
    precision: str = "auto"
    logits_dtype: Optional[str] = None  # None: fused max only; "bf16" / "fp32": also materialise
    activation: str = "none"           # reference applies no activation to the scores
    topk: int = 0                      # 0 = every survivor goes to NMS (reference)
    class_aware: bool = False          # reference NMS is class-agnostic
    max_det: int = 0                   # 0 = capacity for every anchor (reference has no cap)
    fused: bool = True                 # bf16 only: K1+K2 in one kernel when the inputs allow it


def resolve_precision(config: "HeadConfig", level_shapes, num_classes: int, projected: bool = False) -> str:
    """What ``precision="auto"`` means for a problem shape (see ``HeadConfig``)."""
    if config.precision != "auto":
        return config.precision
    if projected:
        return "bf16"                       # the projected similarity is a single bf16 pass by construction
    d = config.embed_dim
    tma_ok = len(level_shapes) <= 4         # levels with unaligned rows are re-pitched first (ops.tma_addressable)
    if config.fused and tma_ok and d == 512 and config.logits_dtype != "fp32":
        return "fp16"                       # one pass of the CTA-pair kernel, ~1e-5 (max 7e-5)
    return "fp32"                           # three passes: fused streaming mode (<= 128 classes) or K1 -> K2


class HeadPipeline:
    def __init__(self, batch: int, level_shapes: Sequence[Tuple[int, int]], num_classes: int,
                 config: HeadConfig = HeadConfig(), device="cuda", per_image_text: bool = False,
                 projections: Optional[Sequence[Tuple[torch.Tensor, Optional[torch.Tensor]]]] = None):
        """``projections``: per level ``(weight, bias)`` of the head's last 1x1 convolution
        (text_contrastive.py:67).  When given, ``run`` takes the HIDDEN features of
        ``obj_embed_conv`` instead of its output and the projection is folded into the similarity
        (``ops.similarity_projected``, bf16 precision only)."""
        self.cfg = config
        self.batch = batch
        self.level_shapes = [tuple(s) for s in level_shapes]
        self.anchors = sum(h * w for h, w in self.level_shapes)
        self.num_classes = num_classes
        self.device = torch.device(device)
        self.per_image_text = per_image_text
        if config.precision not in ("auto", "bf16", "fp16", "fp32"):
            raise ValueError("ovdet: precision is 'auto', 'bf16', 'fp16' or 'fp32'")
        self._auto = config.precision == "auto"
        self._twin = None                  # "auto" + bf16 activations: the bf16-operand pipeline of the same shape
        self._vocab_text = None
        self._ctor = (batch, [tuple(s) for s in level_shapes], num_classes, device, per_image_text, projections)
        if config.precision == "auto":
            import dataclasses
            config = dataclasses.replace(config, precision=resolve_precision(
                config, level_shapes, num_classes, projected=projections is not None))
            self.cfg = config
        self.split = config.precision == "fp32"
        self.f16 = config.precision == "fp16"
        self.projections = None
        if projections is not None:
            if self.split or self.f16:
                raise ValueError("ovdet: the projected similarity is a bf16 path")
            self.projections = [(w.detach().to(device, torch.float32),
                                 None if b is None else b.detach().to(device, torch.float32))
                                for w, b in projections]
            assert len(self.projections) == len(self.level_shapes)
        self.level_ops = None
        d, a, dev = config.embed_dim, self.anchors, self.device
        kop = d * (2 if self.split else 1)
        self.want_fused = config.fused and not self.split and d % 64 == 0 and d <= 512
        if self.f16 and not (config.fused and d == 512):
            raise ValueError("ovdet: precision 'fp16' is a mode of the fused CTA-pair kernel (embed_dim 512)")
        # fp32 precision with a single class tile: the fused kernel's streaming three-pass mode
        self.want_fused_fp32 = (config.fused and self.split and d % 64 == 0 and d <= 512 and
                                (num_classes <= 128 or d <= 128))
        self.text_op3 = None
        self._kop = kop
        self.regions_op = None             # bf16 operand of the two-kernel path, allocated on first use
        self.inv_norm = torch.empty(batch, a, device=dev, dtype=torch.float32)
        self.text_op = torch.empty(batch if per_image_text else 1, num_classes, kop, device=dev,
                                   dtype=torch.float16 if self.f16 else torch.bfloat16)
        self.scores = torch.empty(batch, a, device=dev, dtype=torch.float32)
        self.class_ids = torch.empty(batch, a, device=dev, dtype=torch.int32)
        self.logits = None
        if config.logits_dtype is not None:
            dt = {"bf16": torch.bfloat16, "fp32": torch.float32}[config.logits_dtype]
            # rows padded to 16 bytes so that the kernel can use 16-byte stores; `logits` is the
            # [B, A, C] view of the padded buffer (same shape as the reference's, wider row pitch)
            per16 = 16 // torch.empty(0, dtype=dt).element_size()
            ldc = (num_classes + per16 - 1) // per16 * per16
            self.logits = torch.empty(batch, a, ldc, device=dev, dtype=dt)[..., :num_classes]
        self.boxes = torch.empty(batch, a, 4, device=dev, dtype=torch.float32)
        self.scores_act = (torch.empty(batch, a, device=dev, dtype=torch.float32)
                           if config.activation == "sigmoid" else None)
        self.pass_mask = torch.empty(batch, (a + 31) // 32, device=dev, dtype=torch.int32)
        self.max_det = config.max_det if config.max_det > 0 else a
        md = self.max_det
        self.result = ops.NmsResult(
            boxes=torch.zeros(batch, md, 4, device=dev, dtype=torch.float32),
            scores=torch.zeros(batch, md, device=dev, dtype=torch.float32),
            classes=torch.zeros(batch, md, device=dev, dtype=torch.int32),
            anchor=torch.zeros(batch, md, device=dev, dtype=torch.int32),
            keep=torch.zeros(batch, md, device=dev, dtype=torch.int32),
            count=torch.zeros(batch, device=dev, dtype=torch.int32),
            candidates=torch.zeros(batch, device=dev, dtype=torch.int32))
        self.workspace = torch.empty(max(16, ops.nms_workspace_bytes(batch, a)), device=dev,
                                     dtype=torch.uint8)
        self.scale = torch.ones(batch, device=dev, dtype=torch.float32)
        self.clip_wh = torch.zeros(batch, 2, device=dev, dtype=torch.float32)
        self.use_geometry = False
        self._vocab_ready = False
        self.last_path = None              # "fused" | "split": what the previous run() launched
        self.last_single_call = False      # the previous run() was ONE C call (ovdet_head_step)
        self._step_args = None             # ovdet_head_step_args, filled on first use
        self._parallel_decode = False      # set by capture(): decode forked beside the similarity kernel
        self._pad_bufs = []                # re-pitched copies of levels TMA cannot address as they are (odd H*W)
        self._side = None
        self._fork = None

    # -- one-off / per-call host parameters -------------------------------------------------
    def set_vocabulary(self, text: torch.Tensor) -> None:
        """Normalise a shared ``[C, D]`` vocabulary once (the reference re-normalises it three
        times per forward, text_contrastive.py:138)."""
        assert not self.per_image_text
        self._vocab_text = text
        if self._twin is not None:
            self._twin.set_vocabulary(text)
        if self.projections is not None:
            self.level_ops = [ops.project_vocabulary(text, w, b) for w, b in self.projections]
        else:
            ops.l2norm_text(text, split=self._text_mode(), operand=self.text_op)
            if self.want_fused_fp32:
                self.text_op3 = ops.l2norm_text(text, split=3, operand=self.text_op3)
        self._vocab_ready = True

    def _text_mode(self):
        return "fp16" if self.f16 else self.split

    def set_geometry(self, orig_sizes: Sequence[Tuple[int, int]], scale_factors: Sequence[float]) -> None:
        """Per-image ``(orig_h, orig_w)`` and letterbox scale (detector.py:193-202).  The scale
        is rounded to float32 exactly as numpy's weak python-float promotion does."""
        import numpy as np
        scale = np.asarray([np.float32(s) for s in scale_factors], dtype=np.float32)
        wh = np.asarray([[float(w), float(h)] for (h, w) in orig_sizes], dtype=np.float32)
        self.scale.copy_(torch.from_numpy(scale))
        self.clip_wh.copy_(torch.from_numpy(wh))
        self.use_geometry = True

    def clear_geometry(self) -> None:
        """Back to "boxes stay in input-image pixels, no clipping" (no ``orig_sizes`` given)."""
        self.use_geometry = False

    def check_inputs(self, obj_embeds: Sequence[torch.Tensor], box_preds: Sequence[torch.Tensor],
                     text: Optional[torch.Tensor] = None) -> None:
        """The planned shapes are what the tensor maps, the decode kernel and the workspaces are
        sized for: anything else would read or write past an allocation, so it is an error here."""
        n = len(self.level_shapes)
        if len(obj_embeds) != n or len(box_preds) != n:
            raise ValueError(f"ovdet: expected {n} levels, got {len(obj_embeds)} / {len(box_preds)}")
        dim = self.projections[0][0].shape[1] if self.projections is not None else self.cfg.embed_dim
        ch = 4 * (self.cfg.reg_max + 1)
        for l, ((h, w), e, p) in enumerate(zip(self.level_shapes, obj_embeds, box_preds)):
            if tuple(e.shape) != (self.batch, dim, h, w):
                raise ValueError(f"ovdet: level {l} embeddings are {tuple(e.shape)}, the pipeline was planned for "
                                 f"{(self.batch, dim, h, w)}")
            if tuple(p.shape) != (self.batch, ch, h, w):
                raise ValueError(f"ovdet: level {l} box_preds are {tuple(p.shape)}, the pipeline was planned for "
                                 f"{(self.batch, ch, h, w)}")
            if e.device != self.device or p.device != self.device:
                raise ValueError(f"ovdet: level {l} tensors live on {e.device} / {p.device}, the pipeline on {self.device}")
        if text is not None:
            want_d = self.cfg.embed_dim
            ok = text.shape[-2:] == (self.num_classes, want_d) and (
                text.dim() == 2 or (text.dim() == 3 and text.shape[0] in (1, self.batch)))
            if not ok:
                raise ValueError(f"ovdet: text embeddings are {tuple(text.shape)}, expected [{self.num_classes}, {want_d}] "
                                 f"or [{self.batch}, {self.num_classes}, {want_d}]")
            if self.per_image_text and not (text.dim() == 3 and text.shape[0] == self.batch):
                raise ValueError("ovdet: this pipeline was planned for per-image text [B, C, D]")

    # -- the hot path -----------------------------------------------------------------------
    def run(self, obj_embeds: Sequence[torch.Tensor], box_preds: Sequence[torch.Tensor],
            text: Optional[torch.Tensor] = None, events: Optional[dict] = None) -> ops.NmsResult:
        """One pass of the hot path.  ``events`` (optional dict) receives a pair of CUDA events
        per stage, recorded on the launching stream, for per-kernel timing."""
        _nvtx_push("ovdet.head_step")
        try:
            return self._run(obj_embeds, box_preds, text, events)
        finally:
            _nvtx_pop()

    def _run(self, obj_embeds, box_preds, text, events) -> ops.NmsResult:
        cfg = self.cfg
        self.check_inputs(obj_embeds, box_preds, text)
        if self._auto and self.f16 and obj_embeds[0].dtype == torch.bfloat16:
            return self._run_bf16_twin(obj_embeds, box_preds, text, events)
        if self.want_fused or self.want_fused_fp32 or self.projections is not None:
            obj_embeds = ops.tma_addressable(obj_embeds, self._pad_bufs)

        def mark(name, begin):
            if begin:
                _nvtx_push(_STAGE_NAMES.get(name, name))
            else:
                _nvtx_pop()
            if events is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                if begin:
                    events[name] = [ev, None]
                else:
                    events[name][1] = ev

        self._fork = None
        self.last_single_call = False
        if self._parallel_decode:
            # fork: decode needs nothing from the similarity kernel once K4 owns the threshold
            main = torch.cuda.current_stream(self.device)
            start = torch.cuda.Event()
            start.record(main)
            with torch.cuda.stream(self._side):
                self._side.wait_event(start)
                ops.decode_filter(box_preds, cfg.strides, boxes=self.boxes)
                self._fork = torch.cuda.Event()
                self._fork.record(self._side)
        if self.projections is not None:
            return self._run_projected(obj_embeds, box_preds, text, mark)
        if (events is None and self.want_fused and self.logits is None and not self._parallel_decode
                and ops.fused_supported(obj_embeds) and self._single_call_ok(box_preds)
                and (not self.f16 or ops.fused_fp16_supported(obj_embeds))):
            return self._run_single_call(obj_embeds, box_preds, text)
        fused = self.want_fused and ops.fused_supported(obj_embeds)
        fused32 = self.want_fused_fp32 and ops.fused_supported(obj_embeds)
        if self.f16 and not (fused and ops.fused_fp16_supported(obj_embeds)):
            raise ValueError("ovdet: precision 'fp16' (also what 'auto' resolves to at embed_dim 512) needs fp32 "
                             "activations; for bf16 activations (heads under autocast) construct the pipeline with "
                             "precision='bf16', for other layouts with precision='fp32'")
        self.last_path = "fused" if fused else ("fused_fp32" if fused32 else "split")
        mark("l2norm", True)
        if not fused and not fused32:
            if self.regions_op is None:
                self.regions_op = torch.empty(self.batch, self.anchors, self._kop, device=self.device,
                                              dtype=torch.bfloat16)
            ops.l2norm_regions(obj_embeds, split=self.split, operand=self.regions_op, inv_norm=self.inv_norm)
        if self.per_image_text:
            if fused32:         # one K1b launch straight into the [hi | lo | hi] operand, no ATen op in the step
                self.text_op3 = ops.l2norm_text(text, split=3, operand=self.text_op3)
            else:
                ops.l2norm_text(text, split=self._text_mode(), operand=self.text_op)
        elif text is not None:
            self.set_vocabulary(text)
        elif not self._vocab_ready:
            raise RuntimeError("ovdet: no vocabulary set (call set_vocabulary or pass text)")
        mark("l2norm", False)
        mark("similarity", True)
        if fused32:
            ops.similarity_fused(obj_embeds, self.text_op3, cfg.cls_alpha, cfg.cls_beta, logits_dtype=None,
                                 logits=self.logits, want_max=True, row_max=self.scores,
                                 row_arg=self.class_ids, inv_norm=self.inv_norm, fp32=True)
        elif fused:
            ops.similarity_fused(obj_embeds, self.text_op, cfg.cls_alpha, cfg.cls_beta, logits_dtype=None,
                                 logits=self.logits, want_max=True, row_max=self.scores,
                                 row_arg=self.class_ids, inv_norm=self.inv_norm)
        else:
            ops.similarity(self.regions_op, self.text_op, self.inv_norm, cfg.embed_dim, cfg.cls_alpha,
                           cfg.cls_beta, split=self.split, logits_dtype=None, logits=self.logits,
                           want_max=True, row_max=self.scores, row_arg=self.class_ids)
        mark("similarity", False)
        return self._decode_and_nms(box_preds, mark)

    def _run_bf16_twin(self, obj_embeds, box_preds, text, events) -> ops.NmsResult:
        """``precision="auto"`` resolved to the fp16 tier, but the activations arrive as bf16 (the heads ran
        under autocast): they are already rounded, the bf16-operand kernel multiplies them exactly and only
        the text rows are rounded to bf16.  A second pipeline of the same shape with ``precision="bf16"``
        takes the call; this one's result attributes point at its buffers afterwards."""
        if self._twin is None:
            import dataclasses
            batch, shapes, classes, device, per_image, projections = self._ctor
            self._twin = HeadPipeline(batch, shapes, classes, dataclasses.replace(self.cfg, precision="bf16"),
                                      device=device, per_image_text=per_image, projections=projections)
            if self._vocab_text is not None and not per_image:
                self._twin.set_vocabulary(self._vocab_text)
        twin = self._twin
        twin.use_geometry = self.use_geometry
        if self.use_geometry:
            twin.scale.copy_(self.scale)
            twin.clip_wh.copy_(self.clip_wh)
        res = twin._run(obj_embeds, box_preds, text, events)
        self.scores, self.class_ids, self.boxes, self.inv_norm = twin.scores, twin.class_ids, twin.boxes, twin.inv_norm
        self.pass_mask, self.result, self.logits = twin.pass_mask, twin.result, twin.logits
        self.last_path, self.last_single_call = twin.last_path, twin.last_single_call
        return res

    # -- the bf16 step as ONE C call (ovdet_head_step): same kernels, one host round trip ----------
    def _single_call_ok(self, box_preds) -> bool:
        for p in box_preds:
            if p.dtype not in (torch.float32, torch.bfloat16) or p.stride(3) != 1 or p.stride(2) != p.shape[3] or \
                    p.stride(1) != p.shape[2] * p.shape[3]:
                return False
        return len(box_preds) <= 4

    def _run_single_call(self, obj_embeds, box_preds, text) -> ops.NmsResult:
        cfg = self.cfg
        if self.per_image_text:
            ops.l2norm_text(text, split=self._text_mode(), operand=self.text_op)
        elif text is not None:
            self.set_vocabulary(text)
        elif not self._vocab_ready:
            raise RuntimeError("ovdet: no vocabulary set (call set_vocabulary or pass text)")
        a = self._fill_step_args(obj_embeds, box_preds)
        self.last_path = "fused"
        self.last_single_call = True
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().ovdet_head_step(ctypes.byref(a),
                                                    torch.cuda.current_stream(self.device).cuda_stream),
                        "ovdet_head_step")
        return self.result

    def _fill_step_args(self, obj_embeds, box_preds) -> "_cabi.HeadStepArgs":
        """The ``ovdet_head_step_args`` of this pipeline (built once; per call only the input
        pointers, strides and dtypes are refreshed)."""
        cfg = self.cfg
        a = self._step_args
        if a is None:
            a = _cabi.HeadStepArgs()
            assert ctypes.sizeof(a) == _cabi.lib().ovdet_head_step_args_size()
            n = len(self.level_shapes)
            a.num_levels, a.bins = n, cfg.reg_max + 1
            a.batch, a.dim, a.classes = self.batch, cfg.embed_dim, self.num_classes
            for l, (h, w) in enumerate(self.level_shapes):
                a.heights[l], a.widths[l], a.strides[l] = h, w, int(cfg.strides[l])
            a.text_op = self.text_op.data_ptr()
            a.text_batched = int(self.per_image_text and self.batch > 1)
            a.text_fp16 = int(self.f16)
            a.activation = {"none": _cabi.ACT_NONE, "sigmoid": _cabi.ACT_SIGMOID}[cfg.activation]
            a.class_aware, a.topk = int(cfg.class_aware), int(cfg.topk)
            a.alpha, a.beta, a.conf, a.iou_thr = cfg.cls_alpha, cfg.cls_beta, cfg.conf_threshold, cfg.iou_threshold
            a.max_det = self.max_det
            a.scores, a.class_ids = self.scores.data_ptr(), self.class_ids.data_ptr()
            a.inv_norm, a.boxes = self.inv_norm.data_ptr(), self.boxes.data_ptr()
            a.scores_act = None if self.scores_act is None else self.scores_act.data_ptr()
            a.pass_mask = self.pass_mask.data_ptr()
            r = self.result
            a.out_boxes, a.out_scores, a.out_classes = r.boxes.data_ptr(), r.scores.data_ptr(), r.classes.data_ptr()
            a.out_anchor, a.out_keep = r.anchor.data_ptr(), r.keep.data_ptr()
            a.out_count, a.out_candidates = r.count.data_ptr(), r.candidates.data_ptr()
            a.workspace, a.workspace_bytes = self.workspace.data_ptr(), self.workspace.numel()
            # small launches (batch 1): zeroed scratch for the class split of the similarity kernel
            need = _cabi.lib().ovdet_similarity_split_workspace_bytes(self.batch, self.anchors)
            if need:
                self._sim_ws = torch.zeros(need, device=self.device, dtype=torch.uint8)
                a.sim_workspace, a.sim_workspace_bytes = self._sim_ws.data_ptr(), need
            self._step_args = a
        for l, (e, p) in enumerate(zip(obj_embeds, box_preds)):
            a.obj_embeds[l], a.box_preds[l] = e.data_ptr(), p.data_ptr()
            a.emb_stride_b[l], a.emb_stride_d[l], a.box_stride_b[l] = e.stride(0), e.stride(1), p.stride(0)
        a.embed_dtype = _cabi.OVDET_BF16 if obj_embeds[0].dtype == torch.bfloat16 else _cabi.OVDET_F32
        a.box_dtype = _cabi.OVDET_BF16 if box_preds[0].dtype == torch.bfloat16 else _cabi.OVDET_F32
        a.scale = self.scale.data_ptr() if self.use_geometry else None
        a.clip_wh = self.clip_wh.data_ptr() if self.use_geometry else None
        return a

    def _run_projected(self, hidden, box_preds, text, mark) -> ops.NmsResult:
        cfg = self.cfg
        self.last_path = "projected"
        mark("l2norm", True)
        if self.per_image_text:
            self.level_ops = [ops.project_vocabulary(text, w, b) for w, b in self.projections]
        elif text is not None:
            self.set_vocabulary(text)
        elif not self._vocab_ready:
            raise RuntimeError("ovdet: no vocabulary set (call set_vocabulary or pass text)")
        mark("l2norm", False)
        mark("similarity", True)
        ops.similarity_projected(hidden, self.level_ops, self.num_classes, cfg.cls_alpha, cfg.cls_beta,
                                 row_max=self.scores, row_arg=self.class_ids, inv_norm=self.inv_norm)
        mark("similarity", False)
        return self._decode_and_nms(box_preds, mark)

    def _decode_and_nms(self, box_preds, mark) -> ops.NmsResult:
        cfg = self.cfg
        fork = getattr(self, "_fork", None)
        if fork is not None:
            # latency path (captured graph): the boxes were decoded on the side stream beside the
            # similarity kernel; K4 evaluates the confidence threshold itself
            main = torch.cuda.current_stream(self.device)
            main.wait_event(fork)
            return ops.nms_batched(self.boxes, self.scores, self.class_ids, None,
                                   scale=self.scale if self.use_geometry else None,
                                   clip_wh=self.clip_wh if self.use_geometry else None,
                                   iou_thr=cfg.iou_threshold, class_aware=cfg.class_aware,
                                   topk=cfg.topk, max_det=self.max_det, out=self.result,
                                   workspace=self.workspace, conf=cfg.conf_threshold)
        mark("decode", True)
        ops.decode_filter(box_preds, cfg.strides, scores=self.scores, conf=cfg.conf_threshold,
                          activation=cfg.activation, boxes=self.boxes, scores_act=self.scores_act,
                          pass_mask=self.pass_mask)
        mark("decode", False)
        nms_scores = self.scores_act if self.scores_act is not None else self.scores
        mark("nms", True)
        res = ops.nms_batched(self.boxes, nms_scores, self.class_ids, self.pass_mask,
                              scale=self.scale if self.use_geometry else None,
                              clip_wh=self.clip_wh if self.use_geometry else None,
                              iou_thr=cfg.iou_threshold, class_aware=cfg.class_aware,
                              topk=cfg.topk, max_det=self.max_det, out=self.result,
                              workspace=self.workspace)
        mark("nms", False)
        return res

    # -- CUDA-graph replay (latency path) -----------------------------------------------------
    def capture(self, obj_embeds: Sequence[torch.Tensor], box_preds: Sequence[torch.Tensor],
                text: Optional[torch.Tensor] = None, parallel_decode: bool = False) -> None:
        """Capture one step (its 3-5 launches) into a CUDA graph bound to THESE input tensors:
        every ``replay()`` reads their storage again, so a serving loop writes the next image's
        conv outputs into the same tensors and replays.  Every intermediate and output buffer is
        owned by the pipeline, so the captured addresses stay valid.  At batch 1 the step is
        launch-bound from Python (host time per call > GPU time); the replay halves its p50."""
        self._graph_inputs = (list(obj_embeds), list(box_preds), text)      # keep the storage alive
        # parallel_decode: the decode kernel runs on a second graph branch, concurrently with the
        # similarity kernel (at batch 1 that kernel occupies 68 of the 148 SMs), and K4 applies the
        # confidence threshold itself (ovdet_nms_batched_conf).  Measured: no gain (batch-1 p50
        # 0.054 vs 0.053 ms - the fork/join costs what the overlap saves), so it is off by default.
        # Only without a score activation and for pyramids K4's resident path addresses.
        self._parallel_decode = (parallel_decode and self.cfg.activation == "none" and self.anchors <= 65536)
        if self._parallel_decode and self._side is None:
            self._side = torch.cuda.Stream(self.device)
        stream = torch.cuda.Stream(self.device)
        stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(stream):
            self.run(obj_embeds, box_preds, text)       # warm-up: lazy buffers, kernel attributes
            stream.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                self.run(obj_embeds, box_preds, text)
        torch.cuda.current_stream(self.device).wait_stream(stream)
        self._graph = graph
        self._parallel_decode = False       # eager run() keeps the sequential launch order

    def replay(self) -> ops.NmsResult:
        """Re-run the captured step on the current contents of the captured input tensors."""
        if getattr(self, "_graph", None) is None:
            raise RuntimeError("ovdet: HeadPipeline.replay() before capture()")
        self._graph.replay()
        return self.result

    def outputs(self) -> Dict[str, torch.Tensor]:
        """The reference's forward-dict view of the intermediate tensors (yolo_clip.py:216-223)."""
        return {"boxes": self.boxes, "scores": self.scores, "class_ids": self.class_ids.long()}
