"""Post-processing and the predict entry point behind the reference's detector API.

Mirrors ``YOLOCLIPDetector`` (inference/detector.py:14-393) for the part of it that is on the
hot path: ``postprocess_detections`` (:163-223), ``_nms`` (:225-256) and ``detect``'s
model-output -> detections step (:310-319).  Image loading, the backbone/neck forward and
drawing stay with the caller (out of scope).

Every method accepts the host buffers the reference's methods receive (numpy arrays / CPU
tensors are copied to the device, the kernels run there, results come back) as well as CUDA
tensors; nothing is computed on the CPU.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import ops
from .pipeline import HeadConfig, HeadPipeline

ArrayLike = Union[np.ndarray, torch.Tensor]


def _to_device(x: ArrayLike, device, dtype) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    return x.to(device=device, dtype=dtype, non_blocking=True).contiguous()


class Detector:
    """``conf_threshold`` / ``iou_threshold`` / ``image_size`` / ``class_names`` are the four
    attributes the reference's post-processing reads."""

    def __init__(self, class_names: Optional[Sequence[str]] = None, conf_threshold: float = 0.25,
                 iou_threshold: float = 0.45, image_size: Tuple[int, int] = (640, 640),
                 device: str = "cuda:0", config: Optional[HeadConfig] = None, feature_fn=None):
        self.class_names = list(class_names) if class_names is not None else None
        self.conf_threshold = conf_threshold
        self.iou_threshold = iou_threshold
        self.image_size = image_size
        self.device = torch.device(device)
        self.config = config or HeadConfig(conf_threshold=conf_threshold, iou_threshold=iou_threshold)
        self._pipelines: Dict[tuple, HeadPipeline] = {}
        self._vocabulary: Optional[torch.Tensor] = None
        self._host_state = None
        self.feature_fn = feature_fn

    # ---- reference: YOLOCLIPDetector._nms (detector.py:225-256) ---------------------------
    def _nms(self, boxes: ArrayLike, scores: ArrayLike, iou_threshold: float) -> List[int]:
        """Greedy class-agnostic NMS; returns indices into ``boxes`` in descending-score order."""
        n = int(boxes.shape[0])
        if n == 0:
            return []
        b = _to_device(boxes, self.device, torch.float32).reshape(1, n, 4)
        s = _to_device(scores, self.device, torch.float32).reshape(1, n)
        res = ops.nms_batched(b, s, iou_thr=iou_threshold)
        k = int(res.count[0].item())
        return res.anchor[0, :k].tolist()

    # ---- batched post-processing on device tensors ----------------------------------------
    def postprocess_batch(self, outputs: Dict[str, torch.Tensor],
                          orig_sizes: Sequence[Tuple[int, int]], scale_factors: Sequence[float],
                          class_aware: bool = False, topk: int = 0, activation: str = "none",
                          max_det: Optional[int] = None) -> ops.NmsResult:
        """detector.py:184-208 for every image of ``outputs`` (keys ``boxes`` [B,A,4], ``scores``
        [B,A], ``class_ids`` [B,A]).  ``orig_sizes`` are (height, width) pairs."""
        dev = self.device
        boxes = _to_device(outputs["boxes"], dev, torch.float32)
        scores = _to_device(outputs["scores"], dev, torch.float32)
        classes = _to_device(outputs["class_ids"], dev, torch.int32)
        batch = scores.shape[0]
        scale = torch.from_numpy(np.asarray([np.float32(s) for s in scale_factors], dtype=np.float32)).to(dev)
        wh = torch.tensor([[float(w), float(h)] for (h, w) in orig_sizes], dtype=torch.float32, device=dev)
        assert scale.numel() == batch and wh.shape[0] == batch
        if activation == "none" and scores.shape[1] <= 65536:
            # the reference's case: K4 evaluates `scores > conf` itself (detector.py:184), no mask tensor
            return ops.nms_batched(boxes, scores, classes, None, scale=scale, clip_wh=wh,
                                   iou_thr=self.iou_threshold, class_aware=class_aware, topk=topk,
                                   max_det=max_det, conf=self.conf_threshold)
        if activation == "sigmoid":
            scores = torch.sigmoid(scores)
        mask = _pack_mask(scores > self.conf_threshold)
        return ops.nms_batched(boxes, scores, classes, mask, scale=scale, clip_wh=wh,
                               iou_thr=self.iou_threshold, class_aware=class_aware, topk=topk,
                               max_det=max_det)

    # ---- reference: YOLOCLIPDetector.postprocess_detections (detector.py:163-223) ----------
    def postprocess_detections(self, outputs: Dict[str, ArrayLike], orig_size: Tuple[int, int],
                               scale_factor: float) -> List[Dict]:
        """Image 0 of the batch only, exactly like the reference; returns its list of dicts
        (``box`` int-truncated xyxy, ``score``, ``class_id``, ``class_name``)."""
        first = {k: outputs[k][0:1] for k in ("boxes", "scores", "class_ids")}
        res = self.postprocess_batch(first, [orig_size], [scale_factor])
        return self.to_records(res, 0)

    def to_records(self, res: ops.NmsResult, image: int) -> List[Dict]:
        """detector.py:213-221: the detection dicts of one image (``box`` truncated to int on
        the device, one D2H copy per field)."""
        k = int(res.count[image].item())
        boxes = ops.pack_boxes_i32(res.boxes, res.count)[image, :k].cpu().numpy()
        scores = res.scores[image, :k].cpu().numpy()
        classes = res.classes[image, :k].cpu().numpy()
        records = []
        for i in range(k):
            cid = int(classes[i])
            name = self.class_names[cid] if self.class_names is not None else f"Class {cid}"
            records.append({"box": boxes[i].tolist(), "score": float(scores[i]),
                            "class_id": cid, "class_name": name})
        return records

    # ---- reference: YOLOCLIPDetector.preprocess_image (detector.py:119-161) -----------------
    def preprocess_image(self, image: Union[str, ArrayLike]):
        """Letterbox one image: path or RGB uint8 ``[H, W, 3]`` array / tensor -> (``[1, 3, H, W]``
        fp32 tensor on the device, the original image, scale factor).  Decoding a file stays on
        the host (cv2, like the reference); resize / pad / normalise / transpose run in P1."""
        if isinstance(image, str):
            import cv2                                   # host image decode only (detector.py:133-134)
            image = cv2.cvtColor(cv2.imread(image), cv2.COLOR_BGR2RGB)
        orig = image
        dev_img = _to_device(image, self.device, torch.uint8)
        tensor, scales = ops.letterbox([dev_img], self.image_size)
        if isinstance(orig, np.ndarray):
            orig = orig.copy()
        return tensor, orig, scales[0]

    def preprocess_batch(self, images: Sequence[ArrayLike]):
        """``preprocess_image`` for a list of images of any sizes in one launch: returns the
        ``[N, 3, H, W]`` batch, the (height, width) of every original and the scale factors."""
        dev = [_to_device(im, self.device, torch.uint8) for im in images]
        tensor, scales = ops.letterbox(dev, self.image_size)
        return tensor, [(int(im.shape[0]), int(im.shape[1])) for im in dev], scales

    # ---- reference: YOLOCLIPDetector.detect (detector.py:289-325) ---------------------------
    def detect(self, image: Union[str, ArrayLike], text_prompts=None) -> List[Dict]:
        """preprocess -> model -> post-process for one image, like the reference.  The
        convolutional model (backbone, neck, head convolutions: out of scope here, SURVEY.md
        section 2) is the callable given as ``feature_fn``: it maps the ``[1, 3, H, W]`` tensor
        (and the prompts) to ``(obj_embeds, box_preds, text_embeddings)``."""
        if self.feature_fn is None:
            raise RuntimeError("ovdet: Detector.detect needs a feature_fn (the convolutional model)")
        tensor, orig, scale = self.preprocess_image(image)
        h, w = orig.shape[:2]
        with torch.no_grad():
            obj_embeds, box_preds, text = self.feature_fn(tensor, text_prompts)
        res = self.predict(obj_embeds, box_preds, text, [(h, w)], [scale])
        return self.to_records(res, 0)

    def load_offline_vocabulary(self, path: str) -> None:
        """model/yolo_clip.py:244-263: read the JSON vocabulary; class names come from it."""
        from .vocabulary import Vocabulary
        vocab = Vocabulary.load(path)
        self.class_names = vocab.class_names
        self.set_vocabulary(vocab.embeddings)

    # ---- predict: conv outputs -> detections (detect.py:121-125 -> detector.py:310-319) -----
    def pipeline_for(self, obj_embeds: Sequence[torch.Tensor], num_classes: int,
                     per_image_text: bool, projections=None) -> HeadPipeline:
        shapes = tuple((e.shape[2], e.shape[3]) for e in obj_embeds)
        proj_key = None if projections is None else tuple(w.data_ptr() for w, _ in projections)
        dim = self.config.embed_dim if projections is not None else int(obj_embeds[0].shape[1])
        key = (obj_embeds[0].shape[0], shapes, num_classes, per_image_text, obj_embeds[0].device, proj_key, dim)
        if key not in self._pipelines:
            config = self.config if dim == self.config.embed_dim else dataclasses.replace(self.config, embed_dim=dim)
            self._pipelines[key] = HeadPipeline(key[0], shapes, num_classes, config,
                                                device=obj_embeds[0].device,
                                                per_image_text=per_image_text, projections=projections)
        return self._pipelines[key]

    def set_vocabulary(self, text: torch.Tensor) -> None:
        """Offline vocabulary ``[C, D]`` (model/yolo_clip.py:225-263): kept on the device and
        normalised once per pipeline instead of once per level per forward."""
        self._vocabulary = text.to(self.device, torch.float32)
        for pipe in self._pipelines.values():
            if not pipe.per_image_text:
                pipe.set_vocabulary(self._vocabulary)

    def predict_host(self, obj_embeds: Sequence[torch.Tensor], box_preds: Sequence[torch.Tensor],
                     chunk: int = 32, projections=None,
                     orig_sizes: Optional[Sequence[Tuple[int, int]]] = None,
                     scale_factors: Optional[Sequence[float]] = None) -> Dict[str, torch.Tensor]:
        """``predict`` for HOST buffers (pinned CPU tensors, as a serving front-end holds them):
        the batch is cut into chunks; chunk i+1 is copied host->device on a copy stream while
        chunk i runs K1..K4 on the compute stream, and each chunk's detections are copied back
        as soon as its NMS has finished.  Returns pinned host tensors ``boxes [B,max_det,4]``,
        ``scores``, ``classes`` [B,max_det] and ``count`` [B]; needs ``set_vocabulary`` first.
        With ``projections`` (per level ``(weight, bias)`` of the head's last 1x1 convolution) the
        host buffers are the HIDDEN features and the projection is folded into the similarity:
        half the bytes cross PCIe."""
        if self._vocabulary is None:
            raise RuntimeError("ovdet: predict_host needs set_vocabulary() first")
        batch = obj_embeds[0].shape[0]
        chunk = min(chunk, batch)
        if batch % chunk:
            raise ValueError("ovdet: batch must be a multiple of the chunk size")
        key = (batch, chunk, tuple(tuple(e.shape[1:]) for e in obj_embeds), obj_embeds[0].dtype,
               None if projections is None else tuple(w.data_ptr() for w, _ in projections))
        st = self._host_state
        if st is None or st["key"] != key:
            dev = self.device
            shapes = tuple((e.shape[2], e.shape[3]) for e in obj_embeds)
            pipe = HeadPipeline(chunk, shapes, self._vocabulary.shape[0], self.config, device=dev,
                                projections=projections)
            pipe.set_vocabulary(self._vocabulary)
            md = pipe.max_det
            st = {
                "key": key, "pipe": pipe,
                "copy": torch.cuda.Stream(dev), "compute": torch.cuda.Stream(dev),
                "stage": [([torch.empty((chunk,) + tuple(e.shape[1:]), device=dev, dtype=e.dtype) for e in obj_embeds],
                           [torch.empty((chunk,) + tuple(p.shape[1:]), device=dev, dtype=p.dtype) for p in box_preds])
                          for _ in range(2)],
                "out": {"boxes": torch.empty(batch, md, 4).pin_memory(),
                        "scores": torch.empty(batch, md).pin_memory(),
                        "classes": torch.empty(batch, md, dtype=torch.int32).pin_memory(),
                        "count": torch.empty(batch, dtype=torch.int32).pin_memory()},
            }
            self._host_state = st
        pipe, out = st["pipe"], st["out"]
        copy_s, comp_s = st["copy"], st["compute"]
        cur = torch.cuda.current_stream(self.device)
        copy_s.wait_stream(cur)
        comp_s.wait_stream(cur)
        freed = [None, None]
        for i in range(batch // chunk):
            lo, hi = i * chunk, (i + 1) * chunk
            objs, boxes = st["stage"][i & 1]
            if orig_sizes is not None:
                geo = (tuple(orig_sizes[lo:hi]), tuple(scale_factors[lo:hi]) if scale_factors is not None else None)
                if geo != st.get("geo"):          # unchanged geometry (fixed-size serving) is not re-uploaded
                    with torch.cuda.stream(comp_s):
                        pipe.set_geometry(geo[0], geo[1] if geo[1] is not None else [1.0] * chunk)
                    st["geo"] = geo
            else:
                pipe.clear_geometry()
                st["geo"] = None
            with torch.cuda.stream(copy_s):
                if freed[i & 1] is not None:
                    copy_s.wait_event(freed[i & 1])
                for dst, src in zip(objs + boxes, list(obj_embeds) + list(box_preds)):
                    dst.copy_(src[lo:hi], non_blocking=True)
                copied = torch.cuda.Event()
                copied.record(copy_s)
            with torch.cuda.stream(comp_s):
                comp_s.wait_event(copied)
                res = pipe.run(objs, boxes)
                done = torch.cuda.Event()
                done.record(comp_s)
                freed[i & 1] = done
                out["boxes"][lo:hi].copy_(res.boxes, non_blocking=True)
                out["scores"][lo:hi].copy_(res.scores, non_blocking=True)
                out["classes"][lo:hi].copy_(res.classes, non_blocking=True)
                out["count"][lo:hi].copy_(res.count, non_blocking=True)
        cur.wait_stream(comp_s)
        comp_s.synchronize()
        return out

    def predict(self, obj_embeds: Sequence[torch.Tensor], box_preds: Sequence[torch.Tensor],
                text_embeddings: torch.Tensor, orig_sizes: Optional[Sequence[Tuple[int, int]]] = None,
                scale_factors: Optional[Sequence[float]] = None, projections=None) -> ops.NmsResult:
        """Per-level ``obj_embed [B,D,H,W]`` / ``box_preds [B,4R,H,W]`` (what the head
        convolutions emit) and text embeddings ``[C,D]`` or ``[B,C,D]`` -> per-image kept
        boxes / scores / classes, every image of the batch.  With ``projections`` (per level
        ``(weight, bias)`` of the head's last 1x1 convolution) ``obj_embeds`` are the HIDDEN features
        entering that convolution and it is folded into the similarity ("next" row f-2)."""
        per_image = not ops.shared_text(text_embeddings)
        pipe = self.pipeline_for(obj_embeds, text_embeddings.shape[-2], per_image, projections)
        if orig_sizes is not None:
            pipe.set_geometry(orig_sizes, scale_factors if scale_factors is not None
                              else [1.0] * len(orig_sizes))
        else:
            pipe.clear_geometry()       # pipelines are cached per shape: no geometry left over from an earlier call
        return pipe.run(obj_embeds, box_preds, text_embeddings)


class YOLOCLIPDetector(Detector):
    """``YOLOCLIPDetector`` of inference/detector.py:14-117, 289-325 with the reference's
    constructor arguments and ``detect(image, text_prompts=None) -> List[Dict]``; the call sequence
    of detect.py:92-125 runs against it unchanged.

    What stays outside (backbone, neck, CLIP, head convolutions are
    PyTorch/cuDNN code that is not on the path) is the convolutional model itself, so it is
    injected: ``model`` is a ready ``YOLOCLIP``-shaped module (attributes ``backbone``, ``neck``,
    ``contrastive_heads``, ``box_head``, ``offline_mode``, ``offline_vocabulary``,
    ``text_encoder``), or ``model_factory`` builds one from the reference's keyword arguments;
    with neither, the reference's own class is imported (``yolo_clip_detector`` on ``sys.path``).
    Everything from the letterbox to the detection records except that model runs in
    ``libovdet.so``: P1 letterbox -> [model] -> K1+K2 fused similarity -> K3 decode -> K4 NMS ->
    int-truncated records.  ``precision``: ``"auto"`` (default: every score within 1e-4 of the
    reference's fp32 arithmetic - the fp16 tensor-core tier at embed_dim 512, the three-pass recipe
    otherwise), ``"fp32"``, ``"fp16"`` or ``"bf16"`` (|dscore| <~ 8e-3)."""

    def __init__(self, model_path: Optional[str] = None, class_names: Optional[List[str]] = None,
                 vocab_path: Optional[str] = None, device=None, image_size: Tuple[int, int] = (640, 640),
                 conf_threshold: float = 0.25, iou_threshold: float = 0.45, backbone_variant: str = "n",
                 clip_model: str = "ViT-B/32", embed_dim: int = 512, *, model: Optional[torch.nn.Module] = None,
                 model_factory=None, precision: str = "auto", max_det: int = 0):
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("ovdet: YOLOCLIPDetector needs a CUDA device (no CPU fallback exists)")
            device = "cuda:0"
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("ovdet: YOLOCLIPDetector needs a CUDA device (no CPU fallback exists)")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        config = HeadConfig(embed_dim=embed_dim, conf_threshold=conf_threshold, iou_threshold=iou_threshold,
                            precision=precision, max_det=max_det)
        super().__init__(class_names=class_names, conf_threshold=conf_threshold, iou_threshold=iou_threshold,
                         image_size=image_size, device=device, config=config)
        if model is None:
            kwargs = dict(backbone_variant=backbone_variant, clip_model=clip_model, embed_dim=embed_dim,
                          num_classes=len(class_names) if class_names is not None else 80,
                          offline_mode=vocab_path is not None or class_names is not None)
            if model_factory is None:
                try:
                    from yolo_clip_detector.model.yolo_clip import YOLOCLIP as model_factory
                except ImportError as exc:
                    raise RuntimeError("ovdet: pass `model=` / `model_factory=` (the convolutional model is not "
                                       "part of this library) or put yolo_clip_detector on sys.path") from exc
            model = model_factory(**kwargs)
        self.model = model.to(self.device)
        if model_path is not None:
            self._load_model(model_path)
        self.model.eval()
        self.use_offline_vocab = False
        if vocab_path is not None:
            self.load_offline_vocabulary(vocab_path)
        elif class_names is not None and hasattr(self.model, "set_offline_vocabulary"):
            self.model.set_offline_vocabulary(class_names)        # needs the model's CLIP text encoder
            self.use_offline_vocab = True
        elif getattr(self.model, "offline_mode", False) and getattr(self.model, "offline_vocabulary", None) is not None:
            self.use_offline_vocab = True                         # an injected model that carries its vocabulary
        self.feature_fn = self._features

    def _load_model(self, model_path: str) -> None:
        """inference/detector.py:103-117: a raw state dict or one wrapped as ``model_state_dict``."""
        checkpoint = torch.load(model_path, map_location=self.device)
        state = checkpoint["model_state_dict"] if "model_state_dict" in checkpoint else checkpoint
        self.model.load_state_dict(state)

    def load_offline_vocabulary(self, path: str) -> None:
        """model/yolo_clip.py:244-263 through ``ovdet.vocabulary``: class names and the ``[C, D]`` matrix
        of the reference's JSON file; the model's own ``offline_vocabulary`` is set as well."""
        from .vocabulary import Vocabulary
        vocab = Vocabulary.load(path)
        if self.class_names is None:
            self.class_names = vocab.class_names
        self.model.offline_mode = True
        self.model.offline_vocabulary = vocab.embeddings.to(self.device, torch.float32)
        self.use_offline_vocab = True

    def _features(self, tensor: torch.Tensor, text_prompts=None):
        """model/yolo_clip.py:121-189 up to the convolution outputs: text embeddings, backbone, neck,
        the heads' embedding branch and the box head's convolutions."""
        from .heads import prompt_embeddings
        model = self.model
        if not self.use_offline_vocab and text_prompts is None:
            raise ValueError("Text prompts must be provided in online mode")
        text = prompt_embeddings(model, tensor.shape[0], None if self.use_offline_vocab else text_prompts)
        pan_features, text = model.neck(model.backbone(tensor), text)
        obj_embeds = [head.obj_embed_conv(feat) for feat, head in zip(pan_features, model.contrastive_heads)]
        box_preds = [conv(feat) for conv, feat in zip(model.box_head.box_convs, pan_features)]
        return obj_embeds, box_preds, text


def _pack_mask(passed: torch.Tensor) -> torch.Tensor:
    """bool [B, A] -> int32 bit mask [B, ceil(A/32)] (bit a%32 of word a/32)."""
    b, a = passed.shape
    words = (a + 31) // 32
    padded = torch.zeros(b, words * 32, dtype=torch.int64, device=passed.device)
    padded[:, :a] = passed
    weights = (1 << torch.arange(32, device=passed.device, dtype=torch.int64))
    packed = (padded.view(b, words, 32) * weights).sum(-1)
    packed = torch.where(packed >= 2 ** 31, packed - 2 ** 32, packed)
    return packed.to(torch.int32).contiguous()
