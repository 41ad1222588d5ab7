"""CPU oracle for the open-vocabulary head + post-processing hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  Nothing under the product package imports it, and the
product never falls back to it.

It is a restatement (torch-CPU / numpy, float32) of the reference's algorithm for the
path; every function names the reference lines it follows (paths relative to
``/root/reference/yolo_clip_detector``).  The reference ships no tests or golden vectors
of its own, so the restatement is pinned by fixtures generated from
the live reference in the build container: ``oracle/make_golden.py`` imports the real
reference modules, runs them on seeded inputs and stores inputs+outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` replays them through this file.

Third-party arithmetic used exactly as the reference uses it: torch (F.normalize, matmul,
softmax, exp, max) and numpy (argsort, maximum/minimum, clip) at the installed versions
(torch 2.11.0, numpy 2.3.5); the reference's requirements.txt leaves both unpinned.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# Inference defaults of the reference: config/default_config.py:86-95, model/yolo_clip.py:39,
# model/heads/text_contrastive.py:44-45.
CONF_THRESHOLD = 0.25
IOU_THRESHOLD = 0.45
REG_MAX = 16
STRIDES = (8, 16, 32)
EMBED_DIM = 512


# --------------------------------------------------------------------------------------
# Stage 1+2: L2 normalisation and region x text similarity
# --------------------------------------------------------------------------------------
def compute_similarity(obj_embed: torch.Tensor, text_embed: torch.Tensor,
                       cls_alpha: float = 1.0, cls_beta: float = 0.0) -> torch.Tensor:
    """model/heads/text_contrastive.py:119-153.

    obj_embed [B,D,H,W], text_embed [B,C,D] -> logical [B,C,H,W] whose memory is [B,HW,C].
    """
    b, d, h, w = obj_embed.shape
    regions = obj_embed.permute(0, 2, 3, 1).reshape(b, h * w, d)          # :134
    regions = F.normalize(regions, p=2, dim=-1)                            # :137
    text = F.normalize(text_embed, p=2, dim=-1)                            # :138
    sim = torch.matmul(regions, text.transpose(1, 2))                      # :144
    sim = cls_alpha * sim + cls_beta                                       # :147
    c = text_embed.shape[1]
    return sim.transpose(1, 2).reshape(b, c, h, w)                         # :150-151


def class_max_concat(similarities: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """model/yolo_clip.py:198-206: per-level max/argmax over classes, flatten, concat."""
    score_parts, id_parts = [], []
    for sim in similarities:
        s, i = sim.max(dim=1)
        score_parts.append(s.flatten(1))
        id_parts.append(i.flatten(1))
    return torch.cat(score_parts, dim=1), torch.cat(id_parts, dim=1)


# --------------------------------------------------------------------------------------
# Stage 3a: DFL decode
# --------------------------------------------------------------------------------------
def create_grid(batch: int, height: int, width: int, stride: int) -> torch.Tensor:
    """model/heads/box_head.py:115-148: int64 [B,H,W,3] = (column, row, stride)."""
    rows, cols = torch.meshgrid(torch.arange(height), torch.arange(width), indexing="ij")
    cell = torch.stack([cols, rows, torch.ones_like(cols) * stride], dim=-1)
    return cell.unsqueeze(0).expand(batch, -1, -1, -1)


def decode_boxes(box_preds: Sequence[torch.Tensor], grids: Sequence[torch.Tensor],
                 strides: Sequence[int] = STRIDES) -> torch.Tensor:
    """model/heads/box_head.py:150-218.

    Per level: [B,4*R,H,W] -> softmax over the R bins of each of x,y,w,h -> expectation with
    weights 0..R-1 -> centre = (cell + e_xy) * stride, size = exp(e_wh) * stride ->
    xyxy = centre -/+ size/2; anchors ordered row-major inside a level, levels concatenated.
    """
    out = []
    for level, (pred, grid) in enumerate(zip(box_preds, grids)):
        b, ch, h, w = pred.shape
        stride = strides[level]
        bins = ch // 4
        dist = pred.reshape(b, 4, bins, h, w).softmax(dim=2)                # :182-185
        weights = torch.arange(bins).float()                                # :188
        expect = (dist * weights.view(1, 1, -1, 1, 1)).sum(dim=2)           # :192
        expect = expect.permute(0, 2, 3, 1)                                 # :195
        centre = (grid[..., :2] + expect[..., :2]) * stride                 # :203
        size = torch.exp(expect[..., 2:]) * stride                          # :205
        xyxy = torch.cat([centre - size / 2, centre + size / 2], dim=-1)    # :208-211
        out.append(xyxy.reshape(b, h * w, 4))                               # :214
    return torch.cat(out, dim=1)                                            # :218


def head_tail(obj_embeds: Sequence[torch.Tensor], text_embed: torch.Tensor,
              box_preds: Sequence[torch.Tensor], strides: Sequence[int] = STRIDES,
              cls_alpha: float = 1.0, cls_beta: float = 0.0) -> Dict[str, torch.Tensor]:
    """The tail of YOLOCLIP.forward, model/yolo_clip.py:173-223, from the conv outputs on."""
    sims = [compute_similarity(e, text_embed, cls_alpha, cls_beta) for e in obj_embeds]
    grids = [create_grid(p.shape[0], p.shape[2], p.shape[3], s) for p, s in zip(box_preds, strides)]
    boxes = decode_boxes(box_preds, grids, strides)
    scores, class_ids = class_max_concat(sims)
    return {"boxes": boxes, "scores": scores, "class_ids": class_ids, "similarities": sims}


# --------------------------------------------------------------------------------------
# Stage 3b+4: threshold, rescale/clip, greedy NMS (host numpy in the reference)
# --------------------------------------------------------------------------------------
def compute_iou(box: np.ndarray, boxes: np.ndarray) -> np.ndarray:
    """inference/detector.py:258-287 (float32 throughout, +1e-7 on the union)."""
    ix1 = np.maximum(box[0], boxes[:, 0])
    iy1 = np.maximum(box[1], boxes[:, 1])
    ix2 = np.minimum(box[2], boxes[:, 2])
    iy2 = np.minimum(box[3], boxes[:, 3])
    inter = np.maximum(0, ix2 - ix1) * np.maximum(0, iy2 - iy1)
    area_one = (box[2] - box[0]) * (box[3] - box[1])
    area_many = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    union = area_one + area_many - inter
    return inter / (union + 1e-7)


def nms(boxes: np.ndarray, scores: np.ndarray, iou_threshold: float,
        stable_ties: bool = False) -> List[int]:
    """inference/detector.py:225-256: greedy, class-agnostic, keeps iou <= threshold.

    ``stable_ties`` replaces ``np.argsort(scores)[::-1]`` by the total order
    (score desc, index desc) == ``np.argsort(scores, kind='stable')[::-1]``; identical to the
    reference whenever scores are pairwise distinct (SURVEY.md section 8a tie rule).
    """
    order = (np.argsort(scores, kind="stable") if stable_ties else np.argsort(scores))[::-1]
    keep: List[int] = []
    while order.size > 0:
        head = order[0]
        keep.append(head)
        rest = order[1:]
        ious = compute_iou(boxes[head], boxes[rest])
        order = rest[ious <= iou_threshold]
    return keep


def threshold_rescale_clip(boxes: np.ndarray, scores: np.ndarray, class_ids: np.ndarray,
                           orig_size: Tuple[int, int], scale_factor: float,
                           conf_threshold: float = CONF_THRESHOLD):
    """inference/detector.py:184-202 on one image's arrays; returns the survivor arrays and
    the anchor index of each survivor."""
    passed = scores > conf_threshold
    anchor_idx = np.nonzero(passed)[0]
    boxes = boxes[passed]
    scores = scores[passed]
    class_ids = class_ids[passed]
    orig_h, orig_w = orig_size
    for col in range(4):
        boxes[:, col] = boxes[:, col] / scale_factor
    boxes[:, 0] = np.clip(boxes[:, 0], 0, orig_w)
    boxes[:, 1] = np.clip(boxes[:, 1], 0, orig_h)
    boxes[:, 2] = np.clip(boxes[:, 2], 0, orig_w)
    boxes[:, 3] = np.clip(boxes[:, 3], 0, orig_h)
    return boxes, scores, class_ids, anchor_idx


def postprocess_image(boxes: np.ndarray, scores: np.ndarray, class_ids: np.ndarray,
                      orig_size: Tuple[int, int], scale_factor: float,
                      conf_threshold: float = CONF_THRESHOLD, iou_threshold: float = IOU_THRESHOLD,
                      class_names: Optional[Sequence[str]] = None, *, stable_ties: bool = True,
                      class_aware: bool = False, topk: Optional[int] = None,
                      activation: str = "none") -> Dict[str, object]:
    """inference/detector.py:163-223 for one image given as numpy arrays.

    The reference only handles image 0 (``outputs[k][0]``); batched callers loop this.
    ``class_aware`` / ``topk`` / ``activation='sigmoid'`` are north-star extensions with no
    reference counterpart (SURVEY.md section 8c rules v-vii) and default off.
    """
    boxes = np.array(boxes, dtype=np.float32, copy=True)
    scores = np.array(scores, dtype=np.float32, copy=True)
    class_ids = np.array(class_ids, copy=True)
    if activation == "sigmoid":
        scores = torch.sigmoid(torch.from_numpy(scores)).numpy()
    b, s, c, anchor_idx = threshold_rescale_clip(boxes, scores, class_ids, orig_size,
                                                 scale_factor, conf_threshold)
    if topk is not None and s.size > topk:
        # extension (vi): stable descending order truncated to K *before* NMS, then restored
        # to anchor order so that indices keep the reference meaning.
        top = np.sort((np.argsort(s, kind="stable")[::-1])[:topk])
        b, s, c, anchor_idx = b[top], s[top], c[top], anchor_idx[top]
    if class_aware:
        keep = nms_class_aware(b, s, c, iou_threshold)
    else:
        keep = nms(b, s, iou_threshold, stable_ties=stable_ties)
    keep_arr = np.asarray(keep, dtype=np.int64)
    kb, ks, kc = b[keep_arr], s[keep_arr], c[keep_arr]
    detections = []
    for i in range(len(kb)):                                                # :213-221
        cid = int(kc[i])
        detections.append({
            "box": kb[i].astype(int).tolist(),
            "score": float(ks[i]),
            "class_id": cid,
            "class_name": class_names[cid] if class_names is not None else f"Class {cid}",
        })
    return {"detections": detections, "keep": keep_arr, "boxes": kb, "scores": ks,
            "class_ids": kc, "anchor_idx": anchor_idx[keep_arr], "num_candidates": int(s.size)}


def nms_class_aware(boxes: np.ndarray, scores: np.ndarray, class_ids: np.ndarray,
                    iou_threshold: float) -> List[int]:
    """Extension (v): the reference ``_nms`` run independently per class id, results merged
    by (score desc, index desc)."""
    kept: List[int] = []
    for cid in np.unique(class_ids):
        members = np.nonzero(class_ids == cid)[0]
        local = nms(boxes[members], scores[members], iou_threshold, stable_ties=True)
        kept.extend(int(members[i]) for i in local)
    kept.sort(key=lambda i: (-float(scores[i]), -i))
    # float() of a float32 is exact, so the python sort reproduces the float32 order
    return kept


def postprocess_batch(outputs: Dict[str, torch.Tensor], orig_sizes, scale_factors,
                      **kw) -> List[Dict[str, object]]:
    """Loop ``postprocess_image`` over the batch (the reference handles image 0 only)."""
    boxes = outputs["boxes"].detach().cpu().numpy()
    scores = outputs["scores"].detach().cpu().numpy()
    class_ids = outputs["class_ids"].detach().cpu().numpy()
    results = []
    for i in range(boxes.shape[0]):
        size = orig_sizes[i] if isinstance(orig_sizes[0], (tuple, list)) else orig_sizes
        scale = scale_factors[i] if isinstance(scale_factors, (tuple, list, np.ndarray)) else scale_factors
        results.append(postprocess_image(boxes[i], scores[i], class_ids[i], tuple(size), float(scale), **kw))
    return results


# --------------------------------------------------------------------------------------
# "Next" rows: pre-processing, result records, vocabulary format
# --------------------------------------------------------------------------------------
def cv2_resize_linear_u8(src: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    """``cv2.resize(src, (dst_w, dst_h))`` for uint8 HWC input with the default INTER_LINEAR.

    Third-party algorithm: OpenCV (``opencv-python``; the reference's requirements.txt leaves
    it unpinned, 4.13.0 is installed) ``modules/imgproc/src/resize.cpp``: fixed-point bilinear
    interpolation with INTER_RESIZE_COEF_BITS = 11 (``HResizeLinear`` / ``VResizeLinear``), and
    ``cv::resize`` switching to the INTER_AREA fast path for an exact 2x decimation.  OpenCV is
    not part of /root/reference; this restatement is pinned against the installed cv2 by
    ``oracle/make_golden.py`` (300 random shapes, bit-exact) and by the committed fixtures.
    Call site in the reference: inference/detector.py:144.
    """
    sh, sw = src.shape[:2]
    if sw == 2 * dst_w and sh == 2 * dst_h:                       # is_area_fast, iscale == 2
        s = src.astype(np.int32)
        out = (s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2
        return out.astype(np.uint8)
    scale_x = 1.0 / (dst_w / sw)
    scale_y = 1.0 / (dst_h / sh)

    def coeffs(dn, sn, scale, clamp_frac):
        d = np.arange(dn, dtype=np.float64)
        f = ((d + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int32)
        f = (f - s.astype(np.float32)).astype(np.float32)
        if clamp_frac:                                            # columns: fx = 0 at the borders
            lo, hi = s < 0, s >= sn - 1
            f[lo], s[lo] = 0, 0
            f[hi], s[hi] = 0, sn - 1
        a0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int32)
        a1 = np.rint(f * np.float32(2048)).astype(np.int32)
        return np.clip(s, 0, sn - 1), np.clip(s + 1, 0, sn - 1), a0, a1

    sx0, sx1, ax0, ax1 = coeffs(dst_w, sw, scale_x, True)
    sy0, sy1, ay0, ay1 = coeffs(dst_h, sh, scale_y, False)        # rows: indices clamp, fy stays
    s = src.astype(np.int32)
    rows = s[:, sx0, :] * ax0[None, :, None] + s[:, sx1, :] * ax1[None, :, None]
    r0, r1 = rows[sy0], rows[sy1]
    out = (((ay0[:, None, None] * (r0 >> 4)) >> 16) + ((ay1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def letterbox_geometry(orig_h: int, orig_w: int, image_size: Tuple[int, int]) -> Tuple[float, int, int]:
    """inference/detector.py:139-142: scale factor (python float) and resized size."""
    input_h, input_w = image_size
    scale_factor = min(input_h / orig_h, input_w / orig_w)
    return scale_factor, int(orig_h * scale_factor), int(orig_w * scale_factor)


def preprocess_image(image: np.ndarray, image_size: Tuple[int, int] = (640, 640)):
    """inference/detector.py:119-161 for an RGB uint8 HWC array: returns the ``[1, 3, H, W]``
    float32 tensor, the original image and the scale factor."""
    orig_h, orig_w = image.shape[:2]
    input_h, input_w = image_size
    scale_factor, rh, rw = letterbox_geometry(orig_h, orig_w, image_size)
    resized = cv2_resize_linear_u8(image, rw, rh)                              # :144
    canvas = np.zeros((input_h, input_w, 3), dtype=np.uint8)                    # :147
    canvas[:rh, :rw, :] = resized                                               # :150
    canvas = canvas.astype(np.float32) / 255.0                                  # :153
    canvas = canvas.transpose(2, 0, 1)                                          # :156
    return torch.from_numpy(np.ascontiguousarray(canvas)).unsqueeze(0), image.copy(), scale_factor


def load_offline_vocabulary(path: str) -> Tuple[List[str], torch.Tensor]:
    """clip/vocab_builder.py:110-130 + model/yolo_clip.py:244-263: JSON ``{class_name:
    [floats]}`` -> class names in file order and the stacked ``[C, D]`` float32 matrix."""
    import json
    with open(path, "r") as f:
        vocab = json.load(f)
    names = list(vocab.keys())
    return names, torch.stack([torch.tensor(vocab[n]) for n in names])


def save_offline_vocabulary(path: str, class_names: Sequence[str], embeddings: torch.Tensor) -> None:
    """clip/vocab_builder.py:90-104: ``json.dump({name: embedding.tolist()})``."""
    import json
    save = {n: embeddings[i].cpu().numpy().tolist() for i, n in enumerate(class_names)}
    with open(path, "w") as f:
        json.dump(save, f)


def max_sigmoid_attention(y: torch.Tensor, projected_text: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """model/repvl_pan.py:80-95 for one bottleneck iteration: ``y [B, c, H, W]``,
    ``projected_text [B, C, c]`` -> (attended ``[B, c, H, W]``, max scores ``[B, HW]``)."""
    b, c, h, w = y.shape
    y_r = y.permute(0, 2, 3, 1).reshape(b, h * w, c)                           # :81
    scores = torch.matmul(y_r, projected_text.transpose(-1, -2))               # :85
    max_scores, _ = torch.max(scores, dim=-1, keepdim=True)                    # :88
    weights = torch.sigmoid(max_scores)                                        # :89
    attended = y_r * weights                                                   # :92
    return attended.reshape(b, h, w, c).permute(0, 3, 1, 2), max_scores.squeeze(-1)   # :95


def project_similarity_max(hidden: Sequence[torch.Tensor], weights: Sequence[torch.Tensor],
                           biases: Sequence[torch.Tensor], text_embed: torch.Tensor,
                           alpha: float = 1.0, beta: float = 0.0):
    """"Next" row f-2: the last layer of ``obj_embed_conv`` (``nn.Conv2d(hidden_dim, embed_dim, 1)``,
    model/heads/text_contrastive.py:67 applied at :112) followed by ``compute_similarity`` (:119-153)
    and the class max / level concat of model/yolo_clip.py:198-206, per level.  ``hidden[l]`` is
    ``[B, hidden_dim, H, W]`` (the output of the second ConvBlock), ``weights[l] [embed, hidden, 1, 1]``
    and ``biases[l] [embed]`` the level's 1x1 convolution.  Returns (scores [B, A], class_ids [B, A],
    obj_embeds list)."""
    embeds = [F.conv2d(h, w, b) for h, w, b in zip(hidden, weights, biases)]
    sims = [compute_similarity(e, text_embed, alpha, beta) for e in embeds]
    scores, ids = class_max_concat(sims)
    return scores, ids, embeds
