"""Generate the golden fixtures under tests/golden/ from the LIVE reference.

TEST INFRASTRUCTURE.  Run in the build container only (``/root/reference`` does not exist on
the GPU box):

    python oracle/make_golden.py

It imports the reference's own modules (``yolo_clip_detector.model.heads.*``,
``yolo_clip_detector.inference.detector``, ``yolo_clip_detector.model.yolo_clip`` with the
``clip`` stub from ``oracle/clip_stub.py``), runs them on seeded inputs and stores inputs and
outputs as small ``.npz`` files.  The reference has no tests or golden vectors of its own
(SURVEY.md section 4), so these files are what pins ``oracle/ref_port.py``.
"""
from __future__ import annotations

import logging
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("OVDET_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import clip_stub  # noqa: E402

clip_stub.install()
logging.disable(logging.CRITICAL)

from yolo_clip_detector.model.heads.text_contrastive import TextContrastiveHead  # noqa: E402
from yolo_clip_detector.model.heads.box_head import BoxHead  # noqa: E402
from yolo_clip_detector.inference.detector import YOLOCLIPDetector  # noqa: E402
from yolo_clip_detector.model.yolo_clip import YOLOCLIP  # noqa: E402

from ovdet import synth  # noqa: E402


def _bare_detector(conf=0.25, iou=0.45, image_size=(640, 640), class_names=None):
    det = YOLOCLIPDetector.__new__(YOLOCLIPDetector)      # no model / CLIP needed for post-process
    det.conf_threshold = conf
    det.iou_threshold = iou
    det.image_size = image_size
    det.class_names = class_names
    return det


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def similarity_cases():
    torch.manual_seed(11)
    for name, (b, d, h, w, c, alpha, beta, shared) in {
        "sim_batched_d512": (2, 512, 5, 4, 7, 1.0, 0.0, False),
        "sim_shared_affine_d64": (3, 64, 6, 6, 33, 1.3, -0.1, True),
    }.items():
        head = TextContrastiveHead(in_channels=8, embed_dim=d, cls_alpha=alpha, cls_beta=beta)
        obj = torch.randn(b, d, h, w) * 3.0
        if shared:
            text = (torch.randn(c, d) * 0.5).unsqueeze(0).expand(b, -1, -1)
        else:
            # the neck hands over a [B,C,D] tensor whose batch is NOT the outer memory dim
            text = torch.randn(c, b, d).transpose(0, 1)
        with torch.no_grad():
            sim = head.compute_similarity(obj, text)
        assert not sim.is_contiguous() and sim.shape == (b, c, h, w)
        assert sim.stride() == (c * h * w, 1, w * c, c)          # memory is [B,HW,C]
        save(name, obj=obj.numpy(), text=text.contiguous().numpy(), sim=sim.contiguous().numpy(),
             alpha=np.float64(alpha), beta=np.float64(beta), strides=np.array(sim.stride()))


def decode_case():
    torch.manual_seed(12)
    inp = synth.make_inputs(batch=2, image_size=64, num_classes=5, embed_dim=16, seed=5)
    head = BoxHead(in_channels=[8, 8, 8])
    grids = [head._create_grid(2, p.shape[2], p.shape[3], s, p.device)
             for p, s in zip(inp.box_preds, head.strides)]
    with torch.no_grad():
        boxes = head.decode_boxes(inp.box_preds, grids)
    noise = [torch.randn_like(p) * 2.0 for p in inp.box_preds]     # pure-noise logits too
    with torch.no_grad():
        boxes_noise = head.decode_boxes(noise, grids)
    save("decode_3level", p0=inp.box_preds[0].numpy(), p1=inp.box_preds[1].numpy(),
         p2=inp.box_preds[2].numpy(), boxes=boxes.numpy(),
         n0=noise[0].numpy(), n1=noise[1].numpy(), n2=noise[2].numpy(), boxes_noise=boxes_noise.numpy(),
         grid0=grids[0].contiguous().numpy())


def nms_cases():
    rng = np.random.default_rng(13)
    det = _bare_detector()
    cases = {}

    def rand_boxes(n, span, size):
        xy = rng.uniform(0, span, (n, 2)).astype(np.float32)
        wh = rng.uniform(1, size, (n, 2)).astype(np.float32)
        return np.concatenate([xy, xy + wh], axis=1).astype(np.float32)

    def distinct_scores(n):
        return rng.permutation(np.linspace(0.26, 0.99, n)).astype(np.float32)

    cases["uniform300"] = (rand_boxes(300, 600, 120), distinct_scores(300), 0.45)
    cases["dense1000"] = (rand_boxes(1000, 200, 150), distinct_scores(1000), 0.45)
    cases["loose_thr"] = (rand_boxes(257, 100, 90), distinct_scores(257), 0.7)
    cases["tight_thr"] = (rand_boxes(129, 100, 90), distinct_scores(129), 0.05)
    b = rand_boxes(64, 50, 40)
    b[10:20] = b[10]                       # exact duplicates
    b[30:34, 2:] = b[30:34, :2]            # zero-area boxes
    b[40, [0, 2]] = b[40, [2, 0]]          # inverted box (negative width)
    cases["degenerate64"] = (b, distinct_scores(64), 0.45)
    cases["single"] = (rand_boxes(1, 10, 5), distinct_scores(1), 0.45)
    cases["empty"] = (np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 0.45)
    out = {}
    for name, (boxes, scores, thr) in cases.items():
        keep = det._nms(boxes.copy(), scores.copy(), thr)
        out[name + "_boxes"] = boxes
        out[name + "_scores"] = scores
        out[name + "_thr"] = np.float64(thr)
        out[name + "_keep"] = np.asarray(keep, dtype=np.int64)
        if len(boxes) > 1:
            out[name + "_iou0"] = det._compute_iou(boxes[0], boxes[1:])
    # documented tie behaviour: equal scores -> higher index first
    tie_scores = np.array([.5, .7, .5, .7, .1], np.float32)
    out["tie_order"] = np.argsort(tie_scores)[::-1].copy()
    save("nms_cases", **out)


def postprocess_case():
    # synthetic head outputs for B=3 at 128x128 with enough survivors for NMS to matter
    from oracle import ref_port
    inp = synth.make_inputs(batch=3, image_size=128, num_classes=40, embed_dim=64, seed=21,
                            plant_frac=0.08)
    tail = ref_port.head_tail(inp.obj_embeds, inp.text_batched(), inp.box_preds)
    names = [f"thing{i}" for i in range(40)]
    out = {"boxes": tail["boxes"].numpy(), "scores": tail["scores"].numpy(),
           "class_ids": tail["class_ids"].numpy()}
    geometry = [((128, 128), 1.0), ((100, 160), 0.8), ((300, 200), 128 / 300)]
    for i, (orig, scale) in enumerate(geometry):
        det = _bare_detector(conf=0.25, iou=0.45, image_size=(128, 128), class_names=names)
        sliced = {k: torch.from_numpy(out[k][i:i + 1].copy()) for k in ("boxes", "scores", "class_ids")}
        dets = det.postprocess_detections(sliced, orig, scale)
        out[f"img{i}_orig"] = np.array(orig)
        out[f"img{i}_scale"] = np.float64(scale)
        out[f"img{i}_box"] = np.array([d["box"] for d in dets], dtype=np.int64).reshape(-1, 4)
        out[f"img{i}_score"] = np.array([d["score"] for d in dets], dtype=np.float64)
        out[f"img{i}_class"] = np.array([d["class_id"] for d in dets], dtype=np.int64)
        out[f"img{i}_name0"] = np.array(dets[0]["class_name"] if dets else "")
        print(f"  postprocess img{i}: {len(dets)} detections")
    save("postprocess_b3", **out)


def forward_tail_case():
    """Full reference YOLOCLIP.forward on a 64x64 image; capture the tensors entering the
    tail (per-level obj_embed, box_preds, neck text) and the dict leaving it."""
    torch.manual_seed(14)
    model = YOLOCLIP(backbone_variant="n", num_classes=6, offline_mode=True).eval()
    model.offline_vocabulary = torch.randn(6, 512)
    captured = {"obj": [], "text": None, "box": None}
    hooks = [h.register_forward_hook(lambda m, i, o: captured["obj"].append(o[0].detach()))
             for h in model.contrastive_heads]
    hooks.append(model.neck.register_forward_hook(
        lambda m, i, o: captured.__setitem__("text", o[1].detach())))
    hooks.append(model.box_head.register_forward_hook(
        lambda m, i, o: captured.__setitem__("box", [t.detach() for t in o[0]])))
    with torch.no_grad():
        out = model(torch.rand(2, 3, 64, 64))
    for h in hooks:
        h.remove()
    arrays = {f"obj{i}": t.numpy() for i, t in enumerate(captured["obj"])}
    arrays.update({f"box{i}": t.numpy() for i, t in enumerate(captured["box"])})
    arrays["text"] = captured["text"].contiguous().numpy()
    arrays["text_strides"] = np.array(captured["text"].stride())
    arrays["boxes"] = out["boxes"].numpy()
    arrays["scores"] = out["scores"].numpy()
    arrays["class_ids"] = out["class_ids"].numpy()
    arrays["keys"] = np.array(sorted(out.keys()))
    save("forward_tail_64", **arrays)


def preprocess_cases():
    """Reference YOLOCLIPDetector.preprocess_image (live cv2.resize) on seeded uint8 images:
    up-scaling, down-scaling, an exact 2x decimation and a 1:1 copy; small canvases keep the
    fixture small.  Also checks the numpy restatement of cv2.resize on 300 random shapes."""
    import cv2
    from oracle import ref_port
    rng = np.random.default_rng(15)
    cases = {"up": ((37, 53), (64, 64)), "down": ((150, 91), (64, 64)), "half": ((128, 96), (64, 64)),
             "same": ((64, 40), (64, 64)), "wide": ((45, 301), (96, 160))}
    arrays = {}
    for name, ((h, w), size) in cases.items():
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        det = _bare_detector(image_size=size)
        det.device = "cpu"
        tensor, orig, scale = det.preprocess_image(img)
        arrays[f"{name}_img"] = img
        arrays[f"{name}_out"] = tensor.numpy()
        arrays[f"{name}_meta"] = np.array([size[0], size[1], scale], dtype=np.float64)
        got, _, s2 = ref_port.preprocess_image(img, size)
        assert s2 == scale and np.array_equal(got.numpy(), tensor.numpy()), name
    save("preprocess_cases", **arrays)
    bad = 0
    for t in range(300):
        sh, sw = (int(v) for v in rng.integers(2, 900, 2))
        if t % 10 == 0:
            sh, sw = 2 * int(rng.integers(2, 300)), 2 * int(rng.integers(2, 300))
            dh, dw = sh // 2, sw // 2
        else:
            _, dh, dw = ref_port.letterbox_geometry(sh, sw, (640, 640))
            if dh < 1 or dw < 1:
                continue
        img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        bad += not np.array_equal(cv2.resize(img, (dw, dh)), ref_port.cv2_resize_linear_u8(img, dw, dh))
    assert bad == 0, f"{bad} resize mismatches against the installed cv2"
    print("cv2.resize restatement: 300 random shapes bit-exact")


def tcsp_case():
    """Live reference TextGuidedCSPLayer (random weights, eval mode) on a small feature map: the
    layer's input/output, its state dict and the intermediate entering / leaving the attention."""
    from yolo_clip_detector.model.repvl_pan import TextGuidedCSPLayer
    torch.manual_seed(17)
    layer = TextGuidedCSPLayer(in_channels=48, out_channels=64, text_dim=512, n_bottlenecks=1).eval()
    for m in layer.modules():                      # non-trivial BN statistics
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.2)
            m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.2)
    x = torch.randn(2, 48, 12, 10)
    text = torch.randn(2, 9, 512)
    captured = {}
    hook = layer.bottlenecks[0].register_forward_hook(lambda m, i, o: captured.__setitem__("y", o.detach()))
    with torch.no_grad():
        out = layer(x, text)
        proj = layer.text_proj(text)
    hook.remove()
    arrays = {"x": x.numpy(), "text": text.numpy(), "out": out.numpy(), "y_temp": captured["y"].numpy(),
              "proj": proj.numpy()}
    arrays.update({"sd/" + k: v.numpy() for k, v in layer.state_dict().items()})
    save("tcsp_layer", **arrays)


def head_projection_case():
    """Live reference TextContrastiveHead: the hidden features entering the last (1x1) layer of
    obj_embed_conv, that layer's weight / bias, forward()'s obj_embed, compute_similarity() and
    the class max of yolo_clip.py:198-202 - the reference for the folded projection (row f-2).
    Small hidden / embed dims keep the fixture small."""
    torch.manual_seed(18)
    head = TextContrastiveHead(in_channels=12, embed_dim=128, hidden_dim=64, cls_alpha=1.2, cls_beta=0.05).eval()
    head.obj_embed_conv[2].bias.data.normal_(0, 0.1)
    x = torch.randn(2, 12, 8, 6)
    text = torch.randn(2, 7, 128)
    captured = {}
    hook = head.obj_embed_conv[1].register_forward_hook(lambda m, i, o: captured.__setitem__("h", o.detach().clone()))
    with torch.no_grad():
        obj_embed, _ = head(x)
        sim = head.compute_similarity(obj_embed, text)
        scores, ids = sim.max(dim=1)
    hook.remove()
    save("head_projection", hidden=captured["h"].numpy(), weight=head.obj_embed_conv[2].weight.detach().numpy(),
         bias=head.obj_embed_conv[2].bias.detach().numpy(), text=text.numpy(), obj_embed=obj_embed.numpy(),
         scores=scores.flatten(1).numpy(), class_ids=ids.flatten(1).numpy(), alpha_beta=np.array([1.2, 0.05]))


def vocabulary_case():
    """Reference VocabBuilder.build_offline_vocabulary -> JSON file (stub CLIP encoder), and the
    matrix YOLOCLIP.load_offline_vocabulary stacks from it."""
    import json
    import tempfile
    torch.manual_seed(16)
    names = ["traffic light", "person", "zebra"]
    model = YOLOCLIP(backbone_variant="n", num_classes=3, offline_mode=True).eval()
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "v", "vocab.json")
        model.set_offline_vocabulary(names, save_path=path)
        model2 = YOLOCLIP(backbone_variant="n", num_classes=3, offline_mode=True).eval()
        model2.load_offline_vocabulary(path)
        with open(path) as f:
            text = f.read()
    with open(os.path.join(OUT, "vocab_3cls.json"), "w") as f:
        f.write(text)
    save("vocab_3cls", matrix=model2.offline_vocabulary.cpu().numpy(), names=np.array(names))


def full_forward_case():
    """The whole boundary against the LIVE reference: ``YOLOCLIP.forward``'s
    six-key dict (model/yolo_clip.py:102-223) on a batch of two 64x64 images and
    ``YOLOCLIPDetector.detect`` (inference/detector.py:289-325) on a 48x80 uint8 image.

    Stored: the tensors entering the tail (the neck's per-level maps and its per-image text), the
    state dicts of the reference's three TextContrastiveHead modules and of its BoxHead, every entry
    of the forward dict, and the detection records.  The heads are the reference's own classes built
    with ``hidden_dim=16`` so that their state dicts stay small (the ctor argument exists for that:
    text_contrastive.py:39-47, box_head.py:38-42); the model, its forward and the detector are the
    reference's code unchanged."""
    torch.manual_seed(19)
    names = ["traffic light", "person", "zebra", "kite", "cup", "dog"]
    model = YOLOCLIP(backbone_variant="n", num_classes=len(names), offline_mode=True)
    in_ch = model.backbone.out_channels
    model.contrastive_heads = torch.nn.ModuleList(
        [TextContrastiveHead(in_channels=c, embed_dim=512, hidden_dim=16, cls_alpha=1.0, cls_beta=0.0) for c in in_ch])
    model.box_head = BoxHead(in_channels=in_ch, hidden_dim=16)
    model.eval()
    model.offline_vocabulary = torch.randn(len(names), 512)
    # A random-init network either collapses to nearly constant features (scores 1e-8 apart) or, with
    # arbitrary BatchNorm statistics, explodes (activations 1e8).  Give every BatchNorm a random affine
    # part and CALIBRATE its running statistics on seeded inputs (train-mode passes), so that every
    # layer's activations are O(1) and anchors differ.
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data.uniform_(0.7, 1.3)
            m.bias.data.normal_(0, 0.3)
            m.momentum = 0.2
        elif isinstance(m, torch.nn.Conv2d) and m.bias is not None:
            m.bias.data.normal_(0, 0.1)
    model.train()
    with torch.no_grad():
        for _ in range(40):
            model(torch.rand(4, 3, 64, 64).round())
    model.eval()
    for branch in model.box_head.box_convs:          # boxes of a few cells instead of exp(8) * stride
        bias = branch[2].bias.data.view(4, 17)
        k = torch.arange(17.0)
        bias[0:2] += -1.2 * k
        bias[2:4] += -0.9 * (k - 1.5).abs()
        branch[2].weight.data.mul_(0.3)
    captured = {}
    hook = model.neck.register_forward_hook(
        lambda m, i, o: captured.update(pan=[t.detach().clone() for t in o[0]], text=o[1].detach()))
    arrays = {"names": np.array(names), "vocabulary": model.offline_vocabulary.numpy()}
    for l, head in enumerate(model.contrastive_heads):
        arrays.update({f"sd_head{l}/" + k: v.numpy() for k, v in head.state_dict().items()})
    arrays.update({"sd_box/" + k: v.numpy() for k, v in model.box_head.state_dict().items()})
    # ---- forward dict, batch 2 --------------------------------------------------------------------
    with torch.no_grad():
        out = model(torch.rand(2, 3, 64, 64).round())       # saturated noise: strong local structure
    assert list(out.keys()) == ["boxes", "scores", "class_ids", "obj_embeddings", "text_embeddings", "box_preds"]
    for l, t in enumerate(captured["pan"]):
        arrays[f"fwd_pan{l}"] = t.numpy()
    arrays["fwd_text"] = captured["text"].contiguous().numpy()
    arrays["fwd_text_strides"] = np.array(captured["text"].stride())
    arrays["fwd_boxes"] = out["boxes"].numpy()
    arrays["fwd_scores"] = out["scores"].numpy()
    arrays["fwd_class_ids"] = out["class_ids"].numpy()
    arrays["fwd_obj_embeddings"] = out["obj_embeddings"].numpy()
    assert out["text_embeddings"] is captured["text"] or torch.equal(out["text_embeddings"], captured["text"])
    for l, t in enumerate(out["box_preds"]):
        arrays[f"fwd_box_preds{l}"] = t.numpy()
    assert out["class_ids"].dtype == torch.int64
    # ---- detect(), one image ----------------------------------------------------------------------
    rng = np.random.default_rng(19)
    img = (rng.integers(0, 2, (48, 80, 3)) * 255).astype(np.uint8)
    det = _bare_detector(conf=0.0, iou=0.45, image_size=(64, 64), class_names=names)
    det.device = "cpu"
    det.model = model
    det.use_offline_vocab = True
    tensor, _, scale = det.preprocess_image(img)
    with torch.no_grad():
        probe = np.sort(model(tensor)["scores"][0].numpy())
    # threshold in the middle of the widest gap between neighbouring scores of the lower half, so
    # that a 1e-5 score difference cannot move an anchor across it
    lo, hi = len(probe) // 5, len(probe) // 2
    g = lo + int(np.argmax(np.diff(probe[lo:hi])))
    conf = float((probe[g] + probe[g + 1]) / 2)
    det.conf_threshold = conf
    dets = det.detect(img)
    arrays["det_image"] = img
    arrays["det_tensor"] = tensor.numpy()
    arrays["det_conf_iou_scale"] = np.array([conf, 0.45, scale], dtype=np.float64)
    for l, t in enumerate(captured["pan"]):
        arrays[f"det_pan{l}"] = t.numpy()
    arrays["det_text"] = captured["text"].contiguous().numpy()
    arrays["det_box"] = np.array([d["box"] for d in dets], dtype=np.int64).reshape(-1, 4)
    arrays["det_score"] = np.array([d["score"] for d in dets], dtype=np.float64)
    arrays["det_class"] = np.array([d["class_id"] for d in dets], dtype=np.int64)
    arrays["det_name"] = np.array([d["class_name"] for d in dets])
    hook.remove()
    print(f"  full forward: scores {out['scores'].min():.3f}..{out['scores'].max():.3f}, conf {conf:.4f}, "
          f"gap {probe[g + 1] - probe[g]:.2e}, {int((probe > conf).sum())} candidates, {len(dets)} detections of 84 anchors; box span {arrays['det_box'].min()}..{arrays['det_box'].max()}")
    save("forward_full_64", **arrays)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    cases = [similarity_cases, decode_case, nms_cases, postprocess_case, forward_tail_case, preprocess_cases,
             vocabulary_case, tcsp_case, head_projection_case, full_forward_case]
    for case in cases:
        if not only or case.__name__ in only:
            case()
