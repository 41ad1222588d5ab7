"""GPU parity tests: every kernel of libovdet.so, called through the C ABI (ctypes), against the
CPU oracle (oracle/ref_port.py) and the golden fixtures generated from the live reference.

Tolerances (BASELINE.json north_star):
  * logits, fp32 recipe (3-pass hi/lo split):  max|d| / max|ref| <= 1e-3  and
    |d| <= 1e-3 * max(|ref|, 0.05) element-wise; measured ~1e-5.
  * logits, bf16 recipe: |d| <= 8e-3 absolute at alpha = 1 (unit-norm operands, K = 512).
  * boxes: 1e-4 relative (+1e-3 px absolute).
  * NMS / ordering / indices / classes: bit-exact on identical inputs.
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ref_port

pytestmark = pytest.mark.gpu

FP32_REL = 1e-3
BF16_ABS = 8e-3


@pytest.fixture(scope="module")
def ov(cuda_device):
    from ovdet import ops, synth, heads, detector, pipeline  # noqa: F401
    import ovdet
    ovdet._cabi.lib()
    return ovdet


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def assert_logits_close(got: torch.Tensor, ref: torch.Tensor, precision: str, alpha: float = 1.0):
    got, ref = got.double().cpu(), ref.double().cpu()
    d = (got - ref).abs()
    if precision == "fp32":
        assert d.max() / ref.abs().max() <= FP32_REL
        assert bool((d <= FP32_REL * torch.clamp(ref.abs(), min=0.05)).all())
    else:
        assert d.max() <= BF16_ABS * max(1.0, abs(alpha))


# ------------------------------------------------------------------------------------------
# K1
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("split", [False, True])
def test_l2norm_regions(ov, cuda_device, split):
    from ovdet import ops
    torch.manual_seed(0)
    shapes = [(9, 7), (5, 5), (1, 3)]          # ragged: not multiples of the 64/32 anchor tiles
    embs = [torch.randn(3, 128, h, w, device=cuda_device) * (1 + l) for l, (h, w) in enumerate(shapes)]
    embs[1][0, :, 2, 2] = 0.0                   # a zero vector: eps clamp
    op, inv = ops.l2norm_regions(embs, split=split)
    flat = torch.cat([e.flatten(2).transpose(1, 2) for e in embs], dim=1)      # [B, A, D]
    ref_inv = 1.0 / flat.norm(dim=-1).clamp_min(1e-12)
    zero = flat.norm(dim=-1) == 0
    torch.testing.assert_close(inv[~zero], ref_inv[~zero], rtol=2e-6, atol=0)
    assert bool((inv[zero] == 1e12).all())
    hi = flat.to(torch.bfloat16)
    assert torch.equal(op[..., :128], hi)
    if split:
        lo = (flat - hi.float()).to(torch.bfloat16)
        assert torch.equal(op[..., 128:], lo)


@pytest.mark.parametrize("split", [False, True])
def test_l2norm_text(ov, cuda_device, split):
    from ovdet import ops
    torch.manual_seed(1)
    base = torch.randn(11, 2, 64, device=cuda_device).transpose(0, 1)         # [B,C,D], batch not outer
    assert base.stride() == (64, 128, 1)
    op = ops.l2norm_text(base, split=split)
    ref = torch.nn.functional.normalize(base, dim=-1)
    hi = op[..., :64].float()
    assert (hi - ref).abs().max() <= 2 ** -8 * ref.abs().max()
    if split:
        assert ((hi + op[..., 64:].float()) - ref).abs().max() <= 2e-5
    shared = torch.randn(5, 64, device=cuda_device)
    assert ops.l2norm_text(shared.unsqueeze(0).expand(4, -1, -1), split=split).shape[0] == 1


# ------------------------------------------------------------------------------------------
# K2
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["sim_batched_d512", "sim_shared_affine_d64"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_similarity_golden(ov, cuda_device, golden_dir, name, precision):
    """compute_similarity through the drop-in module against fixtures from the live reference."""
    from ovdet.heads import TextContrastiveHead
    g = _load(golden_dir, name)
    alpha, beta = float(g["alpha"]), float(g["beta"])
    obj = torch.from_numpy(g["obj"]).to(cuda_device)
    text = torch.from_numpy(g["text"]).to(cuda_device)
    head = TextContrastiveHead(8, embed_dim=obj.shape[1], cls_alpha=alpha, cls_beta=beta,
                               precision=precision)
    sim = head.compute_similarity(obj, text)
    assert sim.shape == g["sim"].shape
    assert tuple(sim.stride()) == tuple(g["strides"])          # memory is [B,HW,C], like the reference
    assert_logits_close(sim.contiguous(), torch.from_numpy(g["sim"]), precision, alpha)


@pytest.mark.parametrize("classes,batched", [(80, False), (1203, False), (300, True), (17, True), (256, False)])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_similarity_vs_oracle(ov, cuda_device, classes, batched, precision):
    from ovdet import ops
    torch.manual_seed(classes)
    b, d = 3, 512
    shapes = [(20, 20), (10, 10), (5, 5)]       # 525 anchors: ragged against the 128-row tile
    embs = [torch.randn(b, d, h, w) for h, w in shapes]
    text = torch.randn(b, classes, d) if batched else torch.randn(classes, d).unsqueeze(0).expand(b, -1, -1)
    ref = torch.cat([ref_port.compute_similarity(e, text, 1.0, 0.0).flatten(2).transpose(1, 2)
                     for e in embs], dim=1)                                    # [B, A, C]
    split = precision == "fp32"
    dev_embs = [e.to(cuda_device) for e in embs]
    rop, inv = ops.l2norm_regions(dev_embs, split=split)
    top = ops.l2norm_text(text.to(cuda_device) if batched else text[0].to(cuda_device), split=split)
    logits, rmax, rarg = ops.similarity(rop, top, inv, d, split=split, logits_dtype=torch.float32,
                                        want_max=True)
    torch.cuda.synchronize()
    assert_logits_close(logits, ref, precision)
    # fused max / argmax is exact with respect to the logits the same launch wrote
    m, a = logits.max(dim=-1)
    assert torch.equal(rmax, m)
    assert torch.equal(rarg.long(), a)
    # and the separate row-max kernel agrees
    m2, a2 = ops.rowmax(logits)
    assert torch.equal(m2, m) and torch.equal(a2.long(), a)
    if precision == "fp32":
        agree = (rarg.cpu().long() == ref.argmax(dim=-1)).float().mean().item()
        assert agree >= 0.999


def test_similarity_bf16_logits_and_max_only(ov, cuda_device):
    from ovdet import ops
    torch.manual_seed(5)
    b, d, c = 2, 512, 1203
    embs = [torch.randn(b, d, 16, 16, device=cuda_device)]
    text = torch.randn(c, d, device=cuda_device)
    rop, inv = ops.l2norm_regions(embs)
    top = ops.l2norm_text(text)
    l32, m32, a32 = ops.similarity(rop, top, inv, d, logits_dtype=torch.float32, want_max=True)
    l16, m16, a16 = ops.similarity(rop, top, inv, d, logits_dtype=torch.bfloat16, want_max=True)
    _, m0, a0 = ops.similarity(rop, top, inv, d, logits_dtype=None, want_max=True)
    assert torch.equal(l16, l32.to(torch.bfloat16))
    assert torch.equal(m16, m32) and torch.equal(a16, a32)       # max is taken before rounding
    assert torch.equal(m0, m32) and torch.equal(a0, a32)


def test_rowmax_ties_lowest_index(ov, cuda_device):
    from ovdet import ops
    x = torch.zeros(4, 100, device=cuda_device)
    x[0, 7] = x[0, 50] = 2.0
    x[1, 99] = 1.0
    x[2] = -3.0
    m, a = ops.rowmax(x)
    assert a.tolist() == [7, 99, 0, 0]
    assert m.tolist() == [2.0, 1.0, -3.0, 0.0]


# ------------------------------------------------------------------------------------------
# K3
# ------------------------------------------------------------------------------------------
def test_decode_golden(ov, cuda_device, golden_dir):
    from ovdet.heads import BoxHead
    g = _load(golden_dir, "decode_3level")
    head = BoxHead([8, 8, 8])
    for keys, want in ((("p0", "p1", "p2"), "boxes"), (("n0", "n1", "n2"), "boxes_noise")):
        preds = [torch.from_numpy(g[k]).to(cuda_device) for k in keys]
        boxes = head.decode_boxes(preds, None)
        torch.testing.assert_close(boxes.cpu(), torch.from_numpy(g[want]), rtol=1e-4, atol=1e-3)
    grid = head._create_grid(2, 8, 8, 8, cuda_device)
    assert grid.dtype == torch.int64
    np.testing.assert_array_equal(grid.cpu().numpy(), g["grid0"])


def test_decode_filter_vs_oracle(ov, cuda_device):
    from ovdet import ops, synth
    inp = synth.make_inputs(batch=2, image_size=320, num_classes=10, embed_dim=64, seed=3)
    grids = [ref_port.create_grid(2, p.shape[2], p.shape[3], s) for p, s in zip(inp.box_preds, inp.strides)]
    ref = ref_port.decode_boxes(inp.box_preds, grids)
    torch.manual_seed(3)                        # the scores do not depend on which tests ran before
    scores = torch.rand(2, ref.shape[1]) - 0.3
    scores[0, 5] = float("nan")
    for act in ("none", "sigmoid"):
        boxes, sact, mask = ops.decode_filter([p.to(cuda_device) for p in inp.box_preds], inp.strides,
                                              scores=scores.to(cuda_device), conf=0.25, activation=act)
        torch.testing.assert_close(boxes.cpu(), ref, rtol=1e-4, atol=1e-3)
        s = torch.sigmoid(scores) if act == "sigmoid" else scores
        want = (s > 0.25)
        bits = _unpack(mask.cpu(), ref.shape[1])
        if act == "sigmoid":
            torch.testing.assert_close(sact.cpu(), s, rtol=1e-6, atol=1e-7, equal_nan=True)
            near = (s - 0.25).abs() < 1e-6
            assert torch.equal(bits[~near], want[~near])
        else:
            assert torch.equal(bits, want)
            assert not bits[0, 5]               # NaN never passes


def _unpack(mask: torch.Tensor, anchors: int) -> torch.Tensor:
    m = mask.to(torch.int64) & 0xffffffff
    bits = (m.unsqueeze(-1) >> torch.arange(32)) & 1
    return bits.reshape(mask.shape[0], -1)[:, :anchors].bool()


# ------------------------------------------------------------------------------------------
# K4
# ------------------------------------------------------------------------------------------
def test_nms_golden_bit_exact(ov, cuda_device, golden_dir):
    """`_nms` on the exact arrays the live reference was run on: same kept indices, same order."""
    from ovdet.detector import Detector
    g = _load(golden_dir, "nms_cases")
    det = Detector(device=str(cuda_device))
    names = sorted({k[:-len("_boxes")] for k in g.files if k.endswith("_boxes")})
    assert "empty" in names and "degenerate64" in names and "dense1000" in names
    for name in names:
        keep = det._nms(g[name + "_boxes"], g[name + "_scores"], float(g[name + "_thr"]))
        np.testing.assert_array_equal(np.asarray(keep, dtype=np.int64), g[name + "_keep"], err_msg=name)


def _rand_boxes(rng, n, span, size):
    xy = rng.uniform(0, span, (n, 2)).astype(np.float32)
    wh = rng.uniform(1, size, (n, 2)).astype(np.float32)
    return np.concatenate([xy, xy + wh], axis=1).astype(np.float32)


@pytest.mark.parametrize("n,span,size", [(513, 400, 120), (1500, 300, 80), (5000, 2000, 150), (9000, 3000, 100)])
def test_nms_multi_chunk_and_big_sort(ov, cuda_device, n, span, size):
    """More candidates than one 512-chunk / one 4096-key sort block; scores pairwise distinct."""
    from ovdet.detector import Detector
    rng = np.random.default_rng(n)
    boxes = _rand_boxes(rng, n, span, size)
    scores = rng.permutation(np.linspace(0.01, 0.99, n)).astype(np.float32)
    assert len(np.unique(scores)) == n
    want = ref_port.nms(boxes.copy(), scores.copy(), 0.45)
    got = Detector(device=str(cuda_device))._nms(boxes, scores, 0.45)
    np.testing.assert_array_equal(np.asarray(got), np.asarray(want))


def test_nms_tie_rule(ov, cuda_device):
    """Equal scores: higher index first (== stable argsort reversed)."""
    from ovdet.detector import Detector
    boxes = np.array([[i * 100, 0, i * 100 + 10, 10] for i in range(5)], np.float32)   # disjoint
    scores = np.array([.5, .7, .5, .7, .1], np.float32)
    got = Detector(device=str(cuda_device))._nms(boxes, scores, 0.45)
    assert got == [3, 1, 2, 0, 4]
    assert got == ref_port.nms(boxes, scores, 0.45, stable_ties=True)


@pytest.mark.parametrize("class_aware,topk", [(False, 0), (True, 0), (False, 100), (True, 37)])
def test_postprocess_batch_vs_oracle(ov, cuda_device, class_aware, topk):
    from ovdet.detector import Detector
    rng = np.random.default_rng(7)
    b, a = 4, 2100
    boxes = np.stack([_rand_boxes(rng, a, 500, 200) for _ in range(b)])
    scores = np.stack([rng.permutation(np.linspace(-0.2, 0.95, a)).astype(np.float32) for _ in range(b)])
    scores[3] = -1.0                                            # an image without survivors
    classes = rng.integers(0, 6, (b, a)).astype(np.int64)
    sizes = [(480, 640), (500, 500), (300, 200), (640, 640)]
    scales = [1.0, 0.8, 640 / 300, 1.0]
    det = Detector(device=str(cuda_device))
    res = det.postprocess_batch({"boxes": torch.from_numpy(boxes), "scores": torch.from_numpy(scores),
                                 "class_ids": torch.from_numpy(classes)}, sizes, scales,
                                class_aware=class_aware, topk=topk)
    for i in range(b):
        want = ref_port.postprocess_image(boxes[i], scores[i], classes[i], sizes[i], scales[i],
                                          class_aware=class_aware, topk=topk or None)
        k = int(res.count[i])
        assert k == len(want["keep"])
        assert int(res.candidates[i]) == int((scores[i] > 0.25).sum())
        np.testing.assert_array_equal(res.keep[i, :k].cpu().numpy(), want["keep"])
        np.testing.assert_array_equal(res.anchor[i, :k].cpu().numpy(), want["anchor_idx"])
        np.testing.assert_array_equal(res.boxes[i, :k].cpu().numpy(), want["boxes"])
        np.testing.assert_array_equal(res.scores[i, :k].cpu().numpy(), want["scores"])
        np.testing.assert_array_equal(res.classes[i, :k].cpu().numpy(), want["class_ids"])


def test_postprocess_golden(ov, cuda_device, golden_dir):
    """postprocess_detections (image 0 of the dict, like the reference) against the dict lists
    the live reference produced."""
    from ovdet.detector import Detector
    g = _load(golden_dir, "postprocess_b3")
    names = [f"thing{i}" for i in range(40)]
    det = Detector(class_names=names, image_size=(128, 128), device=str(cuda_device))
    for i in range(3):
        out = {k: torch.from_numpy(g[k][i:i + 1].copy()) for k in ("boxes", "scores", "class_ids")}
        dets = det.postprocess_detections(out, tuple(int(v) for v in g[f"img{i}_orig"]), float(g[f"img{i}_scale"]))
        assert len(dets) == len(g[f"img{i}_score"]) > 0
        np.testing.assert_array_equal(np.array([d["box"] for d in dets]), g[f"img{i}_box"])
        np.testing.assert_array_equal(np.array([d["score"] for d in dets]), g[f"img{i}_score"])
        np.testing.assert_array_equal(np.array([d["class_id"] for d in dets]), g[f"img{i}_class"])
        assert dets[0]["class_name"] == str(g[f"img{i}_name0"])


def test_max_det_truncates(ov, cuda_device):
    from ovdet import ops
    n = 300
    boxes = torch.tensor([[i * 20.0, 0, i * 20.0 + 10, 10] for i in range(n)], device=cuda_device).reshape(1, n, 4)
    scores = torch.linspace(0.3, 0.9, n, device=cuda_device).reshape(1, n)
    res = ops.nms_batched(boxes, scores, max_det=50)
    assert int(res.count[0]) == 50
    assert res.anchor[0, :50].tolist() == list(range(n - 1, n - 51, -1))


# ------------------------------------------------------------------------------------------
# the whole path
# ------------------------------------------------------------------------------------------
def test_forward_tail_golden(ov, cuda_device, golden_dir):
    """The tensors a real reference forward handed to its tail (per-level obj_embed, the neck's
    per-image text with its odd strides, box_preds) -> boxes / scores / class_ids."""
    from ovdet.heads import head_tail
    g = _load(golden_dir, "forward_tail_64")
    objs = [torch.from_numpy(g[f"obj{i}"]).to(cuda_device) for i in range(3)]
    preds = [torch.from_numpy(g[f"box{i}"]).to(cuda_device) for i in range(3)]
    text = torch.from_numpy(g["text"]).to(cuda_device)
    out = head_tail(objs, text, preds, precision="fp32")
    assert out["boxes"].shape == (2, 84, 4) and out["class_ids"].dtype == torch.int64
    assert_logits_close(out["scores"], torch.from_numpy(g["scores"]), "fp32")
    rel = ((out["boxes"].cpu() - torch.from_numpy(g["boxes"])).abs()
           / torch.from_numpy(g["boxes"]).abs().clamp_min(1.0)).max()
    assert rel <= 1e-4
    assert (out["class_ids"].cpu().numpy() == g["class_ids"]).mean() >= 0.98


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_predict_vs_oracle(ov, cuda_device, precision):
    """Config-2-shaped run at oracle size: synthetic conv outputs -> detections."""
    from ovdet import synth
    from ovdet.detector import Detector
    from ovdet.pipeline import HeadConfig
    inp = synth.make_inputs(batch=4, image_size=320, num_classes=80, seed=99)
    tail = ref_port.head_tail(inp.obj_embeds, inp.text_batched(), inp.box_preds)
    det = Detector(device=str(cuda_device), config=HeadConfig(precision=precision))
    sizes = [(320, 320)] * 4
    res = det.predict([e.to(cuda_device) for e in inp.obj_embeds], [p.to(cuda_device) for p in inp.box_preds],
                      inp.text.to(cuda_device), sizes, [1.0] * 4)
    pipe = next(iter(det._pipelines.values()))
    assert_logits_close(pipe.scores, tail["scores"], precision)
    # (i) bit-exact post-processing when the oracle is fed the scores / boxes the device produced
    fed = {"boxes": pipe.boxes.cpu(), "scores": pipe.scores.cpu(), "class_ids": pipe.class_ids.cpu().long()}
    want = ref_port.postprocess_batch(fed, sizes, [1.0] * 4)
    total = 0
    for i in range(4):
        k = int(res.count[i])
        total += k
        assert k == len(want[i]["keep"])
        np.testing.assert_array_equal(res.keep[i, :k].cpu().numpy(), want[i]["keep"])
        np.testing.assert_array_equal(res.classes[i, :k].cpu().numpy(), want[i]["class_ids"])
        np.testing.assert_array_equal(res.boxes[i, :k].cpu().numpy(), want[i]["boxes"])
    assert total > 20                       # the synthetic inputs give NMS real work
    # (ii) end to end against the pure-oracle path (oracle scores -> oracle NMS): the kept anchor sets
    # differ only where a device score sits within its error of the threshold / of a neighbour's score,
    # or an IoU within rounding distance of 0.45.  Reported per image, bounded per precision.
    pure = ref_port.postprocess_batch(tail, sizes, [1.0] * 4)
    mism, kept_ref = _index_mismatches(res, pure)
    print(f"\n[e2e index mismatches, {precision}] per image {mism} of {kept_ref} kept")
    # fp32: |dscore| ~ 1e-5 - a flip needs two overlapping candidates whose scores are closer than that;
    # bf16: |dscore| up to 4e-3 reorders near-equal overlapping candidates (planted scores span 0.79-0.87)
    bound = 0.01 if precision == "fp32" else 0.15
    assert sum(mism) <= max(1, int(bound * sum(kept_ref))), (mism, kept_ref)


def _index_mismatches(res, pure):
    """Per image: size of the symmetric difference between the device's kept anchor set and the
    oracle's, and the oracle's kept count."""
    mism, kept = [], []
    for i, want in enumerate(pure):
        got = set(res.anchor[i, :int(res.count[i])].tolist())
        ref = set(int(a) for a in want["anchor_idx"].tolist())
        mism.append(len(got ^ ref))
        kept.append(len(ref))
    return mism, kept


def test_config2_shape_fp32_fused_streaming_vs_oracle(ov, cuda_device):
    """BASELINE configs[1] at its real per-image size (640^2: 6400 + 1600 + 400 anchors, 80 prompts,
    fp32-accurate): the fused kernel's streaming three-pass mode over 50 anchor tiles per image of
    the first level (the 8-slot A ring wraps 19 times per CTA), two images, against the oracle:
    logits inside the fp32 bar, classes equal, boxes 1e-4, post-processing bit-exact on identical
    inputs, and the end-to-end kept sets reported."""
    from ovdet import ops, synth
    from ovdet.pipeline import HeadConfig, HeadPipeline
    b, classes, s = 2, 80, 640
    inp = synth.make_inputs(batch=b, image_size=s, num_classes=classes, seed=11)
    tail = ref_port.head_tail(inp.obj_embeds, inp.text_batched(), inp.box_preds)
    shapes = [(s // 8, s // 8), (s // 16, s // 16), (s // 32, s // 32)]
    pipe = HeadPipeline(b, shapes, classes, HeadConfig(precision="fp32", logits_dtype="fp32"), device=cuda_device)
    pipe.set_vocabulary(inp.text.to(cuda_device))
    sizes = [(s, s)] * b
    pipe.set_geometry(sizes, [1.0] * b)
    objs = [e.to(cuda_device) for e in inp.obj_embeds]
    preds = [p.to(cuda_device) for p in inp.box_preds]
    res = pipe.run(objs, preds)
    torch.cuda.synchronize()
    assert pipe.last_path == "fused_fp32"
    ref_logits = torch.cat([ref_port.compute_similarity(e, inp.text_batched()).flatten(2).transpose(1, 2)
                            for e in inp.obj_embeds], dim=1)
    assert_logits_close(pipe.logits, ref_logits, "fp32")
    assert (pipe.logits.cpu() - ref_logits).abs().max().item() <= 3e-5
    assert_logits_close(pipe.scores, tail["scores"], "fp32")
    assert (pipe.class_ids.cpu().long() == tail["class_ids"]).float().mean() >= 0.999
    torch.testing.assert_close(pipe.boxes.cpu(), tail["boxes"], rtol=1e-4, atol=1e-3)
    fed = {"boxes": pipe.boxes.cpu(), "scores": pipe.scores.cpu(), "class_ids": pipe.class_ids.cpu().long()}
    want = ref_port.postprocess_batch(fed, sizes, [1.0] * b)
    for i in range(b):
        k = int(res.count[i])
        assert k == len(want[i]["keep"]) and k > 20
        np.testing.assert_array_equal(res.keep[i, :k].cpu().numpy(), want[i]["keep"])
        np.testing.assert_array_equal(res.classes[i, :k].cpu().numpy(), want[i]["class_ids"])
        np.testing.assert_array_equal(res.boxes[i, :k].cpu().numpy(), want[i]["boxes"])
    mism, kept_ref = _index_mismatches(res, ref_port.postprocess_batch(tail, sizes, [1.0] * b))
    print(f"\n[e2e index mismatches, configs[1] shape, fp32] per image {mism} of {kept_ref} kept")
    assert sum(mism) <= max(1, int(0.01 * sum(kept_ref)))
    # scores-only launch (what the bench times) == the logits launch
    pipe2 = HeadPipeline(b, shapes, classes, HeadConfig(precision="fp32"), device=cuda_device)
    pipe2.set_vocabulary(inp.text.to(cuda_device))
    pipe2.run(objs, preds)
    torch.cuda.synchronize()
    assert pipe2.last_path == "fused_fp32" and torch.equal(pipe2.scores, pipe.scores)


def test_full_size_properties(ov, cuda_device):
    """BASELINE config 2 size (batch 64 @ 640^2, 80 prompts): size-independent properties."""
    from ovdet import synth, ops
    from ovdet.pipeline import HeadConfig, HeadPipeline
    inp = synth.make_inputs(batch=64, image_size=640, num_classes=80, device=cuda_device, seed=2)
    pipe = HeadPipeline(64, [(80, 80), (40, 40), (20, 20)], 80, HeadConfig(logits_dtype="fp32"), device=cuda_device)
    pipe.set_vocabulary(inp.text)
    res = pipe.run(inp.obj_embeds, inp.box_preds)
    torch.cuda.synchronize()
    assert pipe.scores.abs().max() <= 1.0 + 1e-3                       # cosine similarity
    m, a = pipe.logits.max(dim=-1)
    assert torch.equal(m, pipe.scores) and torch.equal(a.int(), pipe.class_ids)
    cnt = res.count.cpu()
    assert (cnt > 10).all() and (cnt <= res.candidates.cpu()).all()
    passed = (pipe.scores > 0.25).sum(dim=1).int()
    assert torch.equal(passed, res.candidates)
    for i in (0, 17, 63):
        k = int(cnt[i])
        s = res.scores[i, :k]
        assert bool((s[:-1] >= s[1:]).all())                            # kept order = score desc
        assert len(set(res.anchor[i, :k].tolist())) == k
        assert torch.equal(pipe.scores[i][res.anchor[i, :k].long()], s)
        b = res.boxes[i, :k]
        area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
        lt = torch.maximum(b[:, None, :2], b[None, :, :2])
        rb = torch.minimum(b[:, None, 2:], b[None, :, 2:])
        inter = (rb - lt).clamp_min(0).prod(-1)
        iou = inter / (area[:, None] + area[None, :] - inter + 1e-7)
        iou.fill_diagonal_(0)
        assert iou.max() <= 0.45 + 1e-6                                  # no kept pair overlaps
    # idempotence: NMS over the kept set keeps every row, in the same order (rows beyond an
    # image's count are zero-area boxes with score 0: they sort last and suppress nothing)
    kmax = int(cnt.max())
    again = ops.nms_batched(res.boxes[:, :kmax].contiguous(), res.scores[:, :kmax].contiguous())
    for i in (0, 63):
        k = int(cnt[i])
        assert again.anchor[i, :k].tolist() == list(range(k))


@pytest.mark.parametrize("precision,classes", [("bf16", 1203), ("fp32", 80), ("fp32", 1203)])
def test_per_image_text_bench_shape_vs_oracle(ov, cuda_device, precision, classes):
    """The reference's real forward hands the head PER-IMAGE text [B, C, D] whose batch is not the
    outer memory dimension (the neck's I-Pooling output, repvl_pan.py:173-182, strides (D, B*D, 1);
    yolo_clip.py:171,182): the batched similarity at the bench's per-image shape (640^2, three images
    so that a CTA pair never straddles two images' text) against the oracle, K1b normalising B*C rows
    inside the step."""
    from ovdet import synth
    from ovdet.pipeline import HeadConfig, HeadPipeline
    b, s = 3, 640
    inp = synth.make_inputs(batch=b, image_size=s, num_classes=classes, seed=23)
    g = torch.Generator().manual_seed(5)
    text = (inp.text.unsqueeze(1) + 0.05 * torch.randn(classes, b, 512, generator=g)).transpose(0, 1)   # [B, C, D]
    assert text.stride() == (512, b * 512, 1) and not torch.equal(text[0], text[1])
    tail = ref_port.head_tail(inp.obj_embeds, text, inp.box_preds)
    shapes = [(s // 8, s // 8), (s // 16, s // 16), (s // 32, s // 32)]
    pipe = HeadPipeline(b, shapes, classes, HeadConfig(precision=precision), device=cuda_device, per_image_text=True)
    sizes = [(s, s)] * b
    pipe.set_geometry(sizes, [1.0] * b)
    text_dev = torch.empty(classes, b, 512, device=cuda_device).transpose(0, 1)
    text_dev.copy_(text)
    assert text_dev.stride() == (512, b * 512, 1)
    objs = [e.to(cuda_device) for e in inp.obj_embeds]
    preds = [p.to(cuda_device) for p in inp.box_preds]
    res = pipe.run(objs, preds, text=text_dev)
    torch.cuda.synchronize()
    if precision == "bf16":
        assert pipe.last_path == "fused" and pipe.last_single_call
    elif classes <= 128:
        assert pipe.last_path == "fused_fp32"
    assert_logits_close(pipe.scores, tail["scores"], precision)
    agree = (pipe.class_ids.cpu().long() == tail["class_ids"]).float().mean()
    assert agree >= (0.999 if precision == "fp32" else 0.97)
    torch.testing.assert_close(pipe.boxes.cpu(), tail["boxes"], rtol=1e-4, atol=1e-3)
    fed = {"boxes": pipe.boxes.cpu(), "scores": pipe.scores.cpu(), "class_ids": pipe.class_ids.cpu().long()}
    want = ref_port.postprocess_batch(fed, sizes, [1.0] * b)
    for i in range(b):
        k = int(res.count[i])
        assert k == len(want[i]["keep"]) and k > 20
        np.testing.assert_array_equal(res.keep[i, :k].cpu().numpy(), want[i]["keep"])
        np.testing.assert_array_equal(res.boxes[i, :k].cpu().numpy(), want[i]["boxes"])
    mism, kept_ref = _index_mismatches(res, ref_port.postprocess_batch(tail, sizes, [1.0] * b))
    print(f"\n[e2e index mismatches, per-image text, {precision}, C={classes}] per image {mism} of {kept_ref} kept")
    assert sum(mism) <= max(1, int((0.01 if precision == "fp32" else 0.15) * sum(kept_ref)))
    # the per-stage launch sequence gives the same bytes as the single C call
    if precision == "bf16":
        scores = pipe.scores.clone()
        pipe.run(objs, preds, text=text_dev, events={})
        torch.cuda.synchronize()
        assert torch.equal(pipe.scores, scores)


@pytest.mark.parametrize("image_size,classes", [(1280, 1203), (640, 4800)])
def test_config4_config5_shapes_vs_oracle(ov, cuda_device, image_size, classes):
    """BASELINE configs[3] (1280^2: 33 600 anchors per image - the multi-chunk K4 path and the
    vectorised K3) and configs[4] (4800 prompts: 38 class tiles per anchor tile) at full per-image
    size, two images: scores inside the bf16 bar, post-processing bit-exact on identical inputs."""
    from ovdet import synth
    from ovdet.pipeline import HeadConfig, HeadPipeline
    b = 2
    inp = synth.make_inputs(batch=b, image_size=image_size, num_classes=classes, seed=7)
    tail = ref_port.head_tail(inp.obj_embeds, inp.text_batched(), inp.box_preds)
    s = image_size
    shapes = [(s // 8, s // 8), (s // 16, s // 16), (s // 32, s // 32)]
    pipe = HeadPipeline(b, shapes, classes, HeadConfig(precision="bf16"), device=cuda_device)
    pipe.set_vocabulary(inp.text.to(cuda_device))
    sizes = [(s, s)] * b
    pipe.set_geometry(sizes, [1.0] * b)
    res = pipe.run([e.to(cuda_device) for e in inp.obj_embeds], [p.to(cuda_device) for p in inp.box_preds])
    torch.cuda.synchronize()
    assert pipe.last_path == "fused"
    assert_logits_close(pipe.scores, tail["scores"], "bf16")
    agree = (pipe.class_ids.cpu().long() == tail["class_ids"]).float().mean()
    assert agree >= 0.97                                  # bf16 near-ties between background classes
    torch.testing.assert_close(pipe.boxes.cpu(), tail["boxes"], rtol=1e-4, atol=1e-3)
    fed = {"boxes": pipe.boxes.cpu(), "scores": pipe.scores.cpu(), "class_ids": pipe.class_ids.cpu().long()}
    want = ref_port.postprocess_batch(fed, sizes, [1.0] * b)
    for i in range(b):
        k = int(res.count[i])
        assert k == len(want[i]["keep"]) and k > 20
        np.testing.assert_array_equal(res.keep[i, :k].cpu().numpy(), want[i]["keep"])
        np.testing.assert_array_equal(res.classes[i, :k].cpu().numpy(), want[i]["class_ids"])
        np.testing.assert_array_equal(res.boxes[i, :k].cpu().numpy(), want[i]["boxes"])


@pytest.mark.parametrize("image_size,precision", [(416, "bf16"), (416, "auto"), (608, "bf16"), (480, "auto")])
def test_odd_level_sizes_stay_on_the_fused_kernel(ov, cuda_device, image_size, precision):
    """Image sizes whose P5 level has an odd number of cells (416 -> 13x13, 480 -> 15x15, 608 -> 19x19):
    TMA needs 16-byte rows, so that level is re-pitched (ovdet_repitch_rows) and the step stays on the
    fused kernel and the single C call; same bars as everywhere else vs the oracle."""
    from ovdet import ops, synth
    from ovdet.pipeline import HeadConfig, HeadPipeline
    b, classes = 3, 300
    inp = synth.make_inputs(batch=b, image_size=image_size, num_classes=classes, seed=21)
    tail = ref_port.head_tail(inp.obj_embeds, inp.text_batched(), inp.box_preds)
    s = image_size
    shapes = [(s // 8, s // 8), (s // 16, s // 16), (s // 32, s // 32)]
    assert (shapes[2][0] * shapes[2][1]) % 4 != 0
    dev_embs = [e.to(cuda_device) for e in inp.obj_embeds]
    assert not ops.fused_supported(dev_embs) and ops.fused_supported(ops.tma_addressable(dev_embs))
    # the helper alone: aligned levels pass through, the odd one becomes a padded view with equal values
    padded = ops.tma_addressable(dev_embs)
    assert padded[0] is dev_embs[0] and padded[1] is dev_embs[1] and padded[2].stride(1) % 4 == 0
    assert torch.equal(padded[2], dev_embs[2])
    pipe = HeadPipeline(b, shapes, classes, HeadConfig(precision=precision), device=cuda_device)
    pipe.set_vocabulary(inp.text.to(cuda_device))
    sizes = [(s, s)] * b
    pipe.set_geometry(sizes, [1.0] * b)
    res = pipe.run(dev_embs, [p.to(cuda_device) for p in inp.box_preds])
    torch.cuda.synchronize()
    assert pipe.last_path == "fused" and pipe.last_single_call
    assert_logits_close(pipe.scores, tail["scores"], "bf16" if precision == "bf16" else "fp32")
    torch.testing.assert_close(pipe.boxes.cpu(), tail["boxes"], rtol=1e-4, atol=1e-3)
    fed = {"boxes": pipe.boxes.cpu(), "scores": pipe.scores.cpu(), "class_ids": pipe.class_ids.cpu().long()}
    want = ref_port.postprocess_batch(fed, sizes, [1.0] * b)
    for i in range(b):
        k = int(res.count[i])
        assert k == len(want[i]["keep"]) and k > 5
        np.testing.assert_array_equal(res.keep[i, :k].cpu().numpy(), want[i]["keep"])
        np.testing.assert_array_equal(res.boxes[i, :k].cpu().numpy(), want[i]["boxes"])
    # the drop-in tail takes the same route
    from ovdet.heads import head_tail
    out = head_tail(dev_embs, inp.text.to(cuda_device).unsqueeze(0).expand(b, -1, -1),
                    [p.to(cuda_device) for p in inp.box_preds], precision=precision)
    assert_logits_close(out["scores"], tail["scores"], "bf16" if precision == "bf16" else "fp32")


# ------------------------------------------------------------------------------------------
# K1+K2 fused (fp32 NCHW in, A operand resident in tensor memory)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("classes,batched,dim", [(80, False, 512), (1203, False, 512), (300, True, 512),
                                                 (17, True, 256), (129, False, 64)])
def test_similarity_fused_vs_oracle(ov, cuda_device, classes, batched, dim):
    from ovdet import ops
    torch.manual_seed(classes + 1)
    b = 3
    shapes = [(20, 20), (10, 12), (4, 5)]       # 400 | 120 | 20 anchors: full, ragged and tiny tiles
    embs = [torch.randn(b, dim, h, w) * (0.5 + l) for l, (h, w) in enumerate(shapes)]
    embs[2][1, :, 1, 1] = 0.0                   # a zero vector: eps clamp, logit = beta
    text = torch.randn(b, classes, dim) if batched else torch.randn(classes, dim).unsqueeze(0).expand(b, -1, -1)
    alpha, beta = 1.7, -0.2
    ref = torch.cat([ref_port.compute_similarity(e, text, alpha, beta).flatten(2).transpose(1, 2)
                     for e in embs], dim=1)
    dev_embs = [e.to(cuda_device) for e in embs]
    assert ops.fused_supported(dev_embs)
    top = ops.l2norm_text(text.to(cuda_device) if batched else text[0].to(cuda_device))
    inv = torch.empty(b, ref.shape[1], device=cuda_device)
    logits, rmax, rarg = ops.similarity_fused(dev_embs, top, alpha, beta, logits_dtype=torch.float32,
                                              want_max=True, inv_norm=inv)
    torch.cuda.synchronize()
    assert_logits_close(logits, ref, "bf16", alpha)
    m, a = logits.max(dim=-1)
    assert torch.equal(rmax, m) and torch.equal(rarg.long(), a)
    flat = torch.cat([e.flatten(2).transpose(1, 2) for e in embs], dim=1)
    ref_inv = 1.0 / flat.norm(dim=-1).clamp_min(1e-12)
    nz = flat.norm(dim=-1) > 0
    torch.testing.assert_close(inv.cpu()[nz], ref_inv[nz], rtol=3e-6, atol=0)
    # the two-kernel path computes the same bf16 products: near-identical logits
    rop, inv2 = ops.l2norm_regions(dev_embs)
    l2, _, _ = ops.similarity(rop, top, inv2, dim, alpha, beta, logits_dtype=torch.float32)
    assert (logits - l2).abs().max() <= 2e-5 * max(1.0, abs(alpha))
    # max-only launch gives the same answer
    _, m0, a0 = ops.similarity_fused(dev_embs, top, alpha, beta, logits_dtype=None, want_max=True)
    assert torch.equal(m0, rmax) and torch.equal(a0, rarg)
    # padded leading dimension (16-byte rows): the 16-byte-store epilogue writes the same values
    for dt, per16 in ((torch.float32, 4), (torch.bfloat16, 8)):
        ldc = (classes + per16 - 1) // per16 * per16
        buf = torch.full((b, ref.shape[1], ldc), float("nan"), device=cuda_device, dtype=dt)
        lp, mp, ap = ops.similarity_fused(dev_embs, top, alpha, beta, logits=buf[..., :classes], want_max=True)
        torch.cuda.synchronize()
        assert lp.data_ptr() == buf.data_ptr() and torch.equal(mp, rmax) and torch.equal(ap, rarg)
        assert torch.equal(lp.float(), logits.to(dt).float())
    # scores only (no argmax): the raw-accumulator fast path gives bit-identical maxima
    _, m1, a1 = ops.similarity_fused(dev_embs, top, alpha, beta, logits_dtype=None, want_max=True, want_arg=False)
    assert a1 is None and torch.equal(m1, rmax)


@pytest.mark.parametrize("batch,shapes", [(1, [(20, 20), (10, 12)]), (5, [(12, 12)]), (1, [(4, 4)])])
def test_similarity_fused_odd_tile_counts(ov, cuda_device, batch, shapes):
    """CTA pairs take tiles 2i / 2i+1: an odd number of 128-anchor tiles (5, 5 x 2 - 1 ... and a
    single tile) leaves the second CTA of the last pair with a tile past the end."""
    from ovdet import ops
    torch.manual_seed(batch)
    embs = [torch.randn(batch, 512, h, w) for h, w in shapes]
    text = torch.randn(77, 512)
    ref = torch.cat([ref_port.compute_similarity(e, text.unsqueeze(0).expand(batch, -1, -1)).flatten(2).transpose(1, 2)
                     for e in embs], dim=1)
    top = ops.l2norm_text(text.to(cuda_device))
    logits, rmax, rarg = ops.similarity_fused([e.to(cuda_device) for e in embs], top,
                                              logits_dtype=torch.float32, want_max=True)
    torch.cuda.synchronize()
    assert_logits_close(logits, ref, "bf16", 1.0)
    m, a = logits.max(dim=-1)
    assert torch.equal(rmax, m) and torch.equal(rarg.long(), a)
    _, m0, a0 = ops.similarity_fused([e.to(cuda_device) for e in embs], top, logits_dtype=None, want_max=True)
    assert torch.equal(m0, rmax) and torch.equal(a0, rarg)


def test_fused_unsupported_shape_is_an_error(ov, cuda_device):
    from ovdet import ops
    embs = [torch.randn(1, 64, 5, 5, device=cuda_device)]      # hw = 25: row stride not 16-byte aligned
    assert not ops.fused_supported(embs)
    top = ops.l2norm_text(torch.randn(3, 64, device=cuda_device))
    with pytest.raises(ValueError):
        ops.similarity_fused(embs, top)


# ------------------------------------------------------------------------------------------
# P1 / P2 ("next" rows): letterbox pre-processing and int-truncated records
# ------------------------------------------------------------------------------------------
def test_letterbox_golden_bit_exact(ov, cuda_device, golden_dir):
    from ovdet.detector import Detector
    g = _load(golden_dir, "preprocess_cases")
    for name in ("up", "down", "half", "same", "wide"):
        h, w, scale = g[f"{name}_meta"]
        det = Detector(image_size=(int(h), int(w)), device=str(cuda_device))
        tensor, orig, s = det.preprocess_image(g[f"{name}_img"])
        assert s == scale and tensor.shape == (1, 3, int(h), int(w))
        np.testing.assert_array_equal(tensor.cpu().numpy(), g[f"{name}_out"])
        np.testing.assert_array_equal(orig, g[f"{name}_img"])


def test_letterbox_batch_vs_oracle_bit_exact(ov, cuda_device):
    from ovdet import ops
    rng = np.random.default_rng(21)
    shapes = [(480, 640), (1280, 1280), (375, 500), (1, 9), (2000, 31), (640, 640), (90, 1300)]
    shapes += [tuple(int(v) for v in rng.integers(2, 1500, 2)) for _ in range(70)]
    shapes = [s for s in shapes if min(ref_port.letterbox_geometry(s[0], s[1], (640, 640))[1:]) >= 1]
    assert len(shapes) > 64                                    # more than one launch's descriptor table
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    out, scales = ops.letterbox([torch.from_numpy(i).to(cuda_device) for i in imgs], (640, 640))
    torch.cuda.synchronize()
    for i, img in enumerate(imgs):
        want, _, s = ref_port.preprocess_image(img, (640, 640))
        assert s == scales[i]
        np.testing.assert_array_equal(out[i].cpu().numpy(), want[0].numpy(), err_msg=str(shapes[i]))
    # a strided view (row pitch > 3 * width) is read in place
    big = torch.from_numpy(rng.integers(0, 256, (300, 500, 3), dtype=np.uint8)).to(cuda_device)
    view = big[10:250, 20:400]
    got, _ = ops.letterbox([view], (320, 320))
    want, _, _ = ref_port.preprocess_image(view.cpu().numpy(), (320, 320))
    np.testing.assert_array_equal(got.cpu().numpy(), want.numpy())


def test_pack_boxes_and_records(ov, cuda_device):
    from ovdet import ops
    from ovdet.detector import Detector
    boxes = torch.tensor([[[1.9, 2.1, 639.99, 0.0], [5.5, 6.5, 7.5, 8.5], [9.0, 9.0, 9.0, 9.0]],
                          [[0.2, 0.9, 3.7, 4.0], [1.0, 1.0, 1.0, 1.0], [2.0, 2.0, 2.0, 2.0]]], device=cuda_device)
    count = torch.tensor([2, 1], device=cuda_device, dtype=torch.int32)
    packed = ops.pack_boxes_i32(boxes, count)
    want = boxes.cpu().numpy().astype(int)
    want[0, 2:] = 0
    want[1, 1:] = 0
    np.testing.assert_array_equal(packed.cpu().numpy(), want)
    det = Detector(class_names=["a", "b"], device=str(cuda_device))
    res = ops.NmsResult(boxes=boxes, scores=torch.full((2, 3), 0.5, device=cuda_device),
                        classes=torch.ones(2, 3, device=cuda_device, dtype=torch.int32), count=count)
    rec = det.to_records(res, 0)
    assert [r["box"] for r in rec] == [[1, 2, 639, 0], [5, 6, 7, 8]] and rec[0]["class_name"] == "b"


def test_vocabulary_operand_and_detect(ov, cuda_device, golden_dir):
    """JSON vocabulary -> cached operand -> Detector.detect on an image through a stand-in
    convolutional model; equals the oracle fed with the same conv outputs."""
    from ovdet import ops, synth
    from ovdet.detector import Detector
    from ovdet.vocabulary import Vocabulary
    vocab = Vocabulary.load(os.path.join(golden_dir, "vocab_3cls.json"))
    op = vocab.operand(cuda_device)
    assert op.shape == (1, 3, 512) and op.dtype == torch.bfloat16 and vocab.operand(cuda_device) is op
    ref = torch.nn.functional.normalize(vocab.embeddings, dim=-1)
    assert (op[0].float().cpu() - ref).abs().max() < 4e-3
    inp = synth.make_inputs(batch=1, image_size=160, num_classes=3, seed=9)
    feats = ([e.to(cuda_device) for e in inp.obj_embeds], [p.to(cuda_device) for p in inp.box_preds],
             inp.text.to(cuda_device))
    det = Detector(image_size=(160, 160), device=str(cuda_device), feature_fn=lambda x, prompts: feats)
    det.load_offline_vocabulary(os.path.join(golden_dir, "vocab_3cls.json"))
    assert det.class_names == ["traffic light", "person", "zebra"]
    img = np.random.default_rng(3).integers(0, 256, (120, 200, 3), dtype=np.uint8)
    records = det.detect(img)
    pipe = next(iter(det._pipelines.values()))
    fed = {"boxes": pipe.boxes.cpu(), "scores": pipe.scores.cpu(), "class_ids": pipe.class_ids.cpu().long()}
    scale, _, _ = ref_port.letterbox_geometry(120, 200, (160, 160))
    want = ref_port.postprocess_image(fed["boxes"][0].numpy(), fed["scores"][0].numpy(),
                                      fed["class_ids"][0].numpy(), (120, 200), scale,
                                      class_names=det.class_names)["detections"]
    assert len(records) == len(want) > 0
    for a, b in zip(records, want):
        assert a["box"] == b["box"] and a["class_id"] == b["class_id"] and a["class_name"] == b["class_name"]
        assert a["score"] == b["score"]


# ------------------------------------------------------------------------------------------
# N1 ("next" row): max-sigmoid text attention of the neck
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("c,classes,shape,batched", [(32, 5, (20, 20), True), (64, 80, (12, 10), False),
                                                     (128, 300, (8, 8), True), (64, 1203, (40, 40), False)])
def test_max_sigmoid_attention_vs_oracle(ov, cuda_device, c, classes, shape, batched):
    from ovdet import ops
    torch.manual_seed(c + classes)
    b = 3
    y = torch.randn(b, c, *shape) * 0.7
    y[1, :, 0, 1] = 0.0                                          # zero activation: weight = sigmoid(0)
    text = torch.randn(b, classes, c) * 0.5 if batched else (torch.randn(classes, c) * 0.5)
    tb = text if batched else text.unsqueeze(0).expand(b, -1, -1)
    want, want_max = ref_port.max_sigmoid_attention(y, tb)
    yd = y.to(cuda_device)
    td = text.to(cuda_device)
    out, smax = ops.max_sigmoid_attention(yd, td, precise=True, return_scores=True)
    torch.cuda.synchronize()
    # fp32 tolerance (north_star: 1e-3 relative); the three-pass product measures ~1e-6
    scale = want_max.abs().max().item()
    assert (smax.cpu() - want_max).abs().max().item() <= 2e-5 * max(scale, 1.0)
    torch.testing.assert_close(out.cpu(), want.contiguous(), rtol=2e-5, atol=2e-6)
    assert out.shape == y.shape and out.is_contiguous()
    # one bf16 pass: stated separately, |dscore| <= 8e-3 * |y| |t'| ; the weight moves by <= 1/4 of that
    out16, smax16 = ops.max_sigmoid_attention(yd, td, precise=False, return_scores=True)
    bound = 8e-3 * (y.flatten(2).norm(dim=1).max() * tb.norm(dim=-1).max()).item()
    assert (smax16.cpu() - want_max).abs().max().item() <= bound
    assert (out16.cpu() - want).abs().max().item() <= 0.25 * bound * y.abs().max().item() + 1e-6


def test_tcsp_layer_golden(ov, cuda_device, golden_dir):
    from ovdet.neck import TextGuidedCSPLayer
    g = _load(golden_dir, "tcsp_layer")
    layer = TextGuidedCSPLayer(48, 64, 512, n_bottlenecks=1).eval()
    layer.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}, strict=True)
    layer = layer.to(cuda_device)
    x, text = torch.from_numpy(g["x"]).to(cuda_device), torch.from_numpy(g["text"]).to(cuda_device)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # the (out-of-scope) convolutions in plain fp32
    try:
        with torch.no_grad():
            out = layer(x, text)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    torch.testing.assert_close(out.cpu(), torch.from_numpy(g["out"]), rtol=1e-4, atol=2e-5)
    # shared vocabulary (stride-0 expand) gives the same result as the materialised batch
    shared = text[:1].expand(2, -1, -1)
    with torch.no_grad():
        a = layer(x, shared)
        bb = layer(x, shared.contiguous())
    torch.testing.assert_close(a, bb, rtol=1e-5, atol=1e-6)


def test_graph_replay_equals_eager(ov, cuda_device):
    """HeadPipeline.capture / replay: same detections as the eager launches, and a replay reads
    the CURRENT contents of the captured input tensors."""
    from ovdet import synth
    from ovdet.pipeline import HeadConfig, HeadPipeline
    shapes = [(32, 32), (16, 16), (8, 8)]            # TMA-addressable: the fused CTA-pair kernel is captured
    a = synth.make_inputs(batch=2, image_size=256, num_classes=90, seed=31, device=cuda_device)
    b = synth.make_inputs(batch=2, image_size=256, num_classes=90, seed=32, device=cuda_device)
    pipe = HeadPipeline(2, shapes, 90, HeadConfig(precision="bf16", max_det=64), device=cuda_device)
    pipe.set_vocabulary(a.text)

    def snapshot(res):
        torch.cuda.synchronize()
        return [t.clone() for t in (res.count, res.boxes, res.scores, res.classes, res.anchor, res.keep)]

    want_a = snapshot(pipe.run(a.obj_embeds, a.box_preds))
    want_b = snapshot(pipe.run(b.obj_embeds, b.box_preds))
    assert int(want_a[0].sum()) > 0 and not torch.equal(want_a[4], want_b[4])
    bufs_e = [t.clone() for t in a.obj_embeds]
    bufs_p = [t.clone() for t in a.box_preds]
    pipe.capture(bufs_e, bufs_p)
    assert pipe.last_path == "fused"
    def same(got, want):                      # rows past count[b] are not written by a step
        assert torch.equal(got[0], want[0])
        for i in range(2):
            k = int(want[0][i])
            for g, w in zip(got[1:], want[1:]):
                assert torch.equal(g[i, :k], w[i, :k])

    same(snapshot(pipe.replay()), want_a)
    for dst, src in zip(bufs_e + bufs_p, b.obj_embeds + b.box_preds):
        dst.copy_(src)
    same(snapshot(pipe.replay()), want_b)
    # the forked variant (decode beside the similarity kernel, threshold inside K4) gives the same lists
    pipe.capture(bufs_e, bufs_p, parallel_decode=True)
    same(snapshot(pipe.replay()), want_b)


def test_single_call_step_with_changing_inputs(ov, cuda_device):
    """ovdet_head_step launches K3 and K4 with programmatic dependent launch (they start while the
    preceding kernel drains).  Every intermediate and every kept row must equal the per-stage
    path when the inputs change from step to step - a K3 that read the scores through the
    read-only cache path passed same-input tests and returned the PREVIOUS step's pass mask."""
    from ovdet import synth
    from ovdet.pipeline import HeadConfig, HeadPipeline
    shapes = [(32, 32), (16, 16), (8, 8)]
    ins = [synth.make_inputs(batch=2, image_size=256, num_classes=90, seed=s, device=cuda_device)
           for s in (41, 42, 43)]
    pipe = HeadPipeline(2, shapes, 90, HeadConfig(precision="bf16", max_det=64), device=cuda_device)
    pipe.set_vocabulary(ins[0].text)
    ref = []
    for x in ins:
        r = pipe.run(x.obj_embeds, x.box_preds, events={})           # one launch per stage, no PDL
        torch.cuda.synchronize()
        ref.append([t.clone() for t in (pipe.scores, pipe.pass_mask, pipe.boxes, r.count,
                                        r.boxes, r.scores, r.classes, r.anchor)])
    assert not torch.equal(ref[0][1], ref[1][1]) and not torch.equal(ref[1][1], ref[2][1])
    for it in range(30):
        i = it % 3
        r = pipe.run(ins[i].obj_embeds, ins[i].box_preds)            # single C call
        torch.cuda.synchronize()
        assert pipe.last_path == "fused"
        got = [pipe.scores, pipe.pass_mask, pipe.boxes, r.count, r.boxes, r.scores, r.classes, r.anchor]
        for j in range(4):
            assert torch.equal(got[j], ref[i][j]), (it, j)
        for b, k in enumerate(ref[i][3].tolist()):
            for j in range(4, 8):
                assert torch.equal(got[j][b, :k], ref[i][j][b, :k]), (it, j, b)


def test_single_call_step_bench_shape_changing_inputs(ov, cuda_device):
    """The same check at the benchmark's per-image shape (640^2, 1203 prompts, batch 32): here the
    similarity kernel runs for hundreds of microseconds, so a K3 / K4 that did not wait for it would
    read a mostly unwritten score array."""
    from ovdet import synth
    from ovdet.pipeline import HeadConfig, HeadPipeline
    shapes = [(80, 80), (40, 40), (20, 20)]
    ins = [synth.make_inputs(batch=32, image_size=640, num_classes=1203, seed=s, device=cuda_device)
           for s in (71, 72)]
    pipe = HeadPipeline(32, shapes, 1203, HeadConfig(precision="bf16", max_det=300), device=cuda_device)
    pipe.set_vocabulary(ins[0].text)
    ref = []
    for x in ins:
        r = pipe.run(x.obj_embeds, x.box_preds, events={})
        torch.cuda.synchronize()
        ref.append([t.clone() for t in (pipe.scores, pipe.pass_mask, pipe.boxes, r.count, r.anchor)])
    assert not torch.equal(ref[0][1], ref[1][1])
    for it in range(8):
        i = it % 2
        r = pipe.run(ins[i].obj_embeds, ins[i].box_preds)
        torch.cuda.synchronize()
        for j, g in enumerate((pipe.scores, pipe.pass_mask, pipe.boxes, r.count)):
            assert torch.equal(g, ref[i][j]), (it, j)
        for b, k in enumerate(ref[i][3].tolist()):
            assert torch.equal(r.anchor[b, :k], ref[i][4][b, :k]), (it, b)


# ------------------------------------------------------------------------------------------
# Vocabulary-parallel exchange (SURVEY 8 e / f-4) on ONE device: several virtual ranks share the
# GPU, their "peer" buffers are plain local pointers - the kernels cannot tell the difference.
# The multi-process path (CUDA IPC mappings over NVLink) is tools/vp_check.py under torchrun.
# ------------------------------------------------------------------------------------------
def _virtual_ranks(world, batch, shapes, classes, cfg, dev):
    from ovdet import vocab_parallel as vp
    heads = [vp.VocabParallelHead(batch, shapes, classes, cfg, device=dev, rank=r, world=world)
             for r in range(world)]
    ptrs = [h.buffer.ptr for h in heads]
    for h in heads:
        h.connect(ptrs)
    return heads


@pytest.mark.parametrize("world,classes", [(2, 1203), (3, 200), (8, 90)])
def test_vocab_parallel_virtual_ranks_equal_full_vocabulary(ov, cuda_device, world, classes):
    """Class shards reduced through the in-kernel key exchange == one launch over the whole
    vocabulary: scores bit-exact, classes equal, and the detections after K3/K4 identical, over
    several steps with changing inputs (the parity hand-back of the key arrays)."""
    from ovdet import synth
    from ovdet.pipeline import HeadConfig, HeadPipeline
    shapes = [(32, 32), (16, 16), (8, 8)]
    cfg = HeadConfig(precision="bf16", max_det=64)
    ins = [synth.make_inputs(batch=2, image_size=256, num_classes=classes, seed=s, device=cuda_device)
           for s in (51, 52, 53)]
    text = ins[0].text
    full = HeadPipeline(2, shapes, classes, cfg, device=cuda_device)
    full.set_vocabulary(text)
    heads = _virtual_ranks(world, 2, shapes, classes, cfg, cuda_device)
    try:
        for h in heads:
            h.set_vocabulary(text)
        for it in range(5):
            x = ins[it % 3]
            r = full.run(x.obj_embeds, x.box_preds, events={})
            torch.cuda.synchronize()
            want = [t.clone() for t in (full.scores, full.class_ids, r.count, r.boxes, r.scores, r.classes, r.anchor)]
            for h in heads:
                h.similarity(x.obj_embeds)
            for h in heads:
                h.signal()
            for h in heads:
                h.merge()
                res = h.finish(x.box_preds)
                torch.cuda.synchronize()
                assert not h.timed_out()
                assert torch.equal(h.scores, want[0]), (it, h.rank)
                assert torch.equal(h.class_ids, want[1]), (it, h.rank)
                assert torch.equal(res.count, want[2])
                for b, k in enumerate(want[2].tolist()):
                    for g, w in zip((res.boxes, res.scores, res.classes, res.anchor), want[3:]):
                        assert torch.equal(g[b, :k], w[b, :k])
        assert int(want[2].sum()) > 0
    finally:
        for h in heads:
            h.close()


def test_vocab_parallel_single_call_and_graph(ov, cuda_device):
    """ovdet_head_step_vp (the whole sharded step behind one C call) and its CUDA-graph replay, with
    a world of one rank - the exchange then runs against the rank's own buffer, through the same
    kernels and counters as across GPUs."""
    from ovdet import synth
    from ovdet import vocab_parallel as vp
    from ovdet.pipeline import HeadConfig, HeadPipeline
    shapes = [(32, 32), (16, 16), (8, 8)]
    cfg = HeadConfig(precision="bf16", max_det=64)
    ins = [synth.make_inputs(batch=2, image_size=256, num_classes=300, seed=s, device=cuda_device)
           for s in (61, 62)]
    full = HeadPipeline(2, shapes, 300, cfg, device=cuda_device)
    full.set_vocabulary(ins[0].text)
    head = vp.VocabParallelHead(2, shapes, 300, cfg, device=cuda_device, rank=0, world=1)
    head.connect([head.buffer.ptr])
    head.set_vocabulary(ins[0].text)
    try:
        def check(res, x):
            r = full.run(x.obj_embeds, x.box_preds, events={})
            torch.cuda.synchronize()
            assert torch.equal(head.scores, full.scores) and torch.equal(head.class_ids, full.class_ids)
            assert torch.equal(res.count, r.count) and int(r.count.sum()) > 0
            for b, k in enumerate(r.count.tolist()):
                assert torch.equal(res.anchor[b, :k], r.anchor[b, :k])
        for it in range(4):
            x = ins[it % 2]
            check(head.run(x.obj_embeds, x.box_preds), x)
        bufs_e = [t.clone() for t in ins[0].obj_embeds]
        bufs_p = [t.clone() for t in ins[0].box_preds]
        head.capture(bufs_e, bufs_p)
        for it in range(5):                                   # odd count: both key-array parities
            x = ins[it % 2]
            for dst, src in zip(bufs_e + bufs_p, x.obj_embeds + x.box_preds):
                dst.copy_(src)
            check(head.replay(), x)
        assert not head.timed_out()
    finally:
        head.close()


def test_vocab_parallel_across_gpus(ov, cuda_device):
    """The real thing when the box has more than one GPU: one process per GPU under torchrun, CUDA
    IPC peer mappings over NVLink, both exchange implementations and the graph replay against one
    GPU holding the whole vocabulary (tools/vp_check.py --assert-parity)."""
    import socket
    import subprocess
    import sys
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (measured runs: profiles/r1_vocab_parallel_n2.jsonl, _n8.jsonl)")
    n = 2 if n < 4 else (4 if n < 8 else 8)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(root, "tools", "vp_check.py"),
           "--classes", "1203", "--batch", "1", "3", "--image-size", "256", "--steps", "20", "--warmup", "3",
           "--assert-parity"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_vocab_parallel_wait_is_bounded(ov, cuda_device):
    """A rank whose peer never signals does not hang the GPU: the wait expires, is reported, the
    step carries NO detections (sentinel scores, nothing unpacked, key arrays not handed back), the
    condition is sticky, the next run() raises, and reset() re-arms the exchange."""
    from ovdet import synth
    from ovdet.pipeline import HeadConfig
    shapes = [(32, 32), (16, 16), (8, 8)]
    x = synth.make_inputs(batch=1, image_size=256, num_classes=40, seed=5, device=cuda_device)
    heads = _virtual_ranks(2, 1, shapes, 40, HeadConfig(precision="bf16", max_det=16), cuda_device)
    try:
        for h in heads:
            h.set_vocabulary(x.text)
            h.timeout_ms = h.first_timeout_ms = 20
        heads[0].similarity(x.obj_embeds)
        heads[0].signal()                       # rank 1 never contributes
        heads[0].merge()
        res = heads[0].finish(x.box_preds)
        torch.cuda.synchronize()
        assert heads[0].timed_out()
        assert bool(torch.isneginf(heads[0].scores).all()) and int(res.count.sum()) == 0
        with pytest.raises(RuntimeError, match="did not signal"):
            heads[0].run(x.obj_embeds, x.box_preds)
        # sticky: a second merge without a reset unpacks nothing either, even with every flag present
        heads[1].similarity(x.obj_embeds)
        heads[1].signal()
        heads[0].merge()
        torch.cuda.synchronize()
        assert bool(torch.isneginf(heads[0].scores).all())
        for h in heads:
            h.reset()
        assert not heads[0].timed_out()
        for h in heads:
            h.similarity(x.obj_embeds)
        for h in heads:
            h.signal()
        for h in heads:
            h.merge()
            r = h.finish(x.box_preds)
        torch.cuda.synchronize()
        assert not heads[0].timed_out() and not heads[1].timed_out()
        assert torch.equal(heads[0].scores, heads[1].scores) and bool(torch.isfinite(heads[0].scores).all())
        assert int(r.count.sum()) > 0
    finally:
        for h in heads:
            h.close()


def test_score_keys_device_equals_host_twin(ov, cuda_device):
    """ovdet_pack_score_keys / ovdet_unpack_score_keys (the all-reduce baseline's kernels) against
    the numpy twin the gloo tests use."""
    from ovdet import vocab_parallel as vp
    g = torch.Generator(device="cpu").manual_seed(9)
    scores = torch.randn(3, 1000, generator=g)
    scores[0, :6] = torch.tensor([0.0, -0.0, float("inf"), float("-inf"), 1e-45, -1e-45])
    classes = torch.randint(0, 100000, (3, 1000), generator=g, dtype=torch.int32)
    keys = vp.pack_score_keys(scores.to(cuda_device), classes.to(cuda_device), 12345)
    want = vp.pack_keys_host(scores.numpy(), classes.numpy(), 12345)
    assert np.array_equal(keys.cpu().numpy(), want)
    s2 = torch.empty_like(scores, device=cuda_device)
    c2 = torch.empty_like(classes, device=cuda_device)
    vp.unpack_score_keys(keys, s2, c2)
    hs, hc = vp.unpack_keys_host(want)
    assert np.array_equal(s2.cpu().numpy(), hs) and np.array_equal(c2.cpu().numpy(), hc)
    assert np.array_equal(hc, classes.numpy() + 12345)


# ------------------------------------------------------------------------------------------
# K4 properties (hypothesis): random box sets, duplicates, degenerate boxes, ties in IoU
# ------------------------------------------------------------------------------------------
def test_nms_properties_hypothesis(ov, cuda_device):
    from hypothesis import given, settings, strategies as st, HealthCheck
    from ovdet import ops

    coord = st.integers(0, 40)

    @st.composite
    def box_sets(draw):
        n = draw(st.integers(1, 80))
        boxes = []
        for _ in range(n):
            x1, y1 = draw(coord), draw(coord)
            w, h = draw(st.integers(0, 25)), draw(st.integers(0, 25))     # zero-area boxes included
            boxes.append([x1, y1, x1 + w, y1 + h])
        scores = draw(st.lists(st.integers(0, 10_000), min_size=n, max_size=n, unique=True))
        thr = draw(st.sampled_from([0.0, 0.3, 0.45, 0.5, 1.0]))
        return np.asarray(boxes, np.float32), np.asarray(scores, np.float32) / 10_000, thr

    @settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck))
    @given(box_sets())
    def check(case):
        boxes, scores, thr = case
        want = ref_port.nms(boxes, scores, thr, stable_ties=True)
        b = torch.from_numpy(boxes).to(cuda_device).reshape(1, -1, 4)
        s = torch.from_numpy(scores).to(cuda_device).reshape(1, -1)
        res = ops.nms_batched(b, s, iou_thr=thr)
        k = int(res.count[0])
        got = res.anchor[0, :k].tolist()
        assert got == [int(i) for i in want]                                 # bit-exact keep list, in order
        kept = boxes[got]
        for i in range(len(got)):                                            # no kept pair above the threshold
            if i:
                assert (ref_port.compute_iou(kept[i], kept[:i]) <= thr).all()
        # idempotence: NMS of the kept set keeps everything
        again = ops.nms_batched(torch.from_numpy(kept).to(cuda_device).reshape(1, -1, 4),
                                torch.from_numpy(scores[got]).to(cuda_device).reshape(1, -1), iou_thr=thr)
        assert int(again.count[0]) == len(got)

    check()


# ------------------------------------------------------------------------------------------
# f-2 ("next" row): the head's 1x1 projection folded into the similarity
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("hidden_dim,classes,batched,batch", [(256, 300, False, 2), (256, 80, True, 3),
                                                             (64, 17, False, 1), (96, 1203, False, 2)])
def test_similarity_projected_vs_oracle(ov, cuda_device, hidden_dim, classes, batched, batch):
    from ovdet import ops
    torch.manual_seed(hidden_dim + classes)
    embed = 512 if hidden_dim == 256 else 128
    shapes = [(20, 20), (10, 12), (4, 5)]                       # 400 | 120 | 20 anchors
    hidden = [torch.nn.functional.silu(torch.randn(batch, hidden_dim, h, w)) for h, w in shapes]
    hidden[1][0, :, 0, 0] = 0.0
    ws = [torch.randn(embed, hidden_dim, 1, 1) * (2.0 / embed) ** 0.5 for _ in shapes]
    bs = [torch.randn(embed) * 0.1 for _ in shapes]
    text = torch.randn(batch, classes, embed) if batched else torch.randn(classes, embed)
    tb = text if batched else text.unsqueeze(0).expand(batch, -1, -1)
    alpha, beta = 1.3, -0.1
    want, want_ids, embeds = ref_port.project_similarity_max(hidden, ws, bs, tb, alpha, beta)
    dev_hidden = [h.to(cuda_device) for h in hidden]
    level_ops = [ops.project_vocabulary(text.to(cuda_device), w.to(cuda_device), b.to(cuda_device))
                 for w, b in zip(ws, bs)]
    inv = torch.empty(batch, want.shape[1], device=cuda_device)
    rmax, rarg = ops.similarity_projected(dev_hidden, level_ops, classes, alpha, beta, inv_norm=inv)
    torch.cuda.synchronize()
    # one bf16 pass, stated separately from the fp32 bar: |dscore| <= 8e-3 * alpha
    assert (rmax.cpu() - want).abs().max().item() <= 8e-3 * alpha
    flat = torch.cat([e.flatten(2).transpose(1, 2) for e in embeds], dim=1)
    torch.testing.assert_close(inv.cpu(), 1.0 / flat.norm(dim=-1).clamp_min(1e-12), rtol=5e-3, atol=0)
    # argmax: equal, or a near-tie within the bf16 tolerance
    ids = rarg.cpu().long()
    sims = torch.cat([ref_port.compute_similarity(e, tb, alpha, beta).flatten(2).transpose(1, 2) for e in embeds], dim=1)
    picked = sims.gather(-1, ids.unsqueeze(-1)).squeeze(-1)
    assert (want - picked).max().item() <= 1.6e-2 * alpha
    assert (ids == want_ids).float().mean().item() >= 0.97
    # scores only
    m1, a1 = ops.similarity_projected(dev_hidden, level_ops, classes, alpha, beta, want_arg=False)
    assert a1 is None and torch.equal(m1, rmax)


def test_projected_pipeline_vs_oracle(ov, cuda_device):
    """Detector.predict with the 1x1 projection folded in: detections equal the oracle's
    post-processing of the pipeline's own scores, and the scores match conv + similarity + max."""
    from ovdet import synth
    from ovdet.detector import Detector
    from ovdet.pipeline import HeadConfig
    pin = synth.make_projected_inputs(batch=2, image_size=256, num_classes=90, seed=12)
    det = Detector(device=str(cuda_device), config=HeadConfig(precision="bf16"))
    proj = [(w.to(cuda_device), b.to(cuda_device)) for w, b in pin.projections()]
    sizes, scales = [(256, 256), (200, 240)], [1.0, 0.8]
    res = det.predict([h.to(cuda_device) for h in pin.hidden], [p.to(cuda_device) for p in pin.box_preds],
                      pin.text.to(cuda_device), sizes, scales, projections=proj)
    torch.cuda.synchronize()
    pipe = next(iter(det._pipelines.values()))
    assert pipe.last_path == "projected"
    want, want_ids, _ = ref_port.project_similarity_max(pin.hidden, pin.weights, pin.biases, pin.text_batched())
    assert (pipe.scores.cpu() - want).abs().max().item() <= 8e-3
    strong = want > 0.25 + 1.6e-2
    assert torch.equal(pipe.class_ids.cpu().long()[strong], want_ids[strong])
    fed = {"boxes": pipe.boxes.cpu(), "scores": pipe.scores.cpu(), "class_ids": pipe.class_ids.cpu().long()}
    ref = ref_port.postprocess_batch(fed, sizes, scales)
    kept = 0
    for i in range(2):
        k = int(res.count[i])
        kept += k
        np.testing.assert_array_equal(res.keep[i, :k].cpu().numpy(), ref[i]["keep"])
        np.testing.assert_array_equal(res.boxes[i, :k].cpu().numpy(), ref[i]["boxes"])
        np.testing.assert_array_equal(res.classes[i, :k].cpu().numpy(), ref[i]["class_ids"])
    assert kept > 0


def test_edge_shapes_empty_and_huge(ov, cuda_device):
    """Empty batch through every stage (a no-op, not an error); more anchors than the resident NMS
    path addresses (> 65536: workspace path with a sparse pass mask); an image with no survivor."""
    from ovdet import ops
    from ovdet.detector import _pack_mask
    dev = cuda_device
    # ---- empty batch
    e = [torch.empty(0, 512, 8, 8, device=dev)]
    top = ops.l2norm_text(torch.randn(5, 512, device=dev))
    _, m, a = ops.similarity_fused(e, top, want_max=True)
    assert m.shape == (0, 64) and a.shape == (0, 64)
    boxes, _, mask = ops.decode_filter([torch.empty(0, 68, 8, 8, device=dev)], (8,), scores=m, conf=0.25)
    assert boxes.shape == (0, 64, 4)
    res = ops.nms_batched(boxes, m, max_det=8)
    torch.cuda.synchronize()
    assert res.count.numel() == 0
    # ---- 70 000 anchors, 300 candidates + one image with none
    rng = np.random.default_rng(5)
    n = 70_000
    bx = torch.from_numpy(np.stack([_rand_boxes(rng, n, 3000, 120) for _ in range(2)])).to(dev)
    sc = torch.from_numpy(np.stack([rng.permutation(np.linspace(0.0, 1.0, n)).astype(np.float32) for _ in range(2)])).to(dev)
    passed = sc > (1.0 - 300.5 / n)
    passed[1] = False
    res = ops.nms_batched(bx, sc, pass_mask=_pack_mask(passed), iou_thr=0.45)
    torch.cuda.synchronize()
    n_pass = int(passed[0].sum())
    assert 290 <= n_pass <= 310
    assert int(res.candidates[0]) == n_pass and int(res.candidates[1]) == 0 and int(res.count[1]) == 0
    idx = torch.nonzero(passed[0]).flatten().cpu().numpy()
    want = ref_port.nms(bx[0].cpu().numpy()[idx], sc[0].cpu().numpy()[idx], 0.45)
    k = int(res.count[0])
    np.testing.assert_array_equal(res.keep[0, :k].cpu().numpy(), np.asarray(want))
    np.testing.assert_array_equal(res.anchor[0, :k].cpu().numpy(), idx[np.asarray(want)])


def test_bench_size_fused_equals_two_kernel_path(ov, cuda_device):
    """At BASELINE's full size (batch 256 @ 640^2, 1203 prompts: 2.15 M anchors, 16 800 CTA-pair
    tiles) the fused kernel and the K1 -> K2 path - independent code, same bf16 products - agree on
    every score to 2e-5 and on every class outside exact near-ties; and the detections of the
    full pipeline are the same lists."""
    from ovdet import ops, synth
    from ovdet.pipeline import HeadConfig, HeadPipeline
    shapes = [(80, 80), (40, 40), (20, 20)]
    inp = synth.make_inputs(batch=256, image_size=640, num_classes=1203, device=cuda_device, seed=1234)
    cfg = HeadConfig(precision="bf16", max_det=300)
    fused = HeadPipeline(256, shapes, 1203, cfg, device=cuda_device)
    split = HeadPipeline(256, shapes, 1203, HeadConfig(precision="bf16", max_det=300, fused=False), device=cuda_device)
    fused.set_vocabulary(inp.text)
    split.set_vocabulary(inp.text)
    rf = fused.run(inp.obj_embeds, inp.box_preds)
    rs = split.run(inp.obj_embeds, inp.box_preds)
    torch.cuda.synchronize()
    assert fused.last_path == "fused" and split.last_path == "split"
    assert (fused.scores - split.scores).abs().max().item() <= 2e-5
    same = fused.class_ids == split.class_ids
    assert same.float().mean().item() >= 0.9999
    # where the class differs the two scores are a near-tie
    assert (fused.scores[~same] - split.scores[~same]).abs().max().item() <= 2e-5 if (~same).any() else True
    assert torch.equal(fused.boxes, split.boxes)
    agree = 0
    for i in range(256):
        k = int(rf.count[i])
        if k == int(rs.count[i]) and torch.equal(rf.anchor[i, :k], rs.anchor[i, :k]):
            agree += 1
    assert agree >= 250          # a threshold-straddling score may flip a candidate in a few images
    assert int(rf.count.sum()) > 30_000


def test_nms_conf_equals_pass_mask(ov, cuda_device):
    """K4 thresholding the scores itself == K4 fed with the packed `scores > conf` mask (NaN and
    the exact-threshold value never pass); resident, multi-chunk and top-k / class-aware paths."""
    from ovdet import ops
    from ovdet.detector import _pack_mask
    rng = np.random.default_rng(11)
    b, a = 3, 8400
    boxes = torch.from_numpy(np.stack([_rand_boxes(rng, a, 640, 120) for _ in range(b)])).to(cuda_device)
    scores = torch.from_numpy(rng.uniform(-0.2, 1.0, (b, a)).astype(np.float32)).to(cuda_device)
    scores[0, 7] = float("nan")
    scores[0, 9] = 0.25
    scores[2] = 0.1                                               # no survivor
    classes = torch.from_numpy(rng.integers(0, 5, (b, a)).astype(np.int32)).to(cuda_device)
    for conf, kw in ((0.25, {}), (0.9, {}), (0.25, {"topk": 300}), (0.5, {"class_aware": True})):
        want = ops.nms_batched(boxes, scores, classes, _pack_mask(scores > conf), iou_thr=0.45, **kw)
        got = ops.nms_batched(boxes, scores, classes, conf=conf, iou_thr=0.45, **kw)
        torch.cuda.synchronize()
        assert torch.equal(got.count, want.count) and torch.equal(got.candidates, want.candidates)
        for i in range(b):
            k = int(want.count[i])
            for f in ("boxes", "scores", "classes", "anchor", "keep"):
                assert torch.equal(getattr(got, f)[i, :k], getattr(want, f)[i, :k]), (conf, kw, f)


@pytest.mark.parametrize("classes,batched,dim", [(80, False, 512), (128, True, 512), (17, False, 256), (5, True, 64)])
def test_similarity_fused_fp32_streaming(ov, cuda_device, classes, batched, dim):
    """The fused kernel's fp32-accurate mode for small vocabularies (BASELINE configs[1]: 80
    prompts): activation blocks stream through the 8-slot TMEM ring, three bf16 passes."""
    from ovdet import ops
    torch.manual_seed(classes + dim)
    b = 3
    shapes = [(20, 20), (10, 12), (4, 5)]
    embs = [torch.randn(b, dim, h, w) * (0.5 + l) for l, (h, w) in enumerate(shapes)]
    embs[2][1, :, 1, 1] = 0.0
    text = torch.randn(b, classes, dim) if batched else torch.randn(classes, dim).unsqueeze(0).expand(b, -1, -1)
    alpha, beta = 1.7, -0.2
    ref = torch.cat([ref_port.compute_similarity(e, text, alpha, beta).flatten(2).transpose(1, 2)
                     for e in embs], dim=1)
    dev_embs = [e.to(cuda_device) for e in embs]
    top3 = ops.text_operand_fp32(text.to(cuda_device) if batched else text[0].to(cuda_device))
    assert top3.shape[-1] == 3 * dim
    logits, rmax, rarg = ops.similarity_fused(dev_embs, top3, alpha, beta, logits_dtype=torch.float32,
                                              want_max=True, fp32=True)
    torch.cuda.synchronize()
    assert_logits_close(logits, ref, "fp32", alpha)            # the fp32 bar: 1e-3 relative (measured ~1e-5)
    assert (logits.cpu() - ref).abs().max().item() <= 3e-5 * alpha
    m, a = logits.max(dim=-1)
    assert torch.equal(rmax, m) and torch.equal(rarg.long(), a)
    _, m0, a0 = ops.similarity_fused(dev_embs, top3, alpha, beta, logits_dtype=None, want_max=True, fp32=True)
    assert torch.equal(m0, rmax)
    assert (a0 == rarg).float().mean().item() >= 0.999


FP16_ABS = 1e-4          # the fp16 tier: |dlogit| at alpha = 1 (measured max ~6e-5 over 10^7 logits, rms ~1e-5)


@pytest.mark.parametrize("classes,batched", [(80, False), (1203, False), (300, True)])
def test_similarity_fused_fp16_tier_vs_oracle(ov, cuda_device, classes, batched):
    """The fp16 tensor-core tier (one pass, fp16 operands, per-row power-of-two scaling): logits against
    the oracle inside the north_star's fp32 bar (1e-3 relative) and inside 1e-4 absolute - two orders
    below the bf16 tier - for activations of any magnitude: levels scaled by 1e-6 and 3e4, a zero
    vector, a row whose first 64 channels are all zero, ragged anchor tiles."""
    from ovdet import ops
    torch.manual_seed(classes)
    b, dim = 3, 512
    shapes = [(20, 20), (10, 12), (4, 5)]
    scales = [1.0, 1e-6, 3e4]
    embs = [torch.randn(b, dim, h, w) * sc for (h, w), sc in zip(shapes, scales)]
    embs[0][1, :, 2, 3] = 0.0                       # zero vector: eps clamp, score = beta
    embs[0][2, :64, 5, 5] = 0.0                     # first block all zero: scale falls back to 1
    embs[1][0, :8, 1, 1] *= 1e-3                    # sampled channels much smaller than the rest of the row
    text = torch.randn(b, classes, dim) if batched else torch.randn(classes, dim).unsqueeze(0).expand(b, -1, -1)
    alpha, beta = 1.0, 0.03
    ref = torch.cat([ref_port.compute_similarity(e, text, alpha, beta).flatten(2).transpose(1, 2)
                     for e in embs], dim=1)
    dev_embs = [e.to(cuda_device) for e in embs]
    top = ops.l2norm_text(text.to(cuda_device) if batched else text[0].to(cuda_device), split="fp16")
    assert top.dtype == torch.float16 and top.shape[-1] == dim
    logits, rmax, rarg = ops.similarity_fused(dev_embs, top, alpha, beta, logits_dtype=torch.float32, want_max=True)
    torch.cuda.synchronize()
    err = (logits.cpu() - ref).abs()
    print(f"\n[fp16 tier, C={classes}] max |dlogit| {err.max():.2e}, rms {err.pow(2).mean().sqrt():.2e}")
    assert err.max() / ref.abs().max() <= FP32_REL          # north_star: within 1e-3 relative
    assert err.max() <= FP16_ABS
    m, a = logits.max(dim=-1)
    assert torch.equal(rmax, m) and torch.equal(rarg.long(), a)
    assert (rarg.cpu().long() == ref.argmax(dim=-1)).float().mean() >= 0.995
    _, m0, a0 = ops.similarity_fused(dev_embs, top, alpha, beta, logits_dtype=None, want_max=True)
    assert (m0 - rmax).abs().max().item() <= 1e-6 and (a0 == rarg).float().mean().item() >= 0.999
    # bf16 logits out of the fp16 product (TMA-store epilogue, two epilogue groups)
    lb = torch.empty(b, logits.shape[1], (classes + 7) // 8 * 8, device=cuda_device, dtype=torch.bfloat16)[..., :classes]
    ops.similarity_fused(dev_embs, top, alpha, beta, logits=lb, want_max=True)
    torch.cuda.synchronize()
    assert (lb.float() - logits).abs().max().item() <= 4e-3


def test_pipeline_fp16_tier_bench_shape_vs_oracle(ov, cuda_device):
    """precision="fp16" through HeadPipeline at the bench's per-image shape (640^2, 1203 prompts): the
    single C call and the per-stage launches agree byte for byte, scores inside 1e-4 of the oracle,
    classes equal outside near-ties, post-processing bit-exact on identical inputs, kept sets reported."""
    from ovdet import synth
    from ovdet.pipeline import HeadConfig, HeadPipeline
    b, classes, s = 2, 1203, 640
    inp = synth.make_inputs(batch=b, image_size=s, num_classes=classes, seed=31)
    tail = ref_port.head_tail(inp.obj_embeds, inp.text_batched(), inp.box_preds)
    shapes = [(s // 8, s // 8), (s // 16, s // 16), (s // 32, s // 32)]
    pipe = HeadPipeline(b, shapes, classes, HeadConfig(precision="fp16"), device=cuda_device)
    pipe.set_vocabulary(inp.text.to(cuda_device))
    sizes = [(s, s)] * b
    pipe.set_geometry(sizes, [1.0] * b)
    objs = [e.to(cuda_device) for e in inp.obj_embeds]
    preds = [p.to(cuda_device) for p in inp.box_preds]
    res = pipe.run(objs, preds)
    torch.cuda.synchronize()
    assert pipe.last_path == "fused" and pipe.last_single_call
    err = (pipe.scores.cpu() - tail["scores"]).abs().max().item()
    print(f"\n[fp16 tier pipeline] max |dscore| {err:.2e}")
    assert err <= FP16_ABS
    assert (pipe.class_ids.cpu().long() == tail["class_ids"]).float().mean() >= 0.998
    fed = {"boxes": pipe.boxes.cpu(), "scores": pipe.scores.cpu(), "class_ids": pipe.class_ids.cpu().long()}
    want = ref_port.postprocess_batch(fed, sizes, [1.0] * b)
    for i in range(b):
        k = int(res.count[i])
        assert k == len(want[i]["keep"]) and k > 20
        np.testing.assert_array_equal(res.keep[i, :k].cpu().numpy(), want[i]["keep"])
    mism, kept_ref = _index_mismatches(res, ref_port.postprocess_batch(tail, sizes, [1.0] * b))
    print(f"[e2e index mismatches, fp16 tier] per image {mism} of {kept_ref} kept")
    assert sum(mism) <= max(1, int(0.02 * sum(kept_ref)))
    scores = pipe.scores.clone()
    pipe.run(objs, preds, events={})
    torch.cuda.synchronize()
    assert torch.equal(pipe.scores, scores)
    with pytest.raises(ValueError, match="fp16"):
        HeadPipeline(b, shapes, classes, HeadConfig(precision="fp16", embed_dim=256), device=cuda_device)


def test_fused_modes_are_deterministic_under_load(ov, cuda_device):
    """Race detector of last resort (compute-sanitizer is closed on this pool): every mode of the
    fused kernel - CTA-pair cosine, projected, streaming fp32, attention - run 12 times on the same
    mid-size input (thousands of tiles per launch, persistent CTAs recycling TMEM / smem rings)
    must give bit-identical scores and classes every time."""
    from ovdet import ops, synth
    torch.manual_seed(5)
    dev = cuda_device
    inp = synth.make_inputs(batch=37, image_size=640, num_classes=1203, device=dev, seed=9)
    top = ops.l2norm_text(inp.text)
    pin = synth.make_projected_inputs(batch=37, image_size=640, num_classes=1203, device=dev, seed=9)
    lops = [ops.project_vocabulary(pin.text, w, b) for w, b in pin.projections()]
    inp80 = synth.make_inputs(batch=37, image_size=640, num_classes=80, device=dev, seed=9)
    top3 = ops.text_operand_fp32(inp80.text)
    y = torch.randn(16, 64, 80, 80, device=dev)
    tproj = torch.randn(1203, 64, device=dev)

    def runs():
        yield "cosine", lambda: ops.similarity_fused(inp.obj_embeds, top, want_max=True)[1:]
        yield "projected", lambda: ops.similarity_projected(pin.hidden, lops, 1203)
        yield "fp32_stream", lambda: ops.similarity_fused(inp80.obj_embeds, top3, want_max=True, fp32=True)[1:]
        yield "attention", lambda: ops.max_sigmoid_attention(y, tproj, precise=True, return_scores=True)

    for name, fn in runs():
        first = [t.clone() for t in fn()]
        for _ in range(11):
            again = fn()
            torch.cuda.synchronize()
            for a, b in zip(first, again):
                assert torch.equal(a, b), name


def test_predict_host_streams_equal_predict(ov, cuda_device):
    """Detector.predict_host (pinned host buffers, chunked H2D on a copy stream overlapped with the
    kernels on a compute stream, D2H of every chunk's detections) returns what predict returns for
    the same images - with and without the folded projection."""
    from ovdet import synth
    from ovdet.detector import Detector
    from ovdet.pipeline import HeadConfig
    cfg = HeadConfig(precision="bf16", max_det=64)
    inp = synth.make_inputs(batch=6, image_size=256, num_classes=90, seed=41)
    pin = synth.make_projected_inputs(batch=6, image_size=256, num_classes=90, seed=41)
    for embeds, preds, text, proj in ((inp.obj_embeds, inp.box_preds, inp.text, None),
                                      (pin.hidden, pin.box_preds, pin.text, pin.projections())):
        det = Detector(device=str(cuda_device), config=cfg)
        det.set_vocabulary(text.to(cuda_device))
        dproj = None if proj is None else [(w.to(cuda_device), b.to(cuda_device)) for w, b in proj]
        want = det.predict([e.to(cuda_device) for e in embeds], [p.to(cuda_device) for p in preds],
                           text.to(cuda_device), projections=dproj)
        torch.cuda.synchronize()
        want = {k: getattr(want, k).cpu().clone() for k in ("boxes", "scores", "classes", "count")}
        host_e = [e.pin_memory() for e in embeds]
        host_p = [p.pin_memory() for p in preds]
        for _ in range(2):                       # second call reuses the staging buffers / streams
            got = det.predict_host(host_e, host_p, chunk=2, projections=dproj)
        assert torch.equal(got["count"], want["count"]) and int(want["count"].sum()) > 0
        for i in range(6):
            k = int(want["count"][i])
            for key in ("boxes", "scores", "classes"):
                assert torch.equal(got[key][i, :k], want[key][i, :k]), key


@pytest.mark.parametrize("classes,batched,dim", [(1203, False, 512), (80, True, 512), (33, False, 128)])
def test_similarity_fused_bf16_activations(ov, cuda_device, classes, batched, dim):
    """bf16 conv outputs (autocast) straight into the fused kernel: same bar as the bf16 path,
    against the oracle fed with the same (exactly representable) values in fp32."""
    from ovdet import ops
    torch.manual_seed(classes)
    b = 2
    shapes = [(24, 24), (16, 8), (4, 4)]                    # 576 | 128 | 16 anchors, H*W multiples of 8
    embs16 = [(torch.randn(b, dim, h, w) * (0.5 + l)).to(torch.bfloat16) for l, (h, w) in enumerate(shapes)]
    text = torch.randn(b, classes, dim) if batched else torch.randn(classes, dim).unsqueeze(0).expand(b, -1, -1)
    ref = torch.cat([ref_port.compute_similarity(e.float(), text, 1.2, 0.1).flatten(2).transpose(1, 2)
                     for e in embs16], dim=1)
    dev = [e.to(cuda_device) for e in embs16]
    assert ops.fused_supported(dev)
    top = ops.l2norm_text(text.to(cuda_device) if batched else text[0].to(cuda_device))
    logits, rmax, rarg = ops.similarity_fused(dev, top, 1.2, 0.1, logits_dtype=torch.float32, want_max=True)
    torch.cuda.synchronize()
    assert_logits_close(logits, ref, "bf16", 1.2)
    m, a = logits.max(dim=-1)
    assert torch.equal(rmax, m) and torch.equal(rarg.long(), a)
    # the activations are already bf16, so the only rounding left is the text operand's: the fp32-input
    # kernel fed with the widened values multiplies the same bf16 products
    l32, _, _ = ops.similarity_fused([e.float() for e in dev], top, 1.2, 0.1, logits_dtype=torch.float32)
    assert (logits - l32).abs().max().item() <= 2e-6


def test_auto_precision_with_autocast_activations(ov, cuda_device):
    """The default configuration (precision="auto" -> fp16 tier) meets bf16 conv outputs (heads under
    autocast): the pipeline hands the call to its bf16-operand twin instead of raising; scores inside the bf16
    bar against the oracle on the same (exactly representable) values, post-processing bit-exact, and the
    same pipeline object keeps serving fp32 activations through the fp16 tier."""
    from ovdet import synth
    from ovdet.pipeline import HeadConfig, HeadPipeline
    b, classes, s = 2, 300, 256
    shapes = [(s // 8, s // 8), (s // 16, s // 16), (s // 32, s // 32)]
    inp = synth.make_inputs(batch=b, image_size=s, num_classes=classes, seed=33)
    embs16 = [e.to(torch.bfloat16) for e in inp.obj_embeds]
    preds16 = [p.to(torch.bfloat16) for p in inp.box_preds]
    tail = ref_port.head_tail([e.float() for e in embs16], inp.text_batched(), [p.float() for p in preds16])
    pipe = HeadPipeline(b, shapes, classes, HeadConfig(), device=cuda_device)
    assert pipe.cfg.precision == "fp16"
    pipe.set_vocabulary(inp.text.to(cuda_device))
    sizes = [(s, s)] * b
    pipe.set_geometry(sizes, [1.0] * b)
    res = pipe.run([e.to(cuda_device) for e in embs16], [p.to(cuda_device) for p in preds16])
    torch.cuda.synchronize()
    assert pipe.last_path == "fused"
    assert_logits_close(pipe.scores, tail["scores"], "bf16")
    fed = {"boxes": pipe.boxes.cpu(), "scores": pipe.scores.cpu(), "class_ids": pipe.class_ids.cpu().long()}
    want = ref_port.postprocess_batch(fed, sizes, [1.0] * b)
    for i in range(b):
        k = int(res.count[i])
        assert k == len(want[i]["keep"]) and k > 3
        np.testing.assert_array_equal(res.keep[i, :k].cpu().numpy(), want[i]["keep"])
    # fp32 activations afterwards: the fp16 tier again, inside the fp32 bar
    tail32 = ref_port.head_tail(inp.obj_embeds, inp.text_batched(), inp.box_preds)
    pipe.run([e.to(cuda_device) for e in inp.obj_embeds], [p.to(cuda_device) for p in inp.box_preds])
    torch.cuda.synchronize()
    assert (pipe.scores.cpu() - tail32["scores"]).abs().max().item() <= 1e-4


@pytest.mark.parametrize("batch,classes", [(1, 1203), (1, 300), (2, 80)])
def test_small_launch_class_split(ov, cuda_device, batch, classes):
    """Batch 1 at 640^2 fills 34 of the 74 CTA pairs: ovdet_head_step splits the class tiles of every
    anchor tile over the idle pairs and the last arriver merges the partial (max, argmax).  Same
    scores and classes as the unsplit kernel (per-stage path), launch after launch (the arrival
    counters are left zero)."""
    from ovdet import synth
    from ovdet.pipeline import HeadConfig, HeadPipeline
    shapes = [(80, 80), (40, 40), (20, 20)]
    inp = synth.make_inputs(batch=batch, image_size=640, num_classes=classes, device=cuda_device, seed=3)
    pipe = HeadPipeline(batch, shapes, classes, HeadConfig(precision="bf16", max_det=300), device=cuda_device)
    pipe.set_vocabulary(inp.text)
    pipe.run(inp.obj_embeds, inp.box_preds, events={})                  # per-stage launches: unsplit kernel
    torch.cuda.synchronize()
    want_s, want_c, want_n = pipe.scores.clone(), pipe.class_ids.clone(), pipe.result.count.clone()
    for _ in range(3):
        pipe.scores.fill_(-7.0)
        pipe.run(inp.obj_embeds, inp.box_preds)                         # one C call: split when it pays
        torch.cuda.synchronize()
        assert torch.equal(pipe.scores, want_s)
        assert (pipe.class_ids != want_c).sum().item() <= 2             # exact ties after the affine map only
        assert torch.equal(pipe.result.count, want_n)
    if batch == 1:
        # the split must actually be in use at batch 1, and the contract of ovdet_similarity_fused_ws -
        # "arrival counters zero on entry, left zero" - holds: the counter block heads the workspace
        ws = getattr(pipe, "_sim_ws", None)
        assert ws is not None and ws.numel() > 0
        tiles = sum((h * w + 127) // 128 for h, w in shapes) * batch
        counters = ws[:(tiles + 2) * 16].view(torch.int32)
        assert int(counters.abs().sum()) == 0


def test_decode_bf16_box_logits(ov, cuda_device):
    """bf16 box logits (head under autocast): decode == the oracle on the same values widened to
    fp32, for the 16-byte-load path, the hoisted small-launch path and the generic path."""
    from ovdet import ops, synth
    for batch, size, strides in ((40, 640, (8, 16, 32)), (2, 320, (8, 16, 32)), (2, 160, (8, 16, 32))):
        inp = synth.make_inputs(batch=batch, image_size=size, num_classes=10, embed_dim=64, seed=3)
        preds16 = [p.to(torch.bfloat16) for p in inp.box_preds]
        grids = [ref_port.create_grid(batch, p.shape[2], p.shape[3], s) for p, s in zip(preds16, strides)]
        ref = ref_port.decode_boxes([p.float() for p in preds16], grids)
        scores = torch.rand(batch, ref.shape[1]) - 0.3
        boxes, _, mask = ops.decode_filter([p.to(cuda_device) for p in preds16], strides,
                                           scores=scores.to(cuda_device), conf=0.25)
        torch.cuda.synchronize()
        torch.testing.assert_close(boxes.cpu(), ref, rtol=1e-4, atol=1e-3)
        assert torch.equal(_unpack(mask.cpu(), ref.shape[1]), scores > 0.25)


def test_soak_fused_protocol_small(ov, cuda_device, monkeypatch, capsys):
    """tools/soak_fused.py for a few iterations: random shapes, every fused launch against the two-kernel
    path (2e-5) and the fp16 tier against fp64 torch (1e-4) - the check that a change to the converter /
    MMA / epilogue hand-shakes did not open a race."""
    import runpy
    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "soak_fused.py")
    monkeypatch.setattr(sys, "argv", [tool, "30", "11"])
    runpy.run_path(tool, run_name="__main__")
    assert '"ok": true' in capsys.readouterr().out

