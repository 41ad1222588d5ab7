"""The CPU oracle (oracle/ref_port.py) against fixtures generated from the LIVE reference
(oracle/make_golden.py).  Float stages use a tight tolerance because torch's CPU kernels pick
different SIMD paths on different hosts; the numpy post-processing is bit-exact."""
import os

import numpy as np
import torch

from oracle import ref_port


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def test_similarity_batched(golden_dir):
    g = _load(golden_dir, "sim_batched_d512")
    sim = ref_port.compute_similarity(torch.from_numpy(g["obj"]), torch.from_numpy(g["text"]),
                                      float(g["alpha"]), float(g["beta"]))
    assert tuple(sim.stride()) == tuple(g["strides"])          # memory is [B,HW,C]
    np.testing.assert_allclose(sim.contiguous().numpy(), g["sim"], rtol=0, atol=2e-6)


def test_similarity_shared_affine(golden_dir):
    g = _load(golden_dir, "sim_shared_affine_d64")
    sim = ref_port.compute_similarity(torch.from_numpy(g["obj"]), torch.from_numpy(g["text"]),
                                      float(g["alpha"]), float(g["beta"]))
    assert tuple(sim.stride()) == tuple(g["strides"])
    np.testing.assert_allclose(sim.contiguous().numpy(), g["sim"], rtol=0, atol=2e-6)


def test_decode(golden_dir):
    g = _load(golden_dir, "decode_3level")
    for keys, want in ((("p0", "p1", "p2"), "boxes"), (("n0", "n1", "n2"), "boxes_noise")):
        preds = [torch.from_numpy(g[k]) for k in keys]
        grids = [ref_port.create_grid(p.shape[0], p.shape[2], p.shape[3], s)
                 for p, s in zip(preds, ref_port.STRIDES)]
        boxes = ref_port.decode_boxes(preds, grids).numpy()
        np.testing.assert_allclose(boxes, g[want], rtol=2e-6, atol=1e-5)
    np.testing.assert_array_equal(
        ref_port.create_grid(2, 8, 8, 8).contiguous().numpy(), g["grid0"])


def test_nms_and_iou_bit_exact(golden_dir):
    g = _load(golden_dir, "nms_cases")
    names = sorted({k[:-len("_boxes")] for k in g.files if k.endswith("_boxes")})
    assert "empty" in names and "degenerate64" in names
    for name in names:
        boxes, scores, thr = g[name + "_boxes"], g[name + "_scores"], float(g[name + "_thr"])
        for stable in (False, True):        # scores are pairwise distinct -> same result
            keep = ref_port.nms(boxes.copy(), scores.copy(), thr, stable_ties=stable)
            np.testing.assert_array_equal(np.asarray(keep, dtype=np.int64), g[name + "_keep"])
        if name + "_iou0" in g.files:
            iou = ref_port.compute_iou(boxes[0], boxes[1:])
            assert iou.dtype == np.float32
            np.testing.assert_array_equal(iou, g[name + "_iou0"])
    # tie rule: equal scores -> higher original index first
    tie = np.array([.5, .7, .5, .7, .1], np.float32)
    np.testing.assert_array_equal(np.argsort(tie, kind="stable")[::-1], g["tie_order"])


def test_postprocess(golden_dir):
    g = _load(golden_dir, "postprocess_b3")
    names = [f"thing{i}" for i in range(40)]
    for i in range(3):
        res = ref_port.postprocess_image(g["boxes"][i], g["scores"][i], g["class_ids"][i],
                                         tuple(int(v) for v in g[f"img{i}_orig"]),
                                         float(g[f"img{i}_scale"]), class_names=names)
        dets = res["detections"]
        assert len(dets) == len(g[f"img{i}_score"]) > 0
        np.testing.assert_array_equal(np.array([d["box"] for d in dets]), g[f"img{i}_box"])
        np.testing.assert_array_equal(np.array([d["score"] for d in dets]), g[f"img{i}_score"])
        np.testing.assert_array_equal(np.array([d["class_id"] for d in dets]), g[f"img{i}_class"])
        assert dets[0]["class_name"] == str(g[f"img{i}_name0"])
        # the survivor -> anchor map is consistent with the thresholded arrays
        assert np.all(g["scores"][i][res["anchor_idx"]] == res["scores"])


def test_forward_tail(golden_dir):
    g = _load(golden_dir, "forward_tail_64")
    objs = [torch.from_numpy(g[f"obj{i}"]) for i in range(3)]
    boxes_in = [torch.from_numpy(g[f"box{i}"]) for i in range(3)]
    out = ref_port.head_tail(objs, torch.from_numpy(g["text"]), boxes_in)
    np.testing.assert_allclose(out["boxes"].numpy(), g["boxes"], rtol=2e-6, atol=1e-3)
    np.testing.assert_allclose(out["scores"].numpy(), g["scores"], rtol=0, atol=2e-6)
    agree = (out["class_ids"].numpy() == g["class_ids"]).mean()
    assert agree == 1.0
    assert set(g["keys"]) == {"boxes", "scores", "class_ids", "obj_embeddings",
                              "text_embeddings", "box_preds"}
    assert out["boxes"].shape == (2, 84, 4)        # 64 | 16 | 4 anchors: P3 | P4 | P5


# ------------------------------------------------------------------------------------------
# "next" rows: pre-processing and vocabulary format
# ------------------------------------------------------------------------------------------
def test_preprocess_matches_live_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "preprocess_cases.npz"))
    for name in ("up", "down", "half", "same", "wide"):
        h, w, scale = g[f"{name}_meta"]
        out, orig, s = ref_port.preprocess_image(g[f"{name}_img"], (int(h), int(w)))
        assert s == scale
        np.testing.assert_array_equal(out.numpy(), g[f"{name}_out"])
        np.testing.assert_array_equal(orig, g[f"{name}_img"])


def test_vocabulary_format(golden_dir, tmp_path):
    g = np.load(os.path.join(golden_dir, "vocab_3cls.npz"))
    names, matrix = ref_port.load_offline_vocabulary(os.path.join(golden_dir, "vocab_3cls.json"))
    assert names == list(g["names"])
    np.testing.assert_array_equal(matrix.numpy(), g["matrix"])
    # the product's reader / writer speak the same format (pure host code, no GPU needed)
    from ovdet.vocabulary import Vocabulary, load_offline_vocabulary
    names2, matrix2 = load_offline_vocabulary(os.path.join(golden_dir, "vocab_3cls.json"))
    assert names2 == names and torch.equal(matrix2, matrix)
    out = str(tmp_path / "sub" / "v.json")
    Vocabulary(names2, matrix2).save(out)
    with open(out) as a, open(os.path.join(golden_dir, "vocab_3cls.json")) as b:
        assert a.read() == b.read()                      # byte-identical to the reference's file
    ref_out = str(tmp_path / "ref.json")
    ref_port.save_offline_vocabulary(ref_out, names, matrix)
    with open(ref_out) as a, open(out) as b:
        assert a.read() == b.read()


def test_tcsp_attention_restatement(golden_dir):
    """Oracle attention (repvl_pan.py:80-95) inside the reference layer's own convolutions
    reproduces the live reference output; the drop-in layer loads the reference state dict."""
    from ovdet.neck import TextGuidedCSPLayer
    g = np.load(os.path.join(golden_dir, "tcsp_layer.npz"))
    layer = TextGuidedCSPLayer(48, 64, 512, n_bottlenecks=1).eval()
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}
    layer.load_state_dict(sd, strict=True)
    x, text = torch.from_numpy(g["x"]), torch.from_numpy(g["text"])
    with torch.no_grad():
        y_temp = layer.bottlenecks[0](layer.cv1(x))
        proj = layer.text_proj(text)
        torch.testing.assert_close(y_temp, torch.from_numpy(g["y_temp"]), rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(proj, torch.from_numpy(g["proj"]), rtol=1e-5, atol=1e-5)
        y1, _ = ref_port.max_sigmoid_attention(torch.from_numpy(g["y_temp"]), torch.from_numpy(g["proj"]))
        out = layer.cv3(torch.cat((y1, layer.cv2(x)), dim=1))
    torch.testing.assert_close(out, torch.from_numpy(g["out"]), rtol=1e-5, atol=1e-5)


def test_head_projection_restatement(golden_dir):
    """oracle.project_similarity_max == live reference head: 1x1 conv on the hidden features,
    compute_similarity, class max (text_contrastive.py:67,112,119-153; yolo_clip.py:198-202)."""
    g = _load(golden_dir, "head_projection")
    alpha, beta = (float(v) for v in g["alpha_beta"])
    scores, ids, embeds = ref_port.project_similarity_max(
        [torch.from_numpy(g["hidden"])], [torch.from_numpy(g["weight"])], [torch.from_numpy(g["bias"])],
        torch.from_numpy(g["text"]), alpha, beta)
    torch.testing.assert_close(embeds[0], torch.from_numpy(g["obj_embed"]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(scores, torch.from_numpy(g["scores"]), rtol=1e-5, atol=2e-6)
    assert torch.equal(ids, torch.from_numpy(g["class_ids"]))
