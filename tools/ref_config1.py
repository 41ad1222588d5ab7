"""BASELINE.json configs[0] / BASELINE.md section 4 item 1: the LIVE reference's own
``YOLOCLIPDetector.detect`` on one synthetic 640x640 image with the 80 COCO prompts, random-init
``n`` weights, CPU.  p50 over >= 20 calls after 3 warm-ups, core and thread count stated.

Build-container tool (``/root/reference`` does not exist on the GPU box); TEST/MEASUREMENT
INFRASTRUCTURE, imports the reference with the ``clip`` stub of ``oracle/clip_stub.py`` exactly as
``oracle/make_golden.py`` does.  The line it prints is committed under ``profiles/``:

    python tools/ref_config1.py > profiles/r2_ref_config1_cpu.json

The same image then goes through this repository's ``preprocess -> tail -> records`` only on a GPU box
(``tests/test_boundary.py``); here nothing of the product is timed.
"""
from __future__ import annotations

import json
import logging
import os
import statistics
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("OVDET_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import clip_stub  # noqa: E402

clip_stub.install()
logging.disable(logging.CRITICAL)

from yolo_clip_detector.inference.detector import YOLOCLIPDetector  # noqa: E402
from yolo_clip_detector.model.yolo_clip import YOLOCLIP  # noqa: E402


def coco_names():
    """The 80 names of config/default_config.py:96-107 (read from the reference's own config)."""
    from yolo_clip_detector.config import default_config as dc
    for v in vars(dc).values():
        if isinstance(v, (list, tuple)) and len(v) == 80 and all(isinstance(s, str) for s in v):
            return list(v)
    for v in vars(dc).values():                      # nested config objects / dicts
        for holder in (getattr(v, "__dict__", None), v if isinstance(v, dict) else None):
            if not holder:
                continue
            for w in holder.values():
                if isinstance(w, (list, tuple)) and len(w) == 80 and all(isinstance(s, str) for s in w):
                    return list(w)
    return [f"class {i}" for i in range(80)]


def main(calls: int = 30, warmup: int = 3) -> None:
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    names = coco_names()
    with tempfile.TemporaryDirectory() as tmp:
        ckpt = os.path.join(tmp, "random_init_n.pt")
        model = YOLOCLIP(backbone_variant="n", num_classes=len(names), offline_mode=True)
        torch.save({"model_state_dict": model.state_dict()}, ckpt)
        det = YOLOCLIPDetector(ckpt, class_names=names, device="cpu", image_size=(640, 640),
                               conf_threshold=0.25, iou_threshold=0.45, backbone_variant="n")
    image = np.random.default_rng(0).integers(0, 256, (640, 640, 3), dtype=np.uint8)
    stage = {}

    def timed(name, fn):
        def wrapper(*a, **k):
            t0 = time.perf_counter()
            out = fn(*a, **k)
            stage.setdefault(name, []).append(time.perf_counter() - t0)
            return out
        return wrapper

    det.preprocess_image = timed("preprocess_image", det.preprocess_image)
    det.postprocess_detections = timed("postprocess_detections", det.postprocess_detections)
    det.model.forward = timed("model.forward", det.model.forward)
    times, n_det = [], 0
    for i in range(warmup + calls):
        t0 = time.perf_counter()
        out = det.detect(image)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            n_det = len(out)
    for k in stage:
        stage[k] = stage[k][warmup:]
    p50 = statistics.median(times)
    print(json.dumps({
        "what": "BASELINE.json configs[0]: live reference YOLOCLIPDetector.detect, 1 synthetic 640x640 uint8 image, "
                "80 COCO prompts (offline vocabulary through the clip stub), random-init 'n' weights, CPU",
        "p50_s": p50, "min_s": min(times), "max_s": max(times), "images_per_s_p50": 1.0 / p50,
        "calls": calls, "warmup": warmup, "detections_last_call": n_det,
        "stage_p50_s": {k: statistics.median(v) for k, v in stage.items()},
        "cores": os.cpu_count(), "torch_threads": torch.get_num_threads(), "torch": torch.__version__,
        "numpy": np.__version__, "host": "build container (no GPU); the GPU box has no /root/reference",
        "reference": "inference/detector.py:289-325 (detect), model/yolo_clip.py:102-223 (forward)",
    }))


if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:3]))
