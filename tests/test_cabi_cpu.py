"""CPU-side checks: the C-ABI library builds, loads and exports exactly what include/ovdet.h
declares; argument validation and the no-fallback contract hold without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def handle():
    from ovdet import build, _cabi
    build.build()
    return _cabi.lib()


def _declared():
    text = open(os.path.join(ROOT, "include", "ovdet.h")).read()
    return sorted(set(re.findall(r"OVDET_API[^;]*?\b(ovdet_\w+)\s*\(", text)))


def test_header_symbols_exported(handle):
    from ovdet import _cabi
    names = _declared()
    assert len(names) >= 11
    for n in names:
        assert hasattr(handle, n), f"{n} declared in ovdet.h but not exported"
    assert sorted(_cabi.PROTOTYPES) == names          # the ctypes table covers the whole header


def test_version_and_strerror(handle):
    assert handle.ovdet_version() == 100
    assert handle.ovdet_strerror(0) == b"ok"
    assert b"no fallback" in handle.ovdet_strerror(-3)
    assert handle.ovdet_strerror(-99) == b"unknown status"


def test_argument_validation_without_gpu(handle):
    # argument checks come before the device check, so they are testable on a CPU-only host
    assert handle.ovdet_rowmax(None, 0, 1, 1, 1, None, None, None) == -1
    assert handle.ovdet_nms_workspace_bytes(0, 100) == 0
    assert handle.ovdet_nms_workspace_bytes(2, 8400) >= 2 * (16384 * 8 + 8400 * 20)
    assert handle.ovdet_similarity(None, None, None, 1, 1, 1, 64, 0, 0, 1.0, 0.0, None, 0, 1,
                                   None, None, None) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only contract")
def test_no_cpu_fallback(handle):
    """Without a B200 every compute entry refuses to run; nothing silently falls back."""
    from ovdet import ops
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.rowmax(torch.zeros(2, 3))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.l2norm_text(torch.zeros(2, 64))
    buf = (ctypes.c_float * 16)()
    out = (ctypes.c_float * 4)()
    rc = handle.ovdet_rowmax(ctypes.addressof(buf), 0, 4, 4, 4, ctypes.addressof(out), None, None)
    assert rc in (-3, -4)                 # wrong arch / no CUDA device: an error, never a result


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "real-time-zero-shot-open-vocabulary-object-detection-using-a-lightweight_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle-sized", ""), f"{f} mentions the oracle"


def test_pack_mask_roundtrip():
    from ovdet.detector import _pack_mask
    rng = np.random.default_rng(0)
    bits = torch.from_numpy(rng.random((3, 100)) > 0.5)
    packed = _pack_mask(bits)
    assert packed.dtype == torch.int32 and packed.shape == (3, 4)
    m = packed.to(torch.int64) & 0xffffffff
    un = ((m.unsqueeze(-1) >> torch.arange(32)) & 1).reshape(3, -1)[:, :100].bool()
    assert torch.equal(un, bits)


def test_module_state_dict_keys_match_reference_layout():
    from ovdet.heads import BoxHead, TextContrastiveHead
    head = TextContrastiveHead(64)
    keys = set(head.state_dict())
    assert "obj_embed_conv.0.conv.weight" in keys and "obj_embed_conv.2.bias" in keys
    assert "box_conv.1.bn.running_mean" in keys
    assert not any("precision" in k for k in keys)
    box = BoxHead([64, 128, 256])
    assert "box_convs.2.2.weight" in box.state_dict()
    grid = box._create_grid(2, 3, 4, 8, torch.device("cpu"))
    assert grid.shape == (2, 3, 4, 3) and grid.dtype == torch.int64
    assert grid[0, 1, 2].tolist() == [2, 1, 8]


def test_new_entries_validate_arguments(handle):
    """Argument checks of the later entry points run before the device check (CPU-testable)."""
    assert handle.ovdet_letterbox_u8(None, None, None, None, None, None, 1, 640, 640, None, None) == -1
    assert handle.ovdet_pack_boxes_i32(None, None, 1, 1, None, None) == -1
    assert handle.ovdet_cast_text(None, 1, 1, 64, 1, 1, None, 64, 0, None) == -1
    assert handle.ovdet_max_sigmoid_attention(None, 1, 64, 16, 1, 1, None, 1, 0, 0, None, None, 1, 1, None) == -1
    assert handle.ovdet_similarity_projected(None, None, None, None, 1, 1, 256, None, 80, 0, 1.0, 0.0,
                                             None, None, None, None) == -1
    assert handle.ovdet_similarity_fused(None, None, None, None, 1, 1, 512, None, 80, 0, 1.0, 0.0, None, 0,
                                         80, None, None, None, None) == -1
    # vocabulary-parallel exchange
    assert handle.ovdet_vp_buffer_bytes(0, 2) == 0 and handle.ovdet_vp_buffer_bytes(100, 9) == 0
    assert handle.ovdet_vp_buffer_bytes(100, 2) == (100 * 16 + 127) // 128 * 128 + 128
    assert handle.ovdet_peer_buffer_create(0, None, None) == -1
    assert handle.ovdet_peer_buffer_open(None, None) == -1
    assert handle.ovdet_vp_buffer_init(None, 10, 2, None) == -1
    assert handle.ovdet_vp_signal(None, 2, 0, 10, None) == -1
    assert handle.ovdet_vp_wait_unpack(None, 2, 10, None, None, None, 0, None) == -1
    assert handle.ovdet_head_step_vp(None, 0, None, 2, 0, None, 0, None) == -1
    assert handle.ovdet_pack_score_keys(None, None, 1, 0, None, None) == -1
    assert handle.ovdet_unpack_score_keys(None, 1, None, None, None) == -1


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs) prints one JSON line with the
    contract's keys; it is the one place outside tests/ that may execute oracle/."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "port"
    assert "workload" in line["config"]


def test_head_step_args_mirror(handle):
    """The ctypes mirror of ovdet_head_step_args has the size the library was compiled with."""
    from ovdet import _cabi
    assert ctypes.sizeof(_cabi.HeadStepArgs) == handle.ovdet_head_step_args_size()
    assert handle.ovdet_head_step(None, None) == -1


def test_precision_auto_rule():
    """What precision="auto" resolves to (pipeline.resolve_precision): the fp16 operand tier for the
    reference's embed_dim 512 - for any class count and any level size (odd levels are re-pitched) - the
    three-pass recipe for other dims and for fp32 logits, bf16 for the projected mode."""
    from ovdet.pipeline import HeadConfig, resolve_precision
    shapes640, shapes416 = [(80, 80), (40, 40), (20, 20)], [(52, 52), (26, 26), (13, 13)]
    assert resolve_precision(HeadConfig(), shapes640, 1203) == "fp16"
    assert resolve_precision(HeadConfig(), shapes640, 80) == "fp16"
    assert resolve_precision(HeadConfig(), shapes416, 80) == "fp16"
    assert resolve_precision(HeadConfig(embed_dim=256), shapes640, 80) == "fp32"
    assert resolve_precision(HeadConfig(logits_dtype="fp32"), shapes640, 1203) == "fp32"
    assert resolve_precision(HeadConfig(fused=False), shapes640, 1203) == "fp32"
    assert resolve_precision(HeadConfig(), shapes640, 1203, projected=True) == "bf16"
    for explicit in ("bf16", "fp16", "fp32"):
        assert resolve_precision(HeadConfig(precision=explicit), shapes640, 1203) == explicit


def test_repitch_rows_validates_arguments(handle):
    """ovdet_repitch_rows without a GPU: argument errors come before the device check."""
    f = handle.ovdet_repitch_rows
    assert f(None, 4, 3, 3, None, 4, 4, None) == -1                 # null pointers
    buf = (ctypes.c_float * 16)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert f(p, 4, 5, 3, p, 8, 4, None) == -1                       # source pitch shorter than the row
    assert f(p, 4, 3, 3, p, 2, 4, None) == -1                       # destination pitch shorter than the row
    assert f(p, 4, 3, 3, p, 4, 3, None) == -1                       # element size must be 2 or 4
