"""CPU tests of the vocabulary-parallel host logic (ovdet/vocab_parallel.py, SURVEY.md section
8 e / f-4): the class partition, the (score, class) key and, with gloo at world_size 2 and 3, that
an all-reduce(MAX) over the keys of class shards equals the reference's max / argmax over the
whole vocabulary (model/yolo_clip.py:198-206, lowest class index on ties).  The kernels that
produce and consume the keys on the GPU are checked against the same numpy twin in
tests/test_gpu_kernels.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ovdet import vocab_parallel as vp


def test_class_range_is_a_balanced_partition():
    for classes in (1, 2, 7, 80, 1203, 4800):
        for world in (1, 2, 3, 4, 8):
            if classes < world:
                with pytest.raises(ValueError):
                    vp.class_range(0, world, classes)
                continue
            spans = [vp.class_range(r, world, classes) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == classes
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert min(sizes) >= 1 and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        vp.class_range(2, 2, 10)


def test_key_roundtrip_and_order():
    rng = np.random.default_rng(0)
    special = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, 1e-45, -1e-45, 3.4e38, -3.4e38,
                        0.25, np.nextafter(np.float32(0.25), np.float32(1))], dtype=np.float32)
    scores = np.concatenate([special, rng.standard_normal(500).astype(np.float32)])
    classes = rng.integers(0, 2 ** 31 - 1, size=scores.shape[0]).astype(np.int64)
    keys = vp.pack_keys_host(scores, classes)
    s2, c2 = vp.unpack_keys_host(keys)
    assert np.array_equal(s2, scores + np.float32(0.0)) and np.array_equal(c2.astype(np.int64), classes)
    # signed key order == (score ascending, class descending)
    order = np.argsort(keys, kind="stable")
    ks, kc = scores[order] + np.float32(0.0), classes[order]
    for i in range(len(order) - 1):
        assert ks[i] < ks[i + 1] or (ks[i] == ks[i + 1] and kc[i] >= kc[i + 1])
    # -0.0 and +0.0 are the same score: the lower class wins
    k = vp.pack_keys_host(np.array([-0.0, 0.0], np.float32), np.array([3, 5]))
    assert k[0] > k[1]


def test_key_max_equals_reference_max_argmax():
    """max over the keys of every class == torch.max over the class axis (values, lowest index on
    ties - the rule SURVEY 8 a-6 probed on the reference's CPU path)."""
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(64, 37, generator=g)
    logits[:, 5] = logits[:, 20]                  # exact ties between two classes
    logits[3] = 0.125                             # a whole row of ties
    want_v, want_i = logits.max(dim=1)
    keys = np.stack([vp.pack_keys_host(logits[:, c].numpy(), np.full(64, c)) for c in range(37)], axis=1)
    v, i = vp.unpack_keys_host(keys.max(axis=1))
    assert np.array_equal(v, want_v.numpy())
    # torch.max returns the first maximal index on CPU
    first = (logits == want_v[:, None]).float().argmax(dim=1)
    assert np.array_equal(i, first.numpy().astype(np.int32))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, classes, path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(11)
        logits = torch.randn(3, 50, classes, generator=g)      # [B, A, C], the same on every rank
        logits[:, :, classes - 1] = logits[:, :, 0]              # ties across the first and last shard
        c0, c1 = vp.class_range(rank, world, classes)
        local_v, local_i = logits[..., c0:c1].max(dim=-1)        # this rank's shard (yolo_clip.py:200)
        keys = torch.from_numpy(vp.pack_keys_host(local_v.numpy(), local_i.numpy(), class_offset=c0))
        dist.all_reduce(keys, op=dist.ReduceOp.MAX)
        v, i = vp.unpack_keys_host(keys.numpy())
        if rank == world - 1:
            torch.save({"v": torch.from_numpy(v.copy()), "i": torch.from_numpy(i.copy())}, path)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,classes", [(2, 80), (3, 7), (2, 3)])
def test_allreduce_of_keys_equals_full_vocabulary(tmp_path, world, classes):
    path = str(tmp_path / "merged.pt")
    mp.spawn(_worker, args=(world, _free_port(), classes, path), nprocs=world, join=True)
    got = torch.load(path)
    g = torch.Generator().manual_seed(11)
    logits = torch.randn(3, 50, classes, generator=g)
    logits[:, :, classes - 1] = logits[:, :, 0]
    want_v, _ = logits.max(dim=-1)
    first = (logits == want_v[..., None]).float().argmax(dim=-1)
    assert torch.equal(got["v"].reshape(3, 50), want_v)
    assert torch.equal(got["i"].reshape(3, 50).long(), first)
    assert int((first == 0).sum()) > 0            # the tie rows resolved to class 0, not C - 1
