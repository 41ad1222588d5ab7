"""Multi-process (world_size 2 and 3, gloo, CPU) tests of the batch-sharding host logic
(ovdet/shard.py).  The per-rank compute on a CPU-only host is the oracle
(tests may use it as the checker); on the B200s it is Detector.predict - the sharding code is
the same and contains no data-path collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_port
from ovdet import shard, synth

MAX_DET = 64


def test_shard_range_covers_batch_exactly():
    for total in (0, 1, 5, 8, 63, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [shard.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
            assert sizes == shard.shard_sizes(total, world)
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_compute(text):
    """Per-rank compute for the CPU test: reference tail + post-process of the local slice,
    packed into fixed-shape per-image tensors like ops.NmsResult."""
    def compute(e0, e1, e2, p0, p1, p2):
        n = e0.shape[0]
        out = {"boxes": torch.zeros(n, MAX_DET, 4), "scores": torch.zeros(n, MAX_DET),
               "classes": torch.zeros(n, MAX_DET, dtype=torch.int32),
               "count": torch.zeros(n, dtype=torch.int32)}
        if n == 0:
            return out
        tail = ref_port.head_tail([e0, e1, e2], text.unsqueeze(0).expand(n, -1, -1), [p0, p1, p2])
        res = ref_port.postprocess_batch(tail, [(160, 160)] * n, [1.0] * n)
        for i, r in enumerate(res):
            k = min(len(r["keep"]), MAX_DET)
            out["boxes"][i, :k] = torch.from_numpy(r["boxes"][:k])
            out["scores"][i, :k] = torch.from_numpy(r["scores"][:k])
            out["classes"][i, :k] = torch.from_numpy(r["class_ids"][:k].astype(np.int32))
            out["count"][i] = k
        return out
    return compute


def _worker(rank, world, port, total, path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        # every rank builds the same global batch (seeded); only its slice is processed
        inp = synth.make_inputs(batch=total, image_size=160, num_classes=40, seed=5)
        vocab = shard.broadcast_vocabulary(inp.text if rank == 0 else None, 40, 512, "cpu")
        assert torch.equal(vocab, inp.text)
        lo, hi = shard.shard_range(total, rank, world)
        local = shard.shard_batch(inp.obj_embeds + inp.box_preds)
        assert all(t.shape[0] == hi - lo for t in local)
        got = shard.run_sharded(_oracle_compute(vocab), inp.obj_embeds + inp.box_preds)
        slowest = shard.max_over_ranks(10.0 + rank, "cpu")
        assert slowest == 10.0 + world - 1
        if rank == 0:
            torch.save(got, path)
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 4), (2, 5), (3, 2)])
def test_sharded_run_equals_unsharded(tmp_path, world, total):
    """Ragged shards (5 over 2) and an empty shard (2 over 3) gather to exactly the un-sharded
    result, in global image order."""
    path = str(tmp_path / "gathered.pt")
    mp.spawn(_worker, args=(world, _free_port(), total, path), nprocs=world, join=True)
    got = torch.load(path)
    inp = synth.make_inputs(batch=total, image_size=160, num_classes=40, seed=5)
    want = _oracle_compute(inp.text)(*(inp.obj_embeds + inp.box_preds))
    assert int(want["count"].sum()) > 0
    for k in want:
        assert torch.equal(got[k], want[k]), k
