"""2+ GPU check of the batch-sharded path (run under torchrun, NCCL): every rank runs
Detector.predict on its slice of a seeded global batch, rank 0 gathers the detections and compares
them with its own un-sharded run of the whole batch - byte for byte.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/shard_check.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from ovdet import shard, synth
from ovdet.detector import Detector
from ovdet.pipeline import HeadConfig

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
TOTAL, CLASSES, MAX_DET = 7, 300, 128                       # ragged: 7 images over `world` ranks
inp = synth.make_inputs(batch=TOTAL, image_size=320, num_classes=CLASSES, seed=21, device=dev)
vocab = shard.broadcast_vocabulary(inp.text if rank == 0 else None, CLASSES, 512, dev)
det = Detector(device=str(dev), config=HeadConfig(precision="bf16", max_det=MAX_DET))


def compute(*tensors):
    embeds, preds = list(tensors[:3]), list(tensors[3:])
    n = embeds[0].shape[0]
    if n == 0:
        z = lambda *s, dt=torch.float32: torch.zeros(*s, device=dev, dtype=dt)
        return {"boxes": z(0, MAX_DET, 4), "scores": z(0, MAX_DET), "classes": z(0, MAX_DET, dt=torch.int32),
                "count": z(0, dt=torch.int32)}
    res = det.predict([e.contiguous() for e in embeds], [p.contiguous() for p in preds], vocab)
    torch.cuda.synchronize()
    return {"boxes": res.boxes.clone(), "scores": res.scores.clone(), "classes": res.classes.clone(),
            "count": res.count.clone()}


got = shard.run_sharded(compute, inp.obj_embeds + inp.box_preds)
if rank == 0:
    want = compute(*(inp.obj_embeds + inp.box_preds))
    assert int(want["count"].sum()) > 0
    assert torch.equal(got["count"], want["count"])
    for i in range(TOTAL):
        k = int(want["count"][i])
        for key in ("boxes", "scores", "classes"):
            assert torch.equal(got[key][i, :k], want[key][i, :k]), (key, i)
    print(f"shard check ok: world {world}, {TOTAL} images, shards {shard.shard_sizes(TOTAL, world)}, "
          f"{int(want['count'].sum())} detections identical to the un-sharded run")
dist.destroy_process_group()
