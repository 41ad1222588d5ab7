"""Host -> device copy bandwidth of this box: the ceiling of bench.py's e2e figure.

    python tools/pcie_probe.py                      # one GPU: pinned, write-combined and pageable sources
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29533 tools/pcie_probe.py --all-gpus     # every GPU copying at the same time

One JSON line.  With --all-gpus every rank binds to its own slice of the host cores (as bench.py does),
allocates after binding, and all ranks copy concurrently between two barriers: `per_gpu_GBps` is each
rank's own rate, `aggregate_GBps` the sum - the host-memory ceiling that the 8-GPU e2e figure runs into
(every pinned H2D byte is a host DRAM read; a virtualised single-NUMA-node host serves all eight links
from one memory system).  `solo_GBps` is rank 0 copying alone, for comparison.
Write-combined staging (cudaHostAllocWriteCombined) is measured through libcudart directly: torch has
no allocator for it.
"""
import argparse
import ctypes
import json
import os

import torch


def _rate(dst, src, reps=8):
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e.record()
    torch.cuda.synchronize()
    return reps * src.numel() * src.element_size() / (s.elapsed_time(e) * 1e-3) / 1e9


def _write_combined_rate(nbytes, dev, reps=8):
    """cudaHostAlloc(..., cudaHostAllocWriteCombined) -> cudaMemcpyAsync; None if libcudart is not loadable."""
    try:
        rt = ctypes.CDLL("libcudart.so")
    except OSError:
        try:
            import glob
            cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + \
                glob.glob("/usr/local/cuda/lib64/libcudart.so*")
            rt = ctypes.CDLL(cands[0])
        except (OSError, IndexError):
            return None
    ptr = ctypes.c_void_p()
    if rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04)) != 0:   # WriteCombined
        return None
    ctypes.memset(ptr, 1, nbytes)
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    for _ in range(2):
        rt.cudaMemcpyAsync(d.data_ptr(), ptr, nbytes, 1, stream)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        rt.cudaMemcpyAsync(d.data_ptr(), ptr, nbytes, 1, stream)
    e.record()
    torch.cuda.synchronize()
    rate = reps * nbytes / (s.elapsed_time(e) * 1e-3) / 1e9
    rt.cudaFreeHost(ptr)
    return rate


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--all-gpus", action="store_true")
    ap.add_argument("--mib", type=int, default=624, help="bytes per copy (624 MiB = one 32-image e2e chunk)")
    args = ap.parse_args()
    nbytes = args.mib * 1024 * 1024
    if not args.all_gpus:
        dev = torch.device("cuda:0")
        out = {}
        for mb in (64, args.mib, 2048):
            h = torch.empty(mb * 1024 * 1024, dtype=torch.uint8).pin_memory()
            d = torch.empty_like(h, device=dev)
            out[f"h2d_pinned_{mb}MiB_GBps"] = _rate(d, h)
        out[f"h2d_write_combined_{args.mib}MiB_GBps"] = _write_combined_rate(nbytes, dev)
        h = torch.empty(nbytes, dtype=torch.uint8)
        out[f"h2d_pageable_{args.mib}MiB_GBps"] = _rate(torch.empty_like(h, device=dev), h, reps=3)
        out["host_cores"] = os.cpu_count()
        print(json.dumps(out))
        return
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpus = sorted(os.sched_getaffinity(0))
    per = max(1, len(cpus) // world)
    os.sched_setaffinity(0, set(cpus[local * per:(local + 1) * per]) or set(cpus))
    dist.init_process_group("nccl", device_id=dev)
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()           # first touched after binding
    h.fill_(rank)
    d = torch.empty_like(h, device=dev)
    solo = None
    for r in range(world):                                            # each rank alone, in turn
        dist.barrier()
        if r == rank:
            mine_alone = _rate(d, h)
        dist.barrier()
    dist.barrier()
    torch.cuda.synchronize()
    together = _rate(d, h, reps=16)                                   # everyone at once
    dist.barrier()
    rates = [None] * world
    dist.all_gather_object(rates, {"rank": rank, "alone_GBps": mine_alone, "together_GBps": together,
                                   "cpus": sorted(os.sched_getaffinity(0))[:2] + ["..."]})
    if rank == 0:
        print(json.dumps({"world": world, "mib_per_copy": args.mib, "host_cores": os.cpu_count(),
                          "per_gpu": rates,
                          "aggregate_together_GBps": sum(r["together_GBps"] for r in rates),
                          "sum_of_alone_GBps": sum(r["alone_GBps"] for r in rates)}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
