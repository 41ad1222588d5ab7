"""Vocabulary-parallel head: the prompts sharded over the GPUs of one NVLink box.

The default multi-GPU mode shards the image batch (``shard.py``; no exchange step).  This module
is the other axis, for a vocabulary that is large against the batch (a latency-bound batch of one
image against 10^4..10^5 prompts): every rank sees the whole batch, owns the contiguous class
range ``class_range(rank, world, C)`` and the per-anchor ``(max, argmax)`` of
model/yolo_clip.py:198-206 is reduced over the class shards.  That reduction is the only exchange
step of the path.  Two implementations of it:

``exchange="fused"`` (the product): the similarity kernel packs each finished row into a 64-bit
key and max-reduces it with system-scope atomics into every rank's key array through NVLink peer
mappings, from the epilogue warp that produced it; a flag handshake (two tiny kernels) replaces
the collective's synchronisation.  No library collective on the data path.

``exchange="allreduce"`` (the baseline it is measured against, and what the gloo CPU tests run):
local ``(max, argmax)`` -> int64 keys -> ``dist.all_reduce(MAX)`` -> unpack.

Both give, on every rank, the scores and global class ids of the full vocabulary; K3 and K4 then
run unchanged.  Key order: score ascending, class index DESCENDING, so the maximum is the best
score and, among equal scores, the lowest class index (torch.max's CPU tie rule).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi, ops
from ._cabi import check, lib
from .pipeline import HeadConfig, HeadPipeline

MAX_PEERS = 8


def class_range(rank: int, world: int, num_classes: int) -> Tuple[int, int]:
    """Contiguous, balanced class shard of ``rank``: sizes differ by at most one, every rank
    non-empty when ``num_classes >= world``."""
    if not (0 <= rank < world) or num_classes < world:
        raise ValueError("ovdet: need 0 <= rank < world <= num_classes")
    base, extra = divmod(num_classes, world)
    c0 = rank * base + min(rank, extra)
    return c0, c0 + base + (1 if rank < extra else 0)


# ---- host-side restatement of the key (used by the CPU tests and by the gloo path) ----------
def pack_keys_host(scores: np.ndarray, class_ids: np.ndarray, class_offset: int = 0) -> np.ndarray:
    """numpy twin of ``vp_pack_key`` (csrc/common.cuh) in its signed-int64 form."""
    s = np.ascontiguousarray(scores, dtype=np.float32) + np.float32(0.0)          # -0.0 -> +0.0
    u = s.view(np.uint32)
    ordered = np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000)).astype(np.uint64)
    low = (np.uint64(0xFFFFFFFF) - (np.asarray(class_ids).astype(np.uint64) + np.uint64(class_offset)))
    key = (ordered << np.uint64(32)) | low
    return (key ^ np.uint64(1 << 63)).view(np.int64)


def unpack_keys_host(keys: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    key = np.ascontiguousarray(keys, dtype=np.int64).view(np.uint64) ^ np.uint64(1 << 63)
    ordered = (key >> np.uint64(32)).astype(np.uint32)
    u = np.where(ordered & np.uint32(0x80000000), ordered & np.uint32(0x7FFFFFFF), ~ordered)
    cls = (np.uint64(0xFFFFFFFF) - (key & np.uint64(0xFFFFFFFF))).astype(np.int32)
    return u.astype(np.uint32).view(np.float32), cls


# ---- device wrappers ----------------------------------------------------------------------
def pack_score_keys(scores: torch.Tensor, class_ids: torch.Tensor, class_offset: int,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    ops._require_cuda(scores, "scores", torch.float32)
    ops._require_cuda(class_ids, "class_ids", torch.int32)
    assert scores.is_contiguous() and class_ids.is_contiguous() and scores.shape == class_ids.shape
    if out is None:
        out = torch.empty(scores.shape, device=scores.device, dtype=torch.int64)
    with torch.cuda.device(scores.device):
        check(lib().ovdet_pack_score_keys(scores.data_ptr(), class_ids.data_ptr(), scores.numel(),
                                          int(class_offset), out.data_ptr(), ops._stream(scores)),
              "ovdet_pack_score_keys")
    return out


def unpack_score_keys(keys: torch.Tensor, scores: torch.Tensor, class_ids: torch.Tensor) -> None:
    ops._require_cuda(keys, "keys", torch.int64)
    assert keys.is_contiguous() and scores.is_contiguous() and class_ids.is_contiguous()
    with torch.cuda.device(keys.device):
        check(lib().ovdet_unpack_score_keys(keys.data_ptr(), keys.numel(), scores.data_ptr(),
                                            class_ids.data_ptr(), ops._stream(keys)),
              "ovdet_unpack_score_keys")


class PeerBuffer:
    """One rank's exchange buffer (device memory owned by the library, exportable over CUDA IPC)."""

    def __init__(self, rows: int, world: int, device):
        self.rows, self.world, self.device = rows, world, torch.device(device)
        self.bytes = lib().ovdet_vp_buffer_bytes(rows, world)
        if self.bytes == 0:
            raise ValueError("ovdet: bad exchange buffer shape")
        ptr = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        with torch.cuda.device(self.device):
            check(lib().ovdet_peer_buffer_create(self.bytes, ctypes.byref(ptr), handle),
                  "ovdet_peer_buffer_create")
            self.ptr = ptr.value
            self.handle = handle.raw
            check(lib().ovdet_vp_buffer_init(self.ptr, rows, world,
                                             torch.cuda.current_stream(self.device).cuda_stream),
                  "ovdet_vp_buffer_init")
            torch.cuda.synchronize(self.device)
        self._opened = []

    def open_peer(self, handle: bytes) -> int:
        ptr = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().ovdet_peer_buffer_open(ctypes.create_string_buffer(handle, 64), ctypes.byref(ptr)),
                  "ovdet_peer_buffer_open")
        self._opened.append(ptr.value)
        return ptr.value

    def close(self) -> None:
        with torch.cuda.device(self.device):
            for p in self._opened:
                lib().ovdet_peer_buffer_close(p)
            self._opened = []
            if self.ptr:
                lib().ovdet_peer_buffer_destroy(self.ptr)
                self.ptr = None


class VocabParallelHead:
    """``HeadPipeline`` with the vocabulary sharded over ``world`` ranks.

    ``group``: a ``torch.distributed`` process group (one process per GPU); ``None`` with explicit
    ``rank`` / ``world`` / ``peers`` builds a *virtual* rank inside one process (the single-GPU
    parity tests: several ranks on one device, driven phase by phase).
    """

    def __init__(self, batch: int, level_shapes: Sequence[Tuple[int, int]], num_classes: int,
                 config: HeadConfig = HeadConfig(), device="cuda", group=None, rank: Optional[int] = None,
                 world: Optional[int] = None, exchange: str = "fused"):
        if config.precision != "bf16" or config.logits_dtype is not None:
            raise ValueError("ovdet: the vocabulary-parallel head is the bf16 fused-max path")
        if exchange not in ("fused", "allreduce"):
            raise ValueError("ovdet: exchange is 'fused' or 'allreduce'")
        self.group = group
        if group is not None or (rank is None and world is None):
            import torch.distributed as dist
            self._dist = dist
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self._dist = None
        if world > MAX_PEERS:
            raise ValueError(f"ovdet: at most {MAX_PEERS} ranks")
        self.rank, self.world, self.exchange = rank, world, exchange
        self.num_classes = num_classes
        self.c0, self.c1 = class_range(rank, world, num_classes)
        self.cfg = config
        # the local pipeline owns every buffer; its similarity stage is replaced below
        self.pipe = HeadPipeline(batch, level_shapes, self.c1 - self.c0, config, device=device)
        self.device = self.pipe.device
        self.rows = batch * self.pipe.anchors
        # sticky status word of the bounded waits, in pinned host memory that the kernels write
        # through the unified address space: the host reads it without synchronising
        self.status = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.timeout_ms = 2000
        self.first_timeout_ms = 30000       # step 1: a peer may still be loading modules / encoding tensor maps
        self._steps = 0
        self._sim_ws = None
        self.buffer = None
        self.peer_ptrs = None
        self._keys = None
        if exchange == "fused":
            self.buffer = PeerBuffer(self.rows, world, self.device)
            if self._dist is not None:
                handles = [None] * world
                self._dist.all_gather_object(handles, self.buffer.handle, group=group)
                ptrs = [self.buffer.ptr if g == rank else self.buffer.open_peer(handles[g])
                        for g in range(world)]
                self.connect(ptrs)
                self._dist.barrier(group=group)        # every buffer is zeroed before step 1
        else:
            self._keys = torch.empty(batch, self.pipe.anchors, device=self.device, dtype=torch.int64)

    # virtual ranks (one process): hand every rank the list of all buffers
    def connect(self, ptrs: Sequence[int]) -> None:
        assert len(ptrs) == self.world
        self.peer_ptrs = (ctypes.c_void_p * self.world)(*ptrs)

    def set_vocabulary(self, text: torch.Tensor) -> None:
        """``text``: the FULL ``[C, D]`` vocabulary (replicated input); this rank keeps its rows."""
        assert text.shape[0] == self.num_classes
        self.pipe.set_vocabulary(text[self.c0:self.c1])

    def set_geometry(self, *a, **k) -> None:
        self.pipe.set_geometry(*a, **k)

    @property
    def result(self):
        return self.pipe.result

    @property
    def scores(self):
        return self.pipe.scores

    @property
    def class_ids(self):
        return self.pipe.class_ids

    # ---- the three phases of the fused exchange (separate so that virtual ranks can interleave) ----
    def similarity(self, obj_embeds: Sequence[torch.Tensor]) -> None:
        pipe, cfg = self.pipe, self.cfg
        if not ops.fused_supported(obj_embeds) or cfg.embed_dim != 512:
            raise ValueError("ovdet: inputs not addressable by the fused similarity kernel")
        first = obj_embeds[0]
        n = len(obj_embeds)
        ptrs = (ctypes.c_void_p * n)(*[e.data_ptr() for e in obj_embeds])
        hw = (ctypes.c_int64 * n)(*[e.shape[2] * e.shape[3] for e in obj_embeds])
        sb = (ctypes.c_int64 * n)(*[e.stride(0) for e in obj_embeds])
        sd = (ctypes.c_int64 * n)(*[e.stride(1) for e in obj_embeds])
        if self._sim_ws is None:
            nbytes = lib().ovdet_similarity_split_workspace_bytes(pipe.batch, pipe.anchors)
            self._sim_ws = torch.zeros(max(16, nbytes), device=self.device, dtype=torch.uint8)
        in16 = first.dtype == torch.bfloat16
        with torch.cuda.device(self.device):
            stream = ops._stream(first)
            if self.exchange == "fused":
                check(lib().ovdet_similarity_fused_vp(
                    ptrs, hw, sb, sd, n, pipe.batch, cfg.embed_dim, pipe.text_op.data_ptr(),
                    self.c1 - self.c0, 0, float(cfg.cls_alpha), float(cfg.cls_beta), pipe.inv_norm.data_ptr(),
                    self._sim_ws.data_ptr(), self._sim_ws.numel(),
                    _cabi.OVDET_BF16 if in16 else _cabi.OVDET_F32, self.c0, self.peer_ptrs, self.world,
                    self.rank, stream), "ovdet_similarity_fused_vp")
            else:
                ops.similarity_fused(obj_embeds, pipe.text_op, cfg.cls_alpha, cfg.cls_beta, logits_dtype=None,
                                     want_max=True, row_max=pipe.scores, row_arg=pipe.class_ids,
                                     inv_norm=pipe.inv_norm)

    def signal(self) -> None:
        with torch.cuda.device(self.device):
            check(lib().ovdet_vp_signal(self.peer_ptrs, self.world, self.rank, self.rows,
                                        torch.cuda.current_stream(self.device).cuda_stream), "ovdet_vp_signal")

    def merge(self) -> None:
        """After this, ``scores`` / ``class_ids`` hold the full-vocabulary result on this rank."""
        pipe = self.pipe
        with torch.cuda.device(self.device):
            if self.exchange == "fused":
                check(lib().ovdet_vp_wait_unpack(self.buffer.ptr, self.world, self.rows,
                                                 pipe.scores.data_ptr(), pipe.class_ids.data_ptr(),
                                                 self.status.data_ptr(), self._timeout(),
                                                 torch.cuda.current_stream(self.device).cuda_stream),
                      "ovdet_vp_wait_unpack")
            else:
                pack_score_keys(pipe.scores, pipe.class_ids, self.c0, out=self._keys)
                self._dist.all_reduce(self._keys, op=self._dist.ReduceOp.MAX, group=self.group)
                unpack_score_keys(self._keys, pipe.scores, pipe.class_ids)

    def _timeout(self) -> int:
        t = self.first_timeout_ms if self._steps == 0 else self.timeout_ms
        self._steps += 1
        return int(t)

    def raise_if_timed_out(self, synchronize: bool = False) -> None:
        """A wait on the peers' flags expired in an earlier step (or, with ``synchronize``, in any
        step enqueued so far): that step and every later one carry NO detections on this rank (the
        status is sticky and the key arrays are no longer handed back).  Raises; recover with
        ``reset()`` on every rank."""
        if synchronize:
            torch.cuda.synchronize(self.device)
        if int(self.status[0]) != 0:
            raise RuntimeError(f"ovdet: rank {self.rank}: a peer did not signal within the exchange timeout; "
                               "the steps since then returned no detections (VocabParallelHead.reset() on "
                               "every rank re-arms the exchange)")

    def reset(self) -> None:
        """Re-arm after a timeout: counters, flags and key arrays back to their initial state.  Every
        rank must call it (collective when a process group is attached)."""
        torch.cuda.synchronize(self.device)
        if self._dist is not None:
            self._dist.barrier(group=self.group)
        if self.buffer is not None:
            with torch.cuda.device(self.device):
                check(lib().ovdet_vp_buffer_init(self.buffer.ptr, self.rows, self.world,
                                                 torch.cuda.current_stream(self.device).cuda_stream),
                      "ovdet_vp_buffer_init")
            torch.cuda.synchronize(self.device)
        self.status.zero_()
        self._steps = 0
        if self._dist is not None:
            self._dist.barrier(group=self.group)

    def run(self, obj_embeds: Sequence[torch.Tensor], box_preds: Sequence[torch.Tensor]):
        """One pass: sharded similarity + exchange, then K3 / K4 on the merged scores.  Every rank
        returns the same detections.  With the fused exchange the whole step is one C call
        (``ovdet_head_step_vp``) whose arguments never change, so it can also be captured in a CUDA
        graph (``capture`` / ``replay``).  Raises when an EARLIER step's wait expired (read from the
        pinned status word, no synchronisation); ``raise_if_timed_out(synchronize=True)`` checks
        the step just enqueued."""
        self.raise_if_timed_out()
        self.pipe.check_inputs(obj_embeds, box_preds)
        if self.exchange == "fused" and self.pipe._single_call_ok(box_preds):
            if not ops.fused_supported(obj_embeds) or self.cfg.embed_dim != 512:
                raise ValueError("ovdet: inputs not addressable by the fused similarity kernel")
            a = self.pipe._fill_step_args(obj_embeds, box_preds)
            with torch.cuda.device(self.device):
                check(lib().ovdet_head_step_vp(ctypes.byref(a), self.c0, self.peer_ptrs, self.world, self.rank,
                                               self.status.data_ptr(), self._timeout(),
                                               torch.cuda.current_stream(self.device).cuda_stream),
                      "ovdet_head_step_vp")
            return self.pipe.result
        self.similarity(obj_embeds)
        if self.exchange == "fused":
            self.signal()
        self.merge()
        return self.finish(box_preds)

    def capture(self, obj_embeds: Sequence[torch.Tensor], box_preds: Sequence[torch.Tensor]) -> None:
        """Record one step into a CUDA graph (fused exchange only: its step counters live in device
        memory).  Every rank must capture, and later replay the same number of times."""
        assert self.exchange == "fused"
        self.run(obj_embeds, box_preds)                       # warm: attributes, workspaces, tensor maps
        torch.cuda.synchronize(self.device)
        if self._dist is not None:
            self._dist.barrier(group=self.group)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self.run(obj_embeds, box_preds)

    def replay(self):
        self.raise_if_timed_out()
        self._graph.replay()
        return self.pipe.result

    def finish(self, box_preds: Sequence[torch.Tensor]):
        return self.pipe._decode_and_nms(box_preds, lambda name, begin: None)

    def timed_out(self) -> bool:
        """Host check (synchronises): did a wait on the peers' flags expire?"""
        torch.cuda.synchronize(self.device)
        return bool(int(self.status[0]))

    def close(self) -> None:
        if self.buffer is not None:
            torch.cuda.synchronize(self.device)
            if self._dist is not None:
                self._dist.barrier(group=self.group)
            self.buffer.close()
            self.buffer = None
