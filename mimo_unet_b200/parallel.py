"""Data-parallel plumbing (one process per GPU, torch.distributed): bucketed all-reduce of the flat gradient
buffer and the [S] per-subnetwork loss exchange that keeps every rank's loss buffer identical.

The reference defines no multi-GPU behaviour (devices=1 everywhere); semantics follow standard DDP: per-rank
BatchNorm statistics, per-rank shuffles, gradients averaged over ranks."""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class GradientSynchronizer:
    """Averages one flat fp32 gradient tensor over all ranks in `num_buckets` contiguous buckets.

    Bucket boundaries are aligned to 128 elements. With NCCL the reductions are issued asynchronously on the
    communicator's stream (so the first buckets travel over NVLink while later ones are still being enqueued) and
    joined in wait(); with gloo (CPU tests) AVG is emulated as SUM / world."""

    def __init__(self, num_buckets: int = 4, group=None):
        self.num_buckets = max(1, int(num_buckets))
        self.group = group
        self._work: List = []
        self._views: List[torch.Tensor] = []

    @staticmethod
    def bucket_ranges(numel: int, num_buckets: int):
        per = (numel + num_buckets - 1) // num_buckets
        per = (per + 127) // 128 * 128
        out, start = [], 0
        while start < numel:
            end = min(numel, start + per)
            out.append((start, end))
            start = end
        return out

    def start(self, flat: torch.Tensor):
        if world_size() == 1:
            return
        avg_ok = flat.is_cuda
        self._views = [flat[a:b] for a, b in self.bucket_ranges(flat.numel(), self.num_buckets)]
        # issue in REVERSE order: the tail of the flat buffer holds decoder gradients, produced first by backward
        for v in reversed(self._views):
            op = dist.ReduceOp.AVG if avg_ok else dist.ReduceOp.SUM
            self._work.append((dist.all_reduce(v, op=op, group=self.group, async_op=True), v, avg_ok))

    def wait(self):
        w = world_size()
        for work, v, avg_ok in self._work:
            work.wait()
            if not avg_ok:
                v.div_(w)
        self._work = []


class OverlappedGradientSynchronizer:
    """Gradient all-reduce overlapped with backward (NCCL over NVLink, one process per GPU).

    The C executor records one CUDA event per backward stage as soon as that stage's parameter gradients are final
    (decoders + heads, core up path, core down path, encoders). Stages are contiguous slices of the flat gradient
    buffer, so each stage is one bucket: right after the (asynchronous) enqueue of backward returns, the buckets are
    all-reduced (mean) on a side stream that waits only for the bucket's event, i.e. the first buckets travel while the
    GPU is still computing the rest of backward. wait() joins them before the optimizer step."""

    NUM_STAGES = 4

    def __init__(self, group=None):
        self.group = group
        self.events = [torch.cuda.Event() for _ in range(self.NUM_STAGES)]
        for e in self.events:
            e.record()  # torch creates the cudaEvent_t lazily on the first record
        # MIMO_DP_PRIORITY=1: high-priority side stream (the collective's CTAs are placed before the next compute kernel's)
        self.side = torch.cuda.Stream(priority=-1) if os.environ.get("MIMO_DP_PRIORITY", "0") == "1" else torch.cuda.Stream()
        # MIMO_DP_BUCKETS = 4 (one bucket per backward stage, default) | 2 (decoders + core up, core down + encoders) | 1 (one
        # all-reduce after backward: no overlap, but no collective kernel competing with the persistent compute kernels for SMs)
        self.buckets = int(os.environ.get("MIMO_DP_BUCKETS", "4"))
        # measurement only (results are wrong): MIMO_DP_DEBUG_SKIP = grad | loss | both drops the named exchange to attribute its cost
        self._skip = os.environ.get("MIMO_DP_DEBUG_SKIP", "")
        self._work: List = []
        self._after_wait = None   # callable run by wait() once the buckets have been joined (gradient-aliasing check)

    def launch(self, flat: torch.Tensor, bounds):
        """bounds[k] = (start, end) element range of stage k inside `flat`."""
        if world_size() == 1 or self._skip in ("grad", "both"):
            return
        groups = {4: [[0], [1], [2], [3]], 2: [[0, 1], [2, 3]], 1: [[0, 1, 2, 3]]}.get(self.buckets, [[0], [1], [2], [3]])
        for g in groups:
            live = [k for k in g if bounds[k][1] > bounds[k][0]]
            if not live:
                continue
            a, b = min(bounds[k][0] for k in live), max(bounds[k][1] for k in live)   # stages are adjacent slices of the flat buffer
            with torch.cuda.stream(self.side):
                self.side.wait_event(self.events[g[-1]])   # the last stage of the group is final last
                self._work.append(dist.all_reduce(flat[a:b], op=dist.ReduceOp.AVG, group=self.group, async_op=True))

    def wait(self):
        for w in self._work:
            w.wait()
        self._work = []
        if self._after_wait is not None:
            cb, self._after_wait = self._after_wait, None
            cb()

    def exchange_loss(self, loss: torch.Tensor, loss_buffer) -> None:
        """Data-parallel loss-buffer update OFF the compute stream: the [S] per-subnetwork loss is averaged over ranks and
        added to the device loss buffer on the side stream; the next step's loss kernel waits for `loss_event` before it reads
        the buffer. (The result is not needed before the next step, so nothing on the compute stream blocks on NCCL.)"""
        if self._skip in ("loss", "both"):
            loss_buffer.add(loss.detach())
            return
        cur = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(cur)
        loss.record_stream(self.side)
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            l = loss.detach().clone()
            allreduce_mean_(l, self.group)
            loss_buffer.add(l)
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.loss_event = ev

    loss_event = None

    def join_loss_exchange(self) -> None:
        if self.loss_event is not None:
            torch.cuda.current_stream().wait_event(self.loss_event)
            self.loss_event = None


def allreduce_mean_(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place mean over ranks (used for the [S] loss vector that feeds the loss buffer)."""
    w = world_size()
    if w == 1:
        return t
    if t.is_cuda:
        dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.div_(w)
    return t
