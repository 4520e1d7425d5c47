"""Data-parallel gradient exchange for the IRFD trainer: bucketed all-reduce(avg) overlapped with backward.

The reference delegates multi-GPU training to HF accelerate -> torch DDP (train.py:333-338, 399-401): batch-sharded
replicas, per-rank BatchNorm statistics (no SyncBN), gradients averaged by all-reduce, BatchNorm buffers broadcast from
rank 0 (DDP's `broadcast_buffers=True` default).  Here the exchange is explicit: one process per GPU,
`torch.distributed` (NCCL over NVLink 5 / NVSwitch on the GPU box; gloo on CPU for the tests).  Every bucket is a FLAT
fp32 gradient buffer the backward kernels write into directly (trainer.py), all-reduced in place on a side stream as
soon as the backward pass has finished producing it — no flatten / unflatten copies.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class GradBuckets:
    """Launch/finish pairs of asynchronous averaged all-reduces over flat buffers (in place).

    On CUDA the collective runs on a dedicated stream ordered after the work already enqueued on the current stream;
    on CPU (gloo) it runs synchronously.  `enabled = False` turns launches into no-ops (bench.py's dp_check captures a
    second, communication-free graph to obtain the per-rank gradients)."""

    def __init__(self, device: torch.device):
        self.world = world_size()
        self.device = device
        self.cuda = device.type == "cuda"
        self.comm_stream = torch.cuda.Stream(device) if (self.cuda and self.world > 1) else None
        self.enabled = True
        # callable returning the streams (besides the current one) that may still be producing the gradients of the
        # bucket being launched: the weight-gradient side stream while it has un-joined work (ops.SideStream)
        self.producers = lambda: []
        self._pending = 0
        self.launched_bytes = 0

    def launch(self, flat: torch.Tensor) -> None:
        if self.world == 1 or not self.enabled or flat is None or flat.numel() == 0:
            return
        self.launched_bytes += flat.numel() * flat.element_size()
        if self.cuda:
            ready = torch.cuda.Event()
            ready.record()
            for ps in self.producers():
                self.comm_stream.wait_stream(ps)
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ready)
                flat.record_stream(self.comm_stream)
                dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat.mul_(1.0 / self.world)
        self._pending += 1

    def finish(self) -> None:
        """Make the averaged gradients visible to the current stream."""
        if self.world == 1 or not self._pending:
            return
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self._pending = 0


class BucketSchedule:
    """Decides WHEN each bucket is complete during IRFD's backward.

    Backward runs the generator first, then the lockstep encoder pass from ResNet stage 4 down to the stem
    (encoder_group._EncoderGroupFn.backward reports "pre", one "stage" event per finished stage, "post"):
      * the generator bucket (all of Gd) is launched when the encoder backward starts;
      * the bucket of stage s (that stage's parameters of all three encoders) when stage s has been written, so only
        the small stem bucket is left exposed after the last backward kernel.
    `final()` launches whatever has not been launched (per-encoder fallback path) and waits for the collectives.
    """

    def __init__(self, buckets: GradBuckets, gd_flat_grad: torch.Tensor, stage_flats: Dict[int, torch.Tensor]):
        self.buckets = buckets
        self.gd_flat_grad = gd_flat_grad
        self.stage_flats = stage_flats  # {stage index (7 = layer4 ... 4 = layer1, 3 = stem): flat gradient buffer}
        self.reset()

    def reset(self):
        self.gd_launched = False
        self.launched = {k: False for k in self.stage_flats}
        self.order: List[str] = []

    def on_event(self, kind: str, stage: Optional[int]) -> None:
        if kind == "pre":
            self._launch_gd()
        elif kind == "stage":
            self._launch_gd()
            self._launch_stage(stage)

    def final(self):
        self._launch_gd()
        for k in sorted(self.stage_flats, reverse=True):
            self._launch_stage(k)
        self.buckets.finish()

    def _launch_gd(self):
        if not self.gd_launched:
            self.gd_launched = True
            self.buckets.launch(self.gd_flat_grad)
            self.order.append("Gd")

    def _launch_stage(self, k):
        if k in self.launched and not self.launched[k]:
            self.launched[k] = True
            self.buckets.launch(self.stage_flats[k])
            self.order.append(f"S{k}")


def broadcast_buffers(modules, src: int = 0) -> None:
    """DDP's `broadcast_buffers=True` (the reference's default, train.py:399): every rank takes rank `src`'s BatchNorm
    running buffers.  DDP does this before every forward; the buffers do not enter train-mode arithmetic, so doing it
    once before an evaluation or a checkpoint leaves every rank with exactly the state DDP would have (rank `src`'s
    own trajectory)."""
    if world_size() == 1:
        return
    bufs = [b for m in modules for b in m.buffers()]
    by_dtype: Dict[torch.dtype, List[torch.Tensor]] = {}
    for b in bufs:
        by_dtype.setdefault(b.dtype, []).append(b)
    for group in by_dtype.values():
        flat = torch.cat([b.reshape(-1) for b in group])
        dist.broadcast(flat, src=src)
        off = 0
        for b in group:
            b.copy_(flat[off: off + b.numel()].view_as(b))
            off += b.numel()
