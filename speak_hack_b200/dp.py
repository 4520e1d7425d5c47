"""Data-parallel gradient exchange for the IRFD trainer: bucketed all-reduce(avg) overlapped with backward.

The reference delegates multi-GPU training to HF accelerate -> torch DDP (train.py:333-338, 399-401): batch-sharded
replicas, per-rank BatchNorm statistics (no SyncBN), gradients averaged by all-reduce.  Here the exchange is explicit:
one process per GPU, `torch.distributed` (NCCL over NVLink 5 / NVSwitch on the GPU box; gloo on CPU for the tests),
each bucket all-reduced on a side stream as soon as the backward pass has finished producing it.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class GradBuckets:
    """Launch/finish pairs of asynchronous averaged all-reduces.

    launch(tensors)      : flatten `tensors` into one buffer and all-reduce it (result copied back on finish()).
    launch(flat=buffer)  : all-reduce an already-flat buffer in place.
    On CUDA the collective runs on a dedicated stream ordered after the work already enqueued on the current stream;
    on CPU (gloo) it runs synchronously.
    """

    def __init__(self, device: torch.device):
        self.world = world_size()
        self.device = device
        self.cuda = device.type == "cuda"
        self.comm_stream = torch.cuda.Stream(device) if (self.cuda and self.world > 1) else None
        # streams (besides the current one) that may have produced the gradients of a bucket: the encoders' backward
        # passes run on their own streams (model.IRFD.encoder_streams)
        self.producer_streams = []
        self._pending = []
        self.launched_bytes = 0

    def launch(self, tensors: Optional[List[torch.Tensor]] = None, flat: Optional[torch.Tensor] = None) -> None:
        if self.world == 1:
            return
        unflatten = None
        if flat is None:
            if not tensors:
                return
            flat = torch.cat([t.reshape(-1) for t in tensors])
            unflatten = list(tensors)
        self.launched_bytes += flat.numel() * flat.element_size()
        if self.cuda:
            ready = torch.cuda.Event()
            ready.record()
            for ps in self.producer_streams:
                self.comm_stream.wait_stream(ps)
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ready)
                flat.record_stream(self.comm_stream)
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
                flat.mul_(1.0 / self.world)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat.mul_(1.0 / self.world)
        self._pending.append((flat, unflatten))

    def finish(self) -> None:
        """Make the averaged gradients visible to the current stream and scatter them back into their tensors."""
        if self.world == 1:
            return
        if self.cuda:
            cur = torch.cuda.current_stream()
            cur.wait_stream(self.comm_stream)
            for ps in self.producer_streams:  # gradients that were not bucketed are still produced there
                cur.wait_stream(ps)
        for flat, unflatten in self._pending:
            if unflatten is not None:
                off, views = 0, []
                for t in unflatten:
                    views.append(flat[off: off + t.numel()].view_as(t))
                    off += t.numel()
                torch._foreach_copy_(unflatten, views)
        self._pending = []


class BucketSchedule:
    """Decides WHEN each bucket is complete during IRFD's backward.

    Backward runs the two generator calls first, then the six encoder passes in reverse forward order
    (Ep(x_t), Ee(x_t), Ei(x_t), Ep(x_s), Ee(x_s), Ei(x_s)).  Encoder passes call `pre()` when they start — every
    gradient produced by earlier autograd nodes has been accumulated by then — and `post(enc)` when they end.
      * the generator bucket is launched at the first `pre()`;
      * encoder E's bucket is launched at the first `pre()`/`final()` after E's second pass.
    """

    def __init__(self, buckets: GradBuckets, gd_flat_grad: torch.Tensor, encoders: List[torch.nn.Module]):
        self.buckets = buckets
        self.gd_flat_grad = gd_flat_grad
        self.encoders = encoders
        self.reset()

    def reset(self):
        self.gd_launched = False
        self.passes = {id(e): 0 for e in self.encoders}
        self.launched = {id(e): False for e in self.encoders}
        self.order: List[str] = []

    def pre(self):
        self._launch_ready(final=False)

    def post(self, enc, passes: int = 1):
        """`passes` = how many of the reference's per-image encoder calls this backward covered (2 for a paired pass)."""
        self.passes[id(enc)] += passes

    def final(self):
        self._launch_ready(final=True)
        self.buckets.finish()

    def _launch_ready(self, final: bool):
        if not self.gd_launched:
            self.gd_launched = True
            self.buckets.launch(flat=self.gd_flat_grad)
            self.order.append("Gd")
        for i, e in enumerate(self.encoders):
            if (self.passes[id(e)] >= 2 or final) and not self.launched[id(e)]:
                self.launched[id(e)] = True
                gs = [p.grad for p in e.parameters() if p.grad is not None]
                self.buckets.launch(gs)
                self.order.append(f"E{i}")
