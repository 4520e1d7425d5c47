"""CPU, world_size 2 over gloo: the data-parallel gradient exchange (dp.py) — in-place averaging of flat buckets, launch
order during a simulated IRFD backward (generator bucket, then one bucket per ResNet stage), DDP-style buffer broadcast,
and shard-average == full-batch gradient."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from speak_hack_b200.dp import BucketSchedule, GradBuckets

        from speak_hack_b200.dp import broadcast_buffers

        dev = torch.device("cpu")
        # --- 1. flat buckets are averaged in place; views into them see the result; disabled buckets stay local
        b = GradBuckets(dev)
        flat = torch.full((10,), float(rank + 1))
        view = flat[2:6].view(2, 2)            # a parameter's .grad living inside the flat bucket
        b.launch(flat)
        b.finish()
        ok1 = torch.allclose(flat, torch.full((10,), 1.5)) and torch.allclose(view, torch.full((2, 2), 1.5))
        b.enabled = False
        local = torch.full((4,), float(rank))
        b.launch(local)
        b.finish()
        ok1 = ok1 and torch.equal(local, torch.full((4,), float(rank)))

        # --- 2. schedule: the lockstep encoder backward reports pre, stage 7..4, stage 3 (stem), post
        gd = torch.full((7,), float(2 * rank))
        stages = {k: torch.full((5,), float(rank + k)) for k in (7, 6, 5, 4, 3)}
        sched = BucketSchedule(GradBuckets(dev), gd, stages)
        sched.on_event("pre", None)
        for k in (7, 6, 5, 4, 3):
            sched.on_event("stage", k)
        sched.on_event("post", None)
        sched.final()
        ok2 = sched.order == ["Gd", "S7", "S6", "S5", "S4", "S3"] and torch.allclose(gd, torch.full((7,), 1.0)) and all(
            torch.allclose(v, torch.full((5,), k + 0.5)) for k, v in stages.items())
        # fallback path (no events): final() launches everything
        sched2 = BucketSchedule(GradBuckets(dev), torch.full((3,), float(rank)), {7: torch.full((2,), float(rank))})
        sched2.final()
        ok2 = ok2 and sched2.order == ["Gd", "S7"] and torch.allclose(sched2.gd_flat_grad, torch.full((3,), 0.5))
        # DDP broadcast_buffers: every rank ends with rank 0's BatchNorm buffers (float and integer ones)
        bn = torch.nn.BatchNorm2d(3)
        bn.running_mean.fill_(float(rank + 1))
        bn.num_batches_tracked.fill_(rank + 4)
        broadcast_buffers([bn])
        ok2 = ok2 and torch.equal(bn.running_mean, torch.ones(3)) and int(bn.num_batches_tracked) == 4

        # --- 3. property: mean of per-shard gradients == gradient of the mean loss over the global batch
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4))
        x, y = torch.randn(8, 8), torch.randn(8, 4)
        full = torch.autograd.grad(torch.nn.functional.mse_loss(model(x), y), list(model.parameters()))
        shard = slice(rank * 4, rank * 4 + 4)
        loss = torch.nn.functional.mse_loss(model(x[shard]), y[shard])
        loss.backward()
        b3 = GradBuckets(dev)
        params = list(model.parameters())
        flat3 = torch.cat([p.grad.reshape(-1) for p in params])
        off = 0
        for p in params:  # re-home the gradients as views of one flat bucket, like the trainer does
            p.grad = flat3[off: off + p.numel()].view_as(p)
            off += p.numel()
        b3.launch(flat3)
        b3.finish()
        ok3 = all(torch.allclose(p.grad, f, atol=1e-6) for p, f in zip(model.parameters(), full))
        q.put((rank, ok1, ok2, ok3))
    finally:
        dist.destroy_process_group()


def test_gradient_buckets_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok1, ok2, ok3 in res:
        assert ok1, f"rank {rank}: bucket averaging failed"
        assert ok2, f"rank {rank}: bucket schedule failed"
        assert ok3, f"rank {rank}: shard-average != full-batch gradient"


def test_single_process_is_a_noop():
    from speak_hack_b200.dp import GradBuckets

    b = GradBuckets(torch.device("cpu"))
    t = torch.ones(3)
    b.launch(t)
    b.finish()
    assert b.world == 1 and torch.equal(t, torch.ones(3))
