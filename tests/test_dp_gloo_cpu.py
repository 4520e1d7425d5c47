"""CPU, world_size 2 over gloo: the data-parallel gradient exchange (dp.py) — bucket averaging, scatter-back into the
original gradient tensors, launch order during a simulated IRFD backward, and shard-average == full-batch gradient."""
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

        dev = torch.device("cpu")
        # --- 1. flat bucket + list bucket are averaged; list bucket is scattered back into the same tensors
        b = GradBuckets(dev)
        flat = torch.full((10,), float(rank + 1))
        g1, g2 = torch.full((3, 2), float(rank)), torch.arange(4.0) * (rank + 1)
        b.launch(flat=flat)
        b.launch([g1, g2])
        b.finish()
        ok1 = torch.allclose(flat, torch.full((10,), 1.5)) and torch.allclose(g1, torch.full((3, 2), 0.5)) and \
            torch.allclose(g2, torch.arange(4.0) * 1.5)

        # --- 2. schedule: simulated backward order Ep_t, Ee_t, Ei_t, Ep_s, Ee_s, Ei_s
        encs = [torch.nn.Linear(4, 4) for _ in range(3)]  # stand-ins for Ei, Ee, Ep
        for e in encs:
            for p in e.parameters():
                p.grad = torch.full_like(p, float(rank + 1))
        gd = torch.full((7,), float(2 * rank))
        sched = BucketSchedule(GradBuckets(dev), gd, encs)
        for e in (encs[2], encs[1], encs[0], encs[2], encs[1], encs[0]):
            sched.pre()
            sched.post(e)
        sched.final()
        ok2 = sched.order == ["Gd", "E2", "E1", "E0"] and torch.allclose(gd, torch.full((7,), 1.0)) and all(
            torch.allclose(p.grad, torch.full_like(p, 1.5)) for e in encs for p in e.parameters())

        # --- 3. property: mean of per-shard gradients == gradient of the mean loss over the global batch
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4))
        x, y = torch.randn(8, 8), torch.randn(8, 4)
        full = torch.autograd.grad(torch.nn.functional.mse_loss(model(x), y), list(model.parameters()))
        shard = slice(rank * 4, rank * 4 + 4)
        loss = torch.nn.functional.mse_loss(model(x[shard]), y[shard])
        loss.backward()
        b3 = GradBuckets(dev)
        b3.launch([p.grad for p in model.parameters()])
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
    b.launch([t])
    b.finish()
    assert b.world == 1 and torch.equal(t, torch.ones(3))
