"""-m gpu: size-independent properties at BASELINE.json's full sizes (B = 32 pairs @256^2), where an element-wise
comparison against the CPU oracle would take minutes: exact homogeneity and batch-permutation equivariance of the
tensor-core convs, additivity of the weight gradient over the batch, run-to-run bit determinism of the whole step."""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


# generator layers at B=32: (H=W, Cin, Cout) — L5.conv1 (halo kernel, N=64), L4.conv1 (halo, N=128), L3.conv1 (per-tap)
FULL = [(256, 128, 64), (128, 256, 128), (64, 512, 256)]


@pytest.mark.parametrize("res,cin,cout", FULL)
def test_conv_homogeneity_and_batch_equivariance_full_size(cuda_device, res, cin, cout):
    from speak_hack_b200 import ops

    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(32, res, res, cin, generator=g, device=dev).to(BF)
    wk = (torch.randn(cout, 9 * cin, generator=g, device=dev) * 0.03).to(BF)
    y = ops.conv_gemm(x, wk, 3, ops.EPI_PLAIN)
    # scaling the weights by 2 is exact in bf16 and in the fp32 accumulator: the output doubles bit for bit
    y2 = ops.conv_gemm(x, (wk.float() * 2).to(BF), 3, ops.EPI_PLAIN)
    assert torch.equal(y2.float(), y.float() * 2)
    # images are independent: permuting the batch permutes the output, bit for bit (same per-pixel accumulation order
    # whatever CTA / tile slot an image lands in)
    perm = torch.randperm(32, generator=torch.Generator().manual_seed(1)).to(dev)
    yp = ops.conv_gemm(x[perm].contiguous(), wk, 3, ops.EPI_PLAIN)
    assert torch.equal(yp, y[perm])
    # zero padding: an all-zero image produces an all-zero output and does not leak into its neighbours
    x0 = x.clone()
    x0[5].zero_()
    y0 = ops.conv_gemm(x0, wk, 3, ops.EPI_PLAIN)
    assert float(y0[5].abs().max()) == 0.0 and torch.equal(y0[4], y[4]) and torch.equal(y0[6], y[6])


@pytest.mark.parametrize("res,cin,cout", FULL)
def test_wgrad_additive_over_batch_full_size(cuda_device, res, cin, cout):
    from speak_hack_b200 import ops

    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(2)
    x = torch.randn(32, res, res, cin, generator=g, device=dev).to(BF)
    dy = torch.randn(32, res, res, cout, generator=g, device=dev).to(BF)
    full = ops.conv_wgrad(x, dy, 3)
    halves = ops.conv_wgrad(x[:16].contiguous(), dy[:16].contiguous(), 3)
    halves = ops.conv_wgrad(x[16:].contiguous(), dy[16:].contiguous(), 3, dw=halves, beta=1.0)
    assert rel_l2(halves, full) < 1e-4  # exact bf16 products; 2M-term fp32 sums in a different order
    again = ops.conv_wgrad(x, dy, 3)
    assert torch.equal(again, full)  # fixed split-K order: bit-deterministic


def test_train_step_bit_deterministic_full_batch(cuda_device):
    """Two independently constructed trainers with the same seeds produce bit-identical losses and Gd parameters for
    two B=32 CUDA-graph steps (every reduction in the step has a fixed order, with five streams in the graph)."""
    import speak_hack_b200 as P
    from speak_hack_b200.trainer import IRFDTrainer

    dev = cuda_device
    g = torch.Generator().manual_seed(7)
    xs = (torch.rand(32, 3, 256, 256, generator=g) * 2 - 1).to(dev)
    xt = (torch.rand(32, 3, 256, 256, generator=g) * 2 - 1).to(dev)
    results = []
    for _ in range(2):
        torch.manual_seed(0)
        torch.cuda.manual_seed(0)
        net = P.IRFD().to(dev).train()
        tr = IRFDTrainer(net, use_cuda_graph=True)
        torch.manual_seed(11)
        torch.cuda.manual_seed(11)
        losses = [float(tr.train_step(xs, xt)) for _ in range(2)]
        torch.cuda.synchronize()
        assert all(l == l and abs(l) != float("inf") for l in losses), losses
        results.append((losses, tr.flat.clone()))
        del tr, net
        torch.cuda.empty_cache()
    assert results[0][0] == results[1][0], (results[0][0], results[1][0])
    assert torch.equal(results[0][1], results[1][1])
