"""-m gpu: tcgen05 implicit-GEMM conv (irfd_conv_gemm) against a torch fp32 conv on the same bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a = a.double()
    b = b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _mk(n, h, w, cin, cout, k, dev, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(n, h, w, cin, generator=g).to(dev).to(torch.bfloat16)
    wt = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(dev).to(torch.bfloat16)
    wk = wt.permute(0, 2, 3, 1).contiguous().reshape(cout, k * k * cin)  # [Cout][tap][Cin]
    return x, wt, wk


def _ref(x, wt, k):
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=k // 2)
    return y.permute(0, 2, 3, 1).contiguous()


CASES = [
    # n, h, w, cin, cout, k
    (1, 1, 256, 64, 64, 1),      # plain GEMM, single k-block
    (1, 1, 1000, 128, 64, 1),    # ragged M (not a multiple of 128)
    (2, 16, 16, 256, 256, 1),
    (2, 8, 8, 64, 64, 3),        # tile spans two images
    (3, 8, 8, 64, 128, 3),       # odd image count -> last tile half out of bounds
    (2, 16, 16, 64, 64, 3),
    (1, 32, 32, 128, 64, 3),
    (1, 64, 64, 64, 128, 3),
    (1, 128, 128, 64, 64, 3),
    (1, 256, 256, 64, 64, 3),    # W > 128: two tiles per row
    (2, 32, 32, 512, 512, 3),
]


@pytest.mark.parametrize("n,h,w,cin,cout,k", CASES)
@pytest.mark.parametrize("block_n", [0, 64, 128, 256])
def test_plain(cuda_device, n, h, w, cin, cout, k, block_n):
    from speak_hack_b200 import ops

    if block_n and cout % block_n:
        pytest.skip("BLOCK_N does not divide Cout")
    x, wt, wk = _mk(n, h, w, cin, cout, k, cuda_device)
    y = ops.conv_gemm(x, wk, k, ops.EPI_PLAIN, force_block_n=block_n)
    torch.cuda.synchronize()
    ref = _ref(x, wt, k)
    assert rel_l2(y.float(), ref) < 4e-3  # bf16 output rounding: 2^-9 per element


@pytest.mark.parametrize("n,h,w,cin,cout,k", [(2, 16, 16, 64, 128, 1), (8, 32, 32, 128, 256, 3), (1, 1, 1000, 128, 64, 1)])
def test_plain_with_bias(cuda_device, monkeypatch, n, h, w, cin, cout, k):
    """PLAIN epilogue with a per-channel bias (C ABI: irfd_conv_gemm mode 0 + bias), per-warp and lockstep epilogues
    (ragged M falls back to the lockstep one)."""
    from speak_hack_b200 import ops

    x, wt, wk = _mk(n, h, w, cin, cout, k, cuda_device, seed=7)
    bias = torch.randn(cout, generator=torch.Generator().manual_seed(70)).to(cuda_device)
    ref = _ref(x, wt, k) + bias.view(1, 1, 1, -1)
    for epi in ("1", "0"):
        monkeypatch.setenv("IRFD_WARP_EPI", epi)
        y = ops.conv_gemm(x, wk, k, ops.EPI_PLAIN, bias=bias)
        torch.cuda.synchronize()
        assert rel_l2(y.float(), ref) < 4e-3, epi


@pytest.mark.parametrize("n,h,w,cin,cout,k", [(2, 16, 16, 256, 256, 1), (2, 32, 32, 512, 512, 3), (8, 32, 32, 128, 256, 3),
                                              (12, 64, 64, 64, 256, 1), (6, 16, 16, 1024, 256, 1)])
def test_kernel_variants_bit_identical(cuda_device, monkeypatch, n, h, w, cin, cout, k):
    """The optional instances of conv_gemm_kernel compute the same tiles with the same MMA sequence:
    IRFD_GEMM_CLUSTER=1 (CTA pairs, each CTA multicasts half of every weight tile into both) must give bit-identical
    outputs and statistics; IRFD_WARP_EPI=0 (the lockstep epilogue) bit-identical outputs and statistics equal up to the
    fp32 order of the per-tile row sums (16-row groups instead of 32-row warps)."""
    from speak_hack_b200 import ops

    x, wt, wk = _mk(n, h, w, cin, cout, k, cuda_device, seed=5)
    base = ops.conv_gemm(x, wk, k, ops.EPI_STATS)
    plain = ops.conv_gemm(x, wk, k, ops.EPI_PLAIN)
    torch.cuda.synchronize()
    assert torch.equal(plain, base[0])
    monkeypatch.setenv("IRFD_GEMM_CLUSTER", "1")
    got = ops.conv_gemm(x, wk, k, ops.EPI_STATS)
    got_plain = ops.conv_gemm(x, wk, k, ops.EPI_PLAIN)
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(got, base)) and torch.equal(got_plain, plain)
    if n % 3 == 0:   # grouped weights: pairs never straddle two weight groups
        wk3 = torch.cat([wk, wk.flip(0), wk.roll(1, 0)]).contiguous()
        monkeypatch.setenv("IRFD_GEMM_CLUSTER", "0")
        g0 = ops.conv_gemm_grouped(x, wk3, k, ops.EPI_STATS, wgroups=3)
        monkeypatch.setenv("IRFD_GEMM_CLUSTER", "1")
        g1 = ops.conv_gemm_grouped(x, wk3, k, ops.EPI_STATS, wgroups=3)
        torch.cuda.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(g0, g1))
    monkeypatch.setenv("IRFD_GEMM_CLUSTER", "0")
    monkeypatch.setenv("IRFD_GEMM_2SM", "1")   # CTA pairs issuing 2-SM MMAs (cta_group::2), half a weight tile per SM
    two = ops.conv_gemm(x, wk, k, ops.EPI_STATS)
    two_plain = ops.conv_gemm(x, wk, k, ops.EPI_PLAIN)
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(two, base)) and torch.equal(two_plain, plain)
    monkeypatch.setenv("IRFD_GEMM_2SM", "0")
    monkeypatch.setenv("IRFD_WARP_EPI", "0")
    old = ops.conv_gemm(x, wk, k, ops.EPI_STATS)
    torch.cuda.synchronize()
    assert torch.equal(old[0], base[0])
    assert torch.allclose(old[1], base[1], rtol=1e-5, atol=1e-3) and torch.allclose(old[2], base[2], rtol=1e-5, atol=1e-2)


@pytest.mark.parametrize("n,h,w,cin,cout,k", [(2, 16, 16, 64, 64, 3), (4, 32, 32, 128, 256, 3), (2, 64, 64, 256, 64, 1)])
def test_stats(cuda_device, n, h, w, cin, cout, k):
    from speak_hack_b200 import ops

    x, wt, wk = _mk(n, h, w, cin, cout, k, cuda_device, seed=1)
    y, ssum, ssq = ops.conv_gemm(x, wk, k, ops.EPI_STATS)
    torch.cuda.synchronize()
    ref = _ref(x, wt, k)
    assert rel_l2(y.float(), ref) < 4e-3
    yf = y.float().reshape(-1, cout)
    assert torch.allclose(ssum.sum(0), yf.sum(0), rtol=1e-4, atol=1e-2)
    assert torch.allclose(ssq.sum(0), (yf * yf).sum(0), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 8, 8, 64, 64), (3, 8, 8, 128, 128), (2, 32, 32, 128, 256), (1, 256, 256, 64, 64)])
def test_style(cuda_device, n, h, w, cin, cout):
    from speak_hack_b200 import ops

    dev = cuda_device
    x, wt, wk = _mk(n, h, w, cin, cout, 3, dev, seed=2)
    g = torch.Generator(device="cpu").manual_seed(3)
    bias = torch.randn(cout, generator=g).to(dev)
    nw = torch.randn(cout, generator=g).to(dev)
    noise = torch.randn(n * h * w, generator=g).to(dev)
    sp1 = (torch.randn(n, cout, generator=g) + 1).to(dev)
    s1 = torch.randn(n, cout, generator=g).to(dev)
    a, y = ops.conv_gemm(x, wk, 3, ops.EPI_STYLE, bias=bias, nw=nw, noise=noise, sp1=sp1, s1=s1)
    torch.cuda.synchronize()
    z = _ref(x, wt, 3) + bias.view(1, 1, 1, -1) + nw.view(1, 1, 1, -1) * noise.view(n, h, w, 1)
    a_ref = F.leaky_relu(z, 0.2)
    y_ref = a_ref * sp1.view(n, 1, 1, cout) + s1.view(n, 1, 1, cout)
    assert rel_l2(a.float(), a_ref) < 4e-3
    assert rel_l2(y.float(), y_ref) < 4e-3


WG_CASES = [
    (1, 1, 512, 64, 64, 1),
    (1, 1, 1000, 128, 64, 1),     # ragged pixel count
    (2, 8, 8, 64, 64, 3),
    (3, 8, 8, 128, 128, 3),       # odd atom count? 9*2=18 atoms (even); odd image count
    (2, 16, 16, 64, 256, 3),      # 9 atoms (odd) -> duplicated tail atom
    (4, 32, 32, 128, 64, 3),
    (2, 64, 64, 64, 128, 3),
    (1, 128, 128, 64, 64, 3),
    (2, 16, 16, 512, 512, 3),
    (2, 16, 16, 1024, 256, 1),
]


@pytest.mark.parametrize("n,h,w,cin,cout,k", WG_CASES)
def test_wgrad(cuda_device, n, h, w, cin, cout, k):
    from speak_hack_b200 import ops

    dev = cuda_device
    g = torch.Generator(device="cpu").manual_seed(4)
    x = torch.randn(n, h, w, cin, generator=g).to(dev).to(torch.bfloat16)
    dy = torch.randn(n, h, w, cout, generator=g).to(dev).to(torch.bfloat16)
    dw = ops.conv_wgrad(x, dy, k)
    torch.cuda.synchronize()
    wt = torch.zeros(cout, cin, k, k, device=dev, requires_grad=True)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wt, padding=k // 2)
    (ref,) = torch.autograd.grad(y, wt, dy.float().permute(0, 3, 1, 2))
    assert rel_l2(dw, ref) < 1e-4  # fp32 accumulation of exact bf16 products; only summation order differs
    # accumulate form: dw = 1.0*dw + grad
    dw2 = ops.conv_wgrad(x, dy, k, dw=dw.clone(), beta=1.0)
    torch.cuda.synchronize()
    assert rel_l2(dw2, 2 * ref) < 1e-4


@pytest.mark.parametrize("n,h,w,cin,cout,k", [(2, 16, 16, 64, 256, 1), (3, 8, 8, 128, 128, 3), (1, 64, 64, 64, 64, 3)])
@pytest.mark.parametrize("with_res,relu", [(False, True), (True, True), (True, False)])
def test_affine_epilogue(cuda_device, n, h, w, cin, cout, k, with_res, relu):
    """mode 3: eval-mode BN + ReLU + residual folded into the conv epilogue."""
    from speak_hack_b200 import ops

    dev = cuda_device
    x, wt, wk = _mk(n, h, w, cin, cout, k, dev, seed=5)
    g = torch.Generator(device="cpu").manual_seed(6)
    scale = (torch.rand(cout, generator=g) + 0.5).to(dev)
    shift = torch.randn(cout, generator=g).to(dev)
    res = torch.randn(n, h, w, cout, generator=g).to(dev).to(torch.bfloat16) if with_res else None
    y = ops.conv_gemm_affine(x, wk, k, scale, shift, res=res, relu=relu)
    torch.cuda.synchronize()
    ref = _ref(x, wt, k) * scale.view(1, 1, 1, -1) + shift.view(1, 1, 1, -1)
    if with_res:
        ref = ref + res.float()
    if relu:
        ref = F.relu(ref)
    assert rel_l2(y.float(), ref) < 4e-3


HALO_CASES = [
    # n, h, w, cin, cout — 3x3 convs eligible for the halo-reuse kernel (W % 128 == 0, H even, Cout 64 / 128)
    (1, 2, 128, 64, 64),       # a single 256-pixel tile: every halo row above/below is out of bounds
    (3, 4, 128, 128, 64),      # odd image count, two cin chunks
    (1, 128, 128, 256, 128),   # generator L4.conv1 shape (one image)
    (2, 6, 256, 64, 128),      # two column segments per row: left/right halo columns come from the neighbour segment
    (1, 256, 256, 128, 64),    # generator L5.conv1 shape (one image)
]


@pytest.mark.parametrize("n,h,w,cin,cout", HALO_CASES)
def test_halo_kernel_matches_per_tap_kernel(cuda_device, monkeypatch, n, h, w, cin, cout):
    """The halo-reuse kernel (one [4 x 130] halo per 64-channel chunk serving all nine taps of two output rows) against
    the per-tap kernel and a torch fp32 conv, all four epilogues' building blocks (PLAIN, STATS, STYLE)."""
    from speak_hack_b200 import ops

    dev = cuda_device
    x, wt, wk = _mk(n, h, w, cin, cout, 3, dev, seed=7)
    ref = _ref(x, wt, 3)
    monkeypatch.setenv("IRFD_CONV_HALO", "0")
    y_tap = ops.conv_gemm(x, wk, 3, ops.EPI_PLAIN)
    monkeypatch.setenv("IRFD_CONV_HALO", "1")
    y_halo = ops.conv_gemm(x, wk, 3, ops.EPI_PLAIN)
    ys, ssum, ssq = ops.conv_gemm(x, wk, 3, ops.EPI_STATS)
    g = torch.Generator(device="cpu").manual_seed(8)
    bias = torch.randn(cout, generator=g).to(dev)
    nw = torch.randn(cout, generator=g).to(dev)
    noise = torch.randn(n * h * w, generator=g).to(dev)
    sp1 = (torch.randn(n, cout, generator=g) + 1).to(dev)
    s1 = torch.randn(n, cout, generator=g).to(dev)
    a, y = ops.conv_gemm(x, wk, 3, ops.EPI_STYLE, bias=bias, nw=nw, noise=noise, sp1=sp1, s1=s1)
    torch.cuda.synchronize()
    assert rel_l2(y_tap.float(), ref) < 4e-3
    assert rel_l2(y_halo.float(), ref) < 4e-3
    # same bf16 products, fp32 accumulation in a different order: equal up to the last bf16 rounding of a few outputs
    assert rel_l2(y_halo.float(), y_tap.float()) < 2e-3
    assert torch.equal(ys, y_halo)
    yf = ys.float().reshape(-1, cout)
    assert ssum.shape[0] == (n * h * w) // 128
    assert torch.allclose(ssum.sum(0), yf.sum(0), rtol=1e-4, atol=1e-2)
    assert torch.allclose(ssq.sum(0), (yf * yf).sum(0), rtol=1e-4, atol=1e-2)
    assert torch.allclose(ssum[1], yf[128:256].sum(0), rtol=1e-4, atol=1e-2)  # per-tile rows land in flat tile order
    z = ref + bias.view(1, 1, 1, -1) + nw.view(1, 1, 1, -1) * noise.view(n, h, w, 1)
    a_ref = F.leaky_relu(z, 0.2)
    assert rel_l2(a.float(), a_ref) < 4e-3
    assert rel_l2(y.float(), a_ref * sp1.view(n, 1, 1, cout) + s1.view(n, 1, 1, cout)) < 4e-3


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 16, 16, 64, 64), (3, 8, 32, 128, 64), (1, 4, 64, 64, 192),
                                             (2, 3, 128, 64, 128), (1, 2, 256, 128, 64)])
def test_wgrad_halo_matches_per_tap_kernel(cuda_device, monkeypatch, n, h, w, cin, cout):
    """Halo-reuse wgrad (all nine taps of a Cin chunk accumulated in TMEM from one halo tile per 128 pixels) against
    the per-tap split-K kernel and torch autograd, for every row-width class (16, 32, 64, 128, 256)."""
    from speak_hack_b200 import ops

    dev = cuda_device
    g = torch.Generator(device="cpu").manual_seed(9)
    x = torch.randn(n, h, w, cin, generator=g).to(dev).to(torch.bfloat16)
    dy = torch.randn(n, h, w, cout, generator=g).to(dev).to(torch.bfloat16)
    monkeypatch.setenv("IRFD_WGRAD_HALO", "0")
    dw_tap = ops.conv_wgrad(x, dy, 3)
    monkeypatch.setenv("IRFD_WGRAD_HALO", "2")  # force the halo kernel on every eligible shape
    dw_halo = ops.conv_wgrad(x, dy, 3)
    torch.cuda.synchronize()
    wt = torch.zeros(cout, cin, 3, 3, device=dev, requires_grad=True)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wt, padding=1)
    (ref,) = torch.autograd.grad(y, wt, dy.float().permute(0, 3, 1, 2))
    assert rel_l2(dw_tap, ref) < 1e-4
    assert rel_l2(dw_halo, ref) < 1e-4
    for tap in range(9):  # every tap on its own: a wrong line offset would only corrupt some taps
        assert rel_l2(dw_halo[:, :, tap // 3, tap % 3], ref[:, :, tap // 3, tap % 3]) < 1e-4
