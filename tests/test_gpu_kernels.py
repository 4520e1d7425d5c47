"""-m gpu: every memory-bound kernel of libirfd_b200.so against the torch fp32 op it replaces (same inputs)."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def gen(seed=0):
    return torch.Generator(device="cpu").manual_seed(seed)


def nhwc(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous()


def nchw(x_nhwc):
    return x_nhwc.permute(0, 3, 1, 2).contiguous()


BF16_TOL = 4e-3  # one bf16 rounding of the output (2^-9 per element)


# ------------------------------------------------------------------------------------------------------------ BN
@pytest.mark.parametrize("n,h,w,c", [(4, 16, 16, 64), (2, 8, 8, 2048), (3, 32, 32, 256)])
def test_bn_train_forward_backward(cuda_device, n, h, w, c):
    from speak_hack_b200 import ops

    dev = cuda_device
    g = gen(1)
    z = (torch.randn(n, h, w, c, generator=g) * 2 + 0.5).to(dev).to(BF)
    res = torch.randn(n, h, w, c, generator=g).to(dev).to(BF)
    gamma = (torch.rand(c, generator=g) + 0.5).to(dev)
    beta = torch.randn(c, generator=g).to(dev)
    rm = torch.zeros(c, device=dev)
    rv = torch.ones(c, device=dev)
    rows = n * h * w
    # statistics: emulate the conv-epilogue partials with 128-row tiles computed by torch
    zf = z.float().reshape(rows, c)
    pad = (-rows) % 128
    zp = torch.cat([zf, torch.zeros(pad, c, device=dev)]) if pad else zf
    ssum = zp.reshape(-1, 128, c).sum(1).contiguous()
    ssq = (zp * zp).reshape(-1, 128, c).sum(1).contiguous()
    mean, rstd = ops.bn_finalize(ssum, ssq, rows, 1e-5, 0.1, rm, rv, running_updates=1)
    out = ops.bn_apply(z, mean, rstd, gamma, beta, res=res, relu=True)
    torch.cuda.synchronize()

    zr = nchw(z.float()).requires_grad_(True)
    bn = torch.nn.BatchNorm2d(c).to(dev)
    with torch.no_grad():
        bn.weight.copy_(gamma)
        bn.bias.copy_(beta)
    bn.train()
    pre = bn(zr) + nchw(res.float())
    ref = F.relu(pre)
    assert torch.allclose(mean, zf.mean(0), atol=1e-4, rtol=1e-4)
    assert torch.allclose(rm, bn.running_mean, atol=1e-5, rtol=1e-4)
    assert torch.allclose(rv, bn.running_var, atol=1e-5, rtol=1e-4)
    assert rel_l2(nchw(out.float()), ref) < BF16_TOL

    # backward: upstream gradient dout w.r.t. relu output; mask from the stored (bf16) output like the product does
    dout = torch.randn(n, h, w, c, generator=g).to(dev).to(BF)
    dz, dgamma, dbeta, g_out = ops.bn_backward(dout, None, out, z, mean, rstd, gamma, want_g_out=True)
    torch.cuda.synchronize()
    mask = (out.float() > 0).float()
    gm = nchw(dout.float() * mask)
    (dzr,) = torch.autograd.grad(bn(zr), zr, gm, retain_graph=False)
    xhat = (zf - mean) * rstd
    gflat = (dout.float() * mask).reshape(rows, c)
    assert rel_l2(g_out.float(), dout.float() * mask) < 1e-6
    assert rel_l2(dbeta, gflat.sum(0)) < 1e-4
    assert rel_l2(dgamma, (gflat * xhat).sum(0)) < 1e-3
    assert rel_l2(nchw(dz.float()), dzr) < 6e-3


def test_bn_apply_dual_and_eval(cuda_device):
    from speak_hack_b200 import ops

    dev = cuda_device
    g = gen(2)
    n, h, w, c = 2, 8, 8, 512
    z = torch.randn(n, h, w, c, generator=g).to(dev).to(BF)
    z2 = torch.randn(n, h, w, c, generator=g).to(dev).to(BF)
    p = [torch.randn(c, generator=g).to(dev) for _ in range(4)]
    q = [torch.randn(c, generator=g).to(dev) for _ in range(4)]
    var1, var2 = p[1].abs() + 0.1, q[1].abs() + 0.1
    r1, r2 = ops.bn_eval_rstd(var1, 1e-5), ops.bn_eval_rstd(var2, 1e-5)
    out = ops.bn_apply(z, p[0], r1, p[2], p[3], res=z2, bn2=(q[0], r2, q[2], q[3]), relu=True)
    torch.cuda.synchronize()
    ref = F.relu(
        F.batch_norm(nchw(z.float()), p[0], var1, p[2], p[3], False, 0.0, 1e-5)
        + F.batch_norm(nchw(z2.float()), q[0], var2, q[2], q[3], False, 0.0, 1e-5)
    )
    assert rel_l2(nchw(out.float()), ref) < BF16_TOL


# ------------------------------------------------------------------------------------------------------------ layout
def test_pack_conv_weight(cuda_device):
    from speak_hack_b200 import ops

    dev = cuda_device
    w = torch.randn(128, 64, 3, 3, generator=gen(3)).to(dev)
    wb = w.to(BF)
    fp = ops.pack_conv_weight(w, ops.PACK_FPROP)
    dg = ops.pack_conv_weight(w, ops.PACK_DGRAD)
    dc = ops.pack_conv_weight(w, ops.PACK_DCOL)
    torch.cuda.synchronize()
    assert torch.equal(fp, wb.permute(0, 2, 3, 1).reshape(128, -1))
    assert torch.equal(dg, wb.flip(2, 3).permute(1, 2, 3, 0).reshape(64, -1))
    assert torch.equal(dc, wb.permute(2, 3, 1, 0).reshape(-1, 128))
    ws = torch.randn(64, 3, 7, 7, generator=gen(4)).to(dev)
    fl = ops.pack_conv_weight(ws, ops.PACK_FLAT, kpad=192)
    torch.cuda.synchronize()
    assert torch.equal(fl[:, :147], ws.to(BF).reshape(64, 147)) and float(fl[:, 147:].float().abs().sum()) == 0.0


def test_stem_conv_via_im2col(cuda_device):
    """7x7/2 stem = im2col + plain tcgen05 GEMM (+ wgrad through the same col matrix)."""
    from speak_hack_b200 import ops

    dev = cuda_device
    g = gen(5)
    x = (torch.rand(2, 3, 64, 64, generator=g) * 2 - 1).to(dev)
    w = (torch.randn(64, 3, 7, 7, generator=g) * 0.1).to(dev)
    col = ops.im2col_stem(x, 192)
    wk = ops.pack_conv_weight(w, ops.PACK_FLAT, kpad=192)
    y, ssum, ssq = ops.gemm_rows(col, wk, ops.EPI_STATS)
    torch.cuda.synchronize()
    ref = F.conv2d(x.to(BF).float(), w.to(BF).float(), stride=2, padding=3)
    assert rel_l2(y.float().view(2, 32, 32, 64), nhwc(ref)) < BF16_TOL
    dy = torch.randn(2, 32, 32, 64, generator=g).to(dev).to(BF)
    dw = ops.conv_wgrad(col.view(1, 1, -1, 192), dy.view(1, 1, -1, 64), 1, reduce_cin=147, reduce_taps=1,
                        out_shape=(64, 3, 7, 7))
    torch.cuda.synchronize()
    wr = torch.zeros_like(w, requires_grad=True)
    (dwr,) = torch.autograd.grad(F.conv2d(x.to(BF).float(), wr, stride=2, padding=3), wr, nchw(dy.float()))
    assert rel_l2(dw, dwr) < 1e-4


def test_conv3x3_stride2_path(cuda_device):
    from speak_hack_b200 import ops

    dev = cuda_device
    g = gen(6)
    n, h, w_, c, co = 2, 16, 16, 128, 128
    a = torch.randn(n, h, w_, c, generator=g).to(dev).to(BF)
    w = (torch.randn(co, c, 3, 3, generator=g) * 0.05).to(dev)
    col = ops.im2col_3x3s2(a)
    y = ops.gemm_rows(col, ops.pack_conv_weight(w, ops.PACK_FPROP))
    torch.cuda.synchronize()
    xr = nchw(a.float()).requires_grad_(True)
    wr = w.to(BF).float().requires_grad_(True)
    ref = F.conv2d(xr, wr, stride=2, padding=1)
    assert rel_l2(y.float().view(n, h // 2, w_ // 2, co), nhwc(ref)) < BF16_TOL
    dy = torch.randn(n, h // 2, w_ // 2, co, generator=g).to(dev).to(BF)
    dxr, dwr = torch.autograd.grad(ref, (xr, wr), nchw(dy.float()))
    dcol = ops.gemm_rows(dy.view(-1, co), ops.pack_conv_weight(w, ops.PACK_DCOL))
    dx = ops.col2im_3x3s2(dcol, n, h, w_, c)
    dw = ops.conv_wgrad(col.view(1, 1, -1, 9 * c), dy.view(1, 1, -1, co), 1, reduce_cin=c, reduce_taps=9,
                        out_shape=(co, c, 3, 3))
    torch.cuda.synchronize()
    assert rel_l2(nchw(dx.float()), dxr) < 8e-3  # dcol is rounded to bf16 before the 4-tap gather
    assert rel_l2(dw, dwr) < 1e-4


def test_subsample_scatter(cuda_device):
    from speak_hack_b200 import ops

    dev = cuda_device
    g = gen(7)
    a = torch.randn(2, 8, 8, 64, generator=g).to(dev).to(BF)
    b = torch.randn(2, 4, 4, 64, generator=g).to(dev).to(BF)
    s = ops.subsample2(a)
    o = ops.scatter_add_s2(a, b)
    o2 = ops.scatter_add_s2(None, b)
    torch.cuda.synchronize()
    assert torch.equal(s, a[:, ::2, ::2].contiguous())
    ref = a.float().clone()
    ref[:, ::2, ::2] += b.float()
    assert rel_l2(o.float(), ref) < BF16_TOL
    z = torch.zeros_like(ref)
    z[:, ::2, ::2] = b.float()
    assert torch.equal(o2.float(), z)


def test_pools(cuda_device):
    from speak_hack_b200 import ops

    dev = cuda_device
    g = gen(8)
    a = F.relu(torch.randn(2, 16, 16, 64, generator=g)).to(dev).to(BF)
    out, arg = ops.maxpool_fwd(a)
    torch.cuda.synchronize()
    ar = nchw(a.float()).requires_grad_(True)
    ref = F.max_pool2d(ar, 3, 2, 1)
    assert torch.equal(nchw(out.float()), ref)
    dout = torch.randn(2, 8, 8, 64, generator=g).to(dev).to(BF)
    dx = ops.maxpool_bwd(dout, arg)
    torch.cuda.synchronize()
    (dxr,) = torch.autograd.grad(ref, ar, nchw(dout.float()))
    # ties (zeros after ReLU) may route gradient to a different zero; both are killed by the ReLU mask downstream
    mask = (ar > 0).float()
    assert rel_l2(nchw(dx.float()) * mask, dxr * mask) < BF16_TOL
    f = ops.avgpool_fwd(a)
    torch.cuda.synchronize()
    assert torch.allclose(f, a.float().mean((1, 2)), atol=1e-5, rtol=1e-5)
    df = torch.randn(2, 64, generator=g).to(dev)
    gg = ops.avgpool_bwd(df, 16, 16)
    torch.cuda.synchronize()
    assert rel_l2(gg.float(), (df / 256).view(2, 1, 1, 64).expand(2, 16, 16, 64)) < BF16_TOL


# ------------------------------------------------------------------------------------------------------------ synthesis
@pytest.mark.parametrize("b,h,w,c", [(2, 4, 4, 512), (2, 16, 16, 64), (1, 64, 64, 128)])
def test_upsample(cuda_device, b, h, w, c):
    from speak_hack_b200 import ops

    dev = cuda_device
    g = gen(9)
    x = torch.randn(b, h, w, c, generator=g).to(dev).to(BF)
    up = ops.upsample2x_fwd(x)
    torch.cuda.synchronize()
    xr = nchw(x.float()).requires_grad_(True)
    ref = F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=False)
    assert rel_l2(nchw(up.float()), ref) < BF16_TOL
    dout = torch.randn(b, 2 * h, 2 * w, c, generator=g).to(dev).to(BF)
    din = ops.upsample2x_bwd(dout)
    torch.cuda.synchronize()
    (dr,) = torch.autograd.grad(ref, xr, nchw(dout.float()))
    assert rel_l2(nchw(din.float()), dr) < BF16_TOL


def test_const_input(cuda_device):
    from speak_hack_b200 import ops

    dev = cuda_device
    g = gen(10)
    b, c = 3, 512
    cst = torch.randn(1, c, 4, 4, generator=g).to(dev).requires_grad_(True)
    bias = torch.randn(c, generator=g).to(dev).requires_grad_(True)
    nw = torch.randn(c, generator=g).to(dev).requires_grad_(True)
    noise = torch.randn(b, 1, 4, 4, generator=g).to(dev)
    sp1 = torch.randn(b, c, generator=g).to(dev).requires_grad_(True)
    s1 = torch.randn(b, c, generator=g).to(dev).requires_grad_(True)
    a0, y0 = ops.const_input_fwd(cst.detach(), bias.detach(), nw.detach(), noise.reshape(-1).contiguous(),
                                 sp1.detach(), s1.detach())
    torch.cuda.synchronize()
    ar = cst.expand(b, -1, -1, -1) + bias.view(1, -1, 1, 1) + nw.view(1, -1, 1, 1) * noise
    yr = ar * sp1.view(b, c, 1, 1) + s1.view(b, c, 1, 1)
    assert rel_l2(nchw(y0.float()), yr) < BF16_TOL
    dy = torch.randn(b, 4, 4, c, generator=g).to(dev).to(BF)
    dsp1, ds1, dconst, dbias, dnw = ops.const_input_bwd(dy, a0, noise.reshape(-1).contiguous(), sp1.detach())
    torch.cuda.synchronize()
    grads = torch.autograd.grad(yr, (sp1, s1, cst, bias, nw), nchw(dy.float()))
    for got, ref in zip((dsp1, ds1, dconst, dbias, dnw), grads):
        assert rel_l2(got, ref) < BF16_TOL  # a0 is stored in bf16


@pytest.mark.parametrize("b,h,w,c", [(2, 8, 8, 512), (2, 64, 64, 128), (1, 256, 256, 64)])
def test_style_bwd(cuda_device, b, h, w, c):
    from speak_hack_b200 import ops

    dev = cuda_device
    g = gen(11)
    z = torch.randn(b, h, w, c, generator=g).to(dev)
    a = F.leaky_relu(z, 0.2).to(BF)
    noise = torch.randn(b * h * w, generator=g).to(dev)
    sp1 = (torch.randn(b, c, generator=g) + 1).to(dev)
    dy = torch.randn(b, h, w, c, generator=g).to(dev).to(BF)
    dz, ds1, dsp1, dbias, dnw = ops.style_bwd(dy, a, noise, sp1)
    torch.cuda.synchronize()
    dyf, af = dy.float(), a.float()
    dz_ref = dyf * sp1.view(b, 1, 1, c) * torch.where(af > 0, 1.0, 0.2)
    assert rel_l2(dz.float(), dz_ref) < BF16_TOL
    assert rel_l2(ds1, dyf.sum((1, 2))) < 1e-4
    assert rel_l2(dsp1, (dyf * af).sum((1, 2))) < 1e-4
    assert rel_l2(dbias, dz_ref.sum((0, 1, 2))) < 1e-4
    assert rel_l2(dnw, (dz_ref * noise.view(b, h, w, 1)).sum((0, 1, 2))) < 1e-3


@pytest.mark.parametrize("b,h,w,c", [(2, 32, 32, 64), (1, 256, 256, 64), (2, 16, 16, 32)])
def test_to_rgb(cuda_device, b, h, w, c):
    from speak_hack_b200 import ops

    dev = cuda_device
    g = gen(12)
    y = torch.randn(b, h, w, c, generator=g).to(dev).to(BF)
    wt = torch.randn(3, c, 1, 1, generator=g).to(dev).requires_grad_(True)
    bias = torch.randn(3, generator=g).to(dev).requires_grad_(True)
    out = ops.to_rgb_fwd(y, wt.detach(), bias.detach())
    torch.cuda.synchronize()
    yr = nchw(y.float()).requires_grad_(True)
    ref = F.conv2d(yr, wt, bias)
    assert rel_l2(out, ref) < 1e-5
    drgb = torch.randn(b, 3, h, w, generator=g).to(dev)
    dy, dw, db = ops.to_rgb_bwd(drgb, y, wt.detach())
    torch.cuda.synchronize()
    dyr, dwr, dbr = torch.autograd.grad(ref, (yr, wt, bias), drgb)
    assert rel_l2(nchw(dy.float()), dyr) < BF16_TOL
    assert rel_l2(dw, dwr) < 1e-4
    assert rel_l2(db, dbr) < 1e-4


# ------------------------------------------------------------------------------------------------------------ dense
@pytest.mark.parametrize("b,n,k", [(2, 512, 6144), (32, 512, 512), (64, 128, 512), (5, 8, 2048), (130, 64, 512)])
def test_linear(cuda_device, b, n, k):
    from speak_hack_b200 import ops

    dev = cuda_device
    g = gen(13)
    x = torch.randn(b, k, generator=g).to(dev).requires_grad_(True)
    w = torch.randn(n, k, generator=g).to(dev).requires_grad_(True)
    bias = torch.randn(n, generator=g).to(dev).requires_grad_(True)
    wmul, bmul = 0.37, 0.01
    y = ops.linear_fwd(x.detach(), w.detach(), bias.detach(), wmul, bmul, lrelu=True)
    torch.cuda.synchronize()
    ref = F.leaky_relu(F.linear(x, w * wmul, bias * bmul), 0.2)
    assert rel_l2(y, ref) < 1e-5
    dy = torch.randn(b, n, generator=g).to(dev)
    dz = ops.lrelu_bwd(dy, y)
    dx, dw, db = ops.linear_bwd(dz, x.detach(), w.detach(), wmul, bmul)
    torch.cuda.synchronize()
    dxr, dwr, dbr = torch.autograd.grad(ref, (x, w, bias), dy)
    assert rel_l2(dx, dxr) < 1e-5
    assert rel_l2(dw, dwr) < 1e-5
    assert rel_l2(db, dbr) < 1e-5
    # accumulate form of dx
    dx2, _, _ = ops.linear_bwd(dz, None, w.detach(), wmul, bmul, dx=dx.clone(), dx_beta=1.0, need_dw=False)
    torch.cuda.synchronize()
    assert rel_l2(dx2, 2 * dxr) < 1e-5
    sm = ops.softmax_rows(y[:, :8].contiguous())
    torch.cuda.synchronize()
    assert torch.allclose(sm, torch.softmax(y[:, :8], 1), atol=1e-6)


def test_style_split_merge_scale(cuda_device):
    from speak_hack_b200 import ops

    dev = cuda_device
    st = torch.randn(3, 256, generator=gen(14)).to(dev)
    sp1, s1 = ops.split_style(st, 128)
    d = ops.merge_style_grad(sp1, s1)
    sc = ops.scale_copy(st, 0.7)
    torch.cuda.synchronize()
    assert torch.equal(sp1, st[:, :128] + 1) and torch.equal(s1, st[:, 128:])
    assert torch.equal(d[:, :128], sp1) and torch.equal(d[:, 128:], s1)
    assert torch.allclose(sc, st * 0.7)


# ------------------------------------------------------------------------------------------------------------ loss / optim
def test_mse_and_adam(cuda_device):
    from speak_hack_b200 import ops

    dev = cuda_device
    g = gen(15)
    a = (torch.randn(2, 3, 256, 256, generator=g) * 100).to(dev).requires_grad_(True)
    b = torch.randn(2, 3, 256, 256, generator=g).to(dev)
    loss = ops.mse_fwd(a.detach(), b)
    torch.cuda.synchronize()
    ref = F.mse_loss(a, b)
    assert abs(loss.item() - ref.item()) <= 1e-6 * abs(ref.item())
    gs = torch.full((1,), 0.5, device=dev)
    da, db = ops.mse_bwd(a.detach(), b, gs, need_da=True, need_db=True)
    torch.cuda.synchronize()
    (dar,) = torch.autograd.grad(ref * 0.5, a)
    assert rel_l2(da, dar) < 1e-6 and rel_l2(db, -dar) < 1e-6
    # Adam with clipping against torch.optim.Adam + clip_grad_norm_
    p = torch.randn(10007, generator=g).to(dev)
    grads = [torch.randn(10007, generator=g).to(dev) * 3 for _ in range(3)]
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=2e-4)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step, gr in enumerate(grads, 1):
        pr.grad = gr.clone()
        torch.nn.utils.clip_grad_norm_([pr], 1.0)
        opt.step()
        ss = ops.sumsq(gr)
        ops.adam_step(p, gr, m, v, 2e-4, 0.9, 0.999, 1e-8, step, total_sumsq=ss, max_norm=1.0)
    torch.cuda.synchronize()
    assert torch.allclose(p, pr.detach(), atol=1e-6, rtol=1e-5)


def test_bn_groups_equal_separate_calls(cuda_device):
    """Two statistic groups in one launch == two launches (forward outputs, running buffers, dz, dgamma/dbeta)."""
    from speak_hack_b200 import ops

    dev = cuda_device
    g = gen(20)
    n, h, w, c = 4, 8, 8, 256  # 2 groups x 128 rows
    rows = n * h * w
    z = (torch.randn(n, h, w, c, generator=g) * 2 + 0.3).to(dev).to(BF)
    res = torch.randn(n, h, w, c, generator=g).to(dev).to(BF)
    dout = torch.randn(n, h, w, c, generator=g).to(dev).to(BF)
    gamma = (torch.rand(c, generator=g) + 0.5).to(dev)
    beta = torch.randn(c, generator=g).to(dev)
    zf = z.float().reshape(rows, c)
    ssum = zf.reshape(-1, 128, c).sum(1).contiguous()
    ssq = (zf * zf).reshape(-1, 128, c).sum(1).contiguous()
    rm2, rv2 = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    mean2, rstd2 = ops.bn_finalize(ssum, ssq, rows // 2, 1e-5, 0.1, rm2, rv2, 1, groups=2)
    out2 = ops.bn_apply(z, mean2, rstd2, gamma, beta, res=res, relu=True, groups=2)
    dz2, dg2, db2, go2 = ops.bn_backward(dout, None, out2, z, mean2, rstd2, gamma, want_g_out=True, groups=2)
    rm1, rv1 = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    dg1, db1 = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
    for gi in range(2):
        sl = slice(gi * 2, gi * 2 + 2)
        m, r = ops.bn_finalize(ssum[gi:gi + 1].contiguous(), ssq[gi:gi + 1].contiguous(), rows // 2, 1e-5, 0.1, rm1, rv1)
        assert torch.equal(m, mean2[gi]) and torch.equal(r, rstd2[gi])
        o = ops.bn_apply(z[sl].contiguous(), m, r, gamma, beta, res=res[sl].contiguous(), relu=True)
        assert torch.equal(o, out2[sl])
        dz, dg, db, go = ops.bn_backward(dout[sl].contiguous(), None, o, z[sl].contiguous(), m, r, gamma, want_g_out=True)
        assert torch.equal(dz, dz2[sl]) and torch.equal(go, go2[sl])
        dg1 += dg
        db1 += db
    torch.cuda.synchronize()
    assert torch.allclose(rm1, rm2, rtol=1e-6, atol=1e-7) and torch.allclose(rv1, rv2, rtol=1e-6, atol=1e-7)
    assert torch.allclose(dg1, dg2, rtol=1e-5, atol=1e-5) and torch.allclose(db1, db2, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("has_g2", [False, True])
@pytest.mark.parametrize("mask", ["none", "act", "from_z"])
@pytest.mark.parametrize("rows,c", [(315, 64), (5000, 128), (777, 1024)])
def test_bn_backward_variants_ragged(cuda_device, has_g2, mask, rows, c):
    """Every (second gradient, ReLU-mask source) instance of the batched BN backward on row counts that are not a
    multiple of the 4-row batch or of the rows a block covers; statistics from many (ragged) 128-row tiles."""
    from speak_hack_b200 import ops

    dev = cuda_device
    g = gen(31)
    z = (torch.randn(1, rows, 1, c, generator=g) * 1.5 + 0.2).to(dev).to(BF)
    gamma = (torch.rand(c, generator=g) + 0.5).to(dev)
    beta = (torch.randn(c, generator=g) * 0.5).to(dev)
    zf = z.float().reshape(rows, c)
    pad = (-rows) % 128
    zp = torch.cat([zf, torch.zeros(pad, c, device=dev)]) if pad else zf
    ssum = zp.reshape(-1, 128, c).sum(1).contiguous()
    ssq = (zp * zp).reshape(-1, 128, c).sum(1).contiguous()
    mean, rstd = ops.bn_finalize(ssum, ssq, rows, 1e-5, 0.1)
    assert torch.allclose(mean, zf.mean(0), atol=1e-4, rtol=1e-4)
    assert torch.allclose(rstd, (zf.var(0, unbiased=False) + 1e-5).rsqrt(), atol=1e-4, rtol=1e-3)
    out = ops.bn_apply(z, mean, rstd, gamma, beta, relu=(mask != "none"))
    xhat = (zf - mean) * rstd
    assert rel_l2(out.float().reshape(rows, c), F.relu(xhat * gamma + beta) if mask != "none" else xhat * gamma + beta) < BF16_TOL
    d1 = torch.randn(1, rows, 1, c, generator=g).to(dev).to(BF)
    d2 = torch.randn(1, rows, 1, c, generator=g).to(dev).to(BF) if has_g2 else None
    dz, dgamma, dbeta, g_out = ops.bn_backward(
        d1, d2, out if mask == "act" else None, z, mean, rstd, gamma, want_g_out=True,
        beta=beta if mask == "from_z" else None)
    torch.cuda.synchronize()
    gsum = d1.float() + (d2.float() if has_g2 else 0)
    if mask == "act":
        gsum = gsum * (out.float() > 0)
    elif mask == "from_z":
        gsum = gsum * ((xhat * gamma + beta).reshape(1, rows, 1, c) > 0)
    gf = gsum.reshape(rows, c)
    ref_db, ref_dg = gf.sum(0), (gf * xhat).sum(0)
    ref_dz = gamma * rstd * (gf - ref_db / rows - xhat * ref_dg / rows)
    if mask != "from_z":  # the recomputed mask may flip where gamma*xhat+beta rounds across 0; compare those loosely
        assert rel_l2(g_out.float().reshape(rows, c), gf) < BF16_TOL
    assert rel_l2(dbeta, ref_db) < 2e-2 if mask == "from_z" else rel_l2(dbeta, ref_db) < 1e-4
    assert rel_l2(dgamma, ref_dg) < 2e-2 if mask == "from_z" else rel_l2(dgamma, ref_dg) < 1e-3
    assert rel_l2(dz.float().reshape(rows, c), ref_dz) < (3e-2 if mask == "from_z" else 6e-3)


@pytest.mark.parametrize("b,h,w,c", [(1, 1, 1, 64), (2, 3, 5, 72), (1, 128, 128, 64)])
def test_upsample_borders(cuda_device, b, h, w, c):
    """Bilinear x2 on odd / degenerate sizes (every border case of the closed-form weights) and the 256^2 layer shape."""
    from speak_hack_b200 import ops

    dev = cuda_device
    g = gen(33)
    x = torch.randn(b, h, w, c, generator=g).to(dev).to(BF)
    up = ops.upsample2x_fwd(x)
    xr = nchw(x.float()).requires_grad_(True)
    ref = F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=False)
    assert rel_l2(nchw(up.float()), ref) < BF16_TOL
    dout = torch.randn(b, 2 * h, 2 * w, c, generator=g).to(dev).to(BF)
    din = ops.upsample2x_bwd(dout)
    torch.cuda.synchronize()
    (dr,) = torch.autograd.grad(ref, xr, nchw(dout.float()))
    assert rel_l2(nchw(din.float()), dr) < BF16_TOL


def test_standalone_generator_modules(cuda_device):
    """ApplyNoise / ApplyStyle / SynthesisBlock called directly (reference-visible classes, styleganv1.py:448-468,
    612-635): same arithmetic as the oracle's modules, forward only; a differentiable call raises instead of returning a
    tensor without a graph."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import irfd_oracle as O
    import speak_hack_b200 as P
    from speak_hack_b200._lib import IrfdError

    dev = cuda_device
    g = torch.Generator().manual_seed(12)
    x = torch.randn(2, 64, 16, 16, generator=g).to(dev)
    an = P.ApplyNoise(64).to(dev)
    with torch.no_grad():
        an.weight.copy_(torch.randn(64, generator=g))
    noise = torch.randn(2, 1, 16, 16, generator=g).to(dev)
    with torch.no_grad():
        out = an(x, noise)
    assert torch.allclose(out, x + an.weight.view(1, -1, 1, 1) * noise, rtol=1e-6, atol=1e-6)
    torch.manual_seed(5)
    with torch.no_grad():
        out2 = an(x, None)   # noise drawn like the reference: torch.randn(B,1,H,W) on the device generator
    torch.manual_seed(5)
    ref_noise = torch.randn(2, 1, 16, 16, device=dev)
    assert torch.allclose(out2, x + an.weight.view(1, -1, 1, 1) * ref_noise, rtol=1e-6, atol=1e-6)
    with pytest.raises(IrfdError):
        an(x.clone().requires_grad_(True), noise)

    torch.manual_seed(1)
    ref_style = O.ApplyStyleRef(512, 64)
    st = P.ApplyStyle(512, 64, use_wscale=True)
    st.load_state_dict(ref_style.state_dict())
    st = st.to(dev)
    lat = torch.randn(2, 512, generator=g)
    with torch.no_grad():
        want = ref_style(x.cpu(), lat)
        got = st(x, lat.to(dev))
    assert rel_l2(got.cpu(), want) < 1e-5

    torch.manual_seed(2)
    ref_blk = O.SynthesisBlockRef(128, 64)
    O.perturb_noise_weights(ref_blk)   # noise1/noise2 weights away from zero
    blk = P.SynthesisBlock(128, 64, 5)
    blk.load_state_dict(ref_blk.state_dict(), strict=False)
    blk = blk.to(dev)
    xin = torch.randn(2, 128, 16, 16, generator=g)
    w = torch.randn(2, 2, 512, generator=g) * 0.3
    bank = []

    def draw(b, h, wd, device, dtype):
        t = torch.randn(b, 1, h, wd, generator=g)
        bank.append(t)
        return t

    with torch.no_grad():
        want = ref_blk(xin, w, draw)
    it = iter(bank)
    orig_randn = torch.randn
    torch.randn = lambda *a, **k: next(it).to(k.get("device", "cpu"))   # feed the block the oracle's noise planes
    try:
        with torch.no_grad():
            got = blk(xin.to(dev), w.to(dev))
    finally:
        torch.randn = orig_randn
    torch.cuda.synchronize()
    assert got.shape == (2, 64, 32, 32) and got.dtype == torch.float32
    e = rel_l2(got.cpu(), want)
    print(f"[parity] standalone SynthesisBlock vs oracle rel-L2 {e:.3e}")
    assert e < 8e-3
