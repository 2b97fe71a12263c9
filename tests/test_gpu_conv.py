"""GPU unit tests of the im2col-free tcgen05 convolution kernels (csrc/conv_tc.cu) and the fused BatchNorm/ELU kernels
(csrc/image_fused.cu) through the C-ABI, against torch float64 on the CPU (MaskedConv2d / PixelCNNBlock semantics:
dec_pixelcnn_v2.py:12-62)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _be():
    import lagvae._backend as be
    return be


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _mask(co, ci, k, mode):
    m = torch.ones(co, ci, k, k, dtype=torch.float64)
    if mode:
        m[:, :, k // 2, k // 2 + (1 if mode == 2 else 0):] = 0
        m[:, :, k // 2 + 1:] = 0
    return m


def _cat(be, x):
    Cc = x.shape[-1]
    cat = torch.empty(*x.shape[:-1], 2 * Cc, dtype=torch.bfloat16, device="cuda")
    be.check(be.lib().lagvae_split_cat(be.ptr(x), x.numel() // Cc, Cc, be.ptr(cat), _st()))
    return cat


@pytest.mark.parametrize("Cc", [32, 64])
def test_split_cat(Cc):
    be = _be()
    x = torch.randn(1000, Cc, device="cuda") * 3
    cat = _cat(be, x)
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    assert torch.equal(cat[:, :Cc], hi) and torch.equal(cat[:, Cc:], lo)


@pytest.mark.parametrize("Cin,Cout,k,mode", [(32, 32, 7, 2), (32, 32, 5, 2), (32, 32, 3, 2), (32, 32, 3, 0), (32, 32, 7, 1),
                                            (64, 32, 1, 0), (32, 64, 1, 0), (64, 64, 1, 0), (64, 64, 3, 0)])
@pytest.mark.parametrize("B", [3, 64])
def test_convtc_forward_dgrad_wgrad(Cin, Cout, k, mode, B):
    be = _be()
    L = be.lib()
    H = W = 28
    assert L.lagvae_convtc_supported(B, H, W, Cin, Cout, k, k) == 1
    g = torch.Generator().manual_seed(100 * k + mode + B + Cin + 3 * Cout)
    x = torch.randn(B, H, W, Cin, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) * 0.1
    dy = torch.randn(B, H, W, Cout, generator=g)
    add_y = torch.randn(B, H, W, Cout, generator=g)
    add_x = torch.randn(B, H, W, Cin, generator=g)
    wm = w.double() * _mask(Cout, Cin, k, mode)
    # float64 reference (NCHW)
    xr = x.double().permute(0, 3, 1, 2).requires_grad_(True)
    wr = wm.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, padding=k // 2)
    yr.backward(dy.double().permute(0, 3, 1, 2))
    y_ref = yr.detach().permute(0, 2, 3, 1)
    dx_ref = xr.grad.permute(0, 2, 3, 1)
    dw_ref = wr.grad                                   # gradients of ALL taps, masked ones included

    xd, dyd, wd = x.cuda(), dy.cuda(), wm.float().cuda().contiguous()
    xcat, dycat = _cat(be, xd), _cat(be, dyd)
    wbuf = torch.empty(L.lagvae_convtc_wbuf_bytes(Cin, Cout, k, k), dtype=torch.uint8, device="cuda")
    be.check(L.lagvae_convtc_prepare_weights(be.ptr(wd), Cout, Cin, k, k, mode, be.ptr(wbuf), _st()))
    y = torch.full((B, H, W, Cout), float("nan"), device="cuda")
    stats = torch.empty(2 * Cout, dtype=torch.float64, device="cuda")
    be.check(L.lagvae_convtc_forward(be.ptr(xcat), be.ptr(wbuf), B, H, W, Cin, Cout, k, k, mode, None, be.ptr(y), be.ptr(stats), _st()))
    torch.cuda.synchronize()
    err = float((y.double().cpu() - y_ref).abs().max() / y_ref.abs().max())
    assert err < 2e-5, ("forward", err)
    yy = y.double().reshape(-1, Cout)
    assert torch.allclose(stats[:Cout], yy.sum(0), rtol=1e-5, atol=1e-2) and torch.allclose(stats[Cout:], (yy * yy).sum(0), rtol=1e-5)
    # with an addend: the statistics are those of the stored value
    y2 = torch.empty_like(y)
    ad = add_y.cuda()
    be.check(L.lagvae_convtc_forward(be.ptr(xcat), be.ptr(wbuf), B, H, W, Cin, Cout, k, k, mode, be.ptr(ad), be.ptr(y2), be.ptr(stats), _st()))
    torch.cuda.synchronize()
    assert torch.allclose(y2, y + ad, rtol=1e-6, atol=1e-6)
    assert torch.allclose(stats[:Cout], y2.double().reshape(-1, Cout).sum(0), rtol=1e-5, atol=1e-2)

    dx = torch.full((B, H, W, Cin), float("nan"), device="cuda")
    adx = add_x.cuda()
    be.check(L.lagvae_convtc_dgrad(be.ptr(dycat), be.ptr(wbuf), B, H, W, Cin, Cout, k, k, mode, be.ptr(adx), be.ptr(dx), _st()))
    torch.cuda.synchronize()
    err = float((dx.double().cpu() - add_x.double() - dx_ref).abs().max() / dx_ref.abs().max())
    assert err < 2e-5, ("dgrad", err)

    dw = torch.full((Cout, Cin, k, k), float("nan"), device="cuda")
    sc = torch.empty(L.lagvae_convtc_wgrad_scratch_bytes(Cin, Cout, k, k), dtype=torch.uint8, device="cuda")
    be.check(L.lagvae_convtc_wgrad(be.ptr(dycat), be.ptr(xcat), B, H, W, Cin, Cout, k, k, be.ptr(dw), be.ptr(sc), _st()))
    torch.cuda.synchronize()
    err = float((dw.double().cpu() - dw_ref).abs().max() / dw_ref.abs().max())
    assert err < 2e-5, ("wgrad", err)


def test_convtc_unsupported_geometry_is_refused():
    be = _be()
    L = be.lib()
    assert L.lagvae_convtc_supported(4, 28, 28, 32, 32, 9, 9) == 0      # > 49 taps
    assert L.lagvae_convtc_supported(4, 28, 28, 32, 32, 4, 4) == 0      # even kernel
    assert L.lagvae_convtc_supported(4, 7, 200, 32, 32, 3, 3) == 0      # a row does not fit a 128-row tile
    assert L.lagvae_convtc_supported(4, 28, 28, 5, 64, 7, 7) == 0       # channel counts other than 32 / 64


@pytest.mark.parametrize("Cc,elu,use_res,from_cat", [(32, 1, 0, 1), (64, 1, 1, 0), (64, 0, 0, 0), (32, 1, 0, 0)])
def test_bnact_forward_backward(Cc, elu, use_res, from_cat):
    """BatchNorm2d(train) [+ residual] [+ ELU] forward (fp32 + cat outputs, running statistics) and backward against
    torch float64 autograd."""
    be = _be()
    L = be.lib()
    R = 3 * 28 * 28
    g = torch.Generator().manual_seed(Cc + elu + 7 * use_res)
    y = torch.randn(R, Cc, generator=g) * 1.7 + 0.3
    res = torch.randn(R, Cc, generator=g) if use_res else None
    gamma, beta = torch.rand(Cc, generator=g) + 0.5, torch.randn(Cc, generator=g) * 0.2
    dout = torch.randn(R, Cc, generator=g)
    rm0, rv0 = torch.randn(Cc, generator=g) * 0.1, torch.rand(Cc, generator=g) + 0.5
    # reference
    yr, gr, br = y.double().requires_grad_(True), gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rr = res.double().requires_grad_(True) if use_res else None
    rm, rv = rm0.double().clone(), rv0.double().clone()
    o = F.batch_norm(yr, rm, rv, gr, br, True, 0.1, 1e-5)
    if use_res:
        o = o + rr
    if elu:
        o = F.elu(o)
    o.backward(dout.double())

    yd = y.cuda()
    stats = torch.cat([yd.double().sum(0), (yd.double() ** 2).sum(0)]).contiguous()
    out = torch.empty(R, Cc, device="cuda")
    cat = torch.empty(R, 2 * Cc, dtype=torch.bfloat16, device="cuda")
    sm, si = torch.empty(Cc, device="cuda"), torch.empty(Cc, device="cuda")
    rmd, rvd = rm0.cuda(), rv0.cuda()
    gd, bd = gamma.cuda(), beta.cuda()
    resd = res.cuda() if use_res else None
    be.check(L.lagvae_bnact_fwd(be.ptr(yd), be.ptr(stats), R, Cc, be.ptr(gd), be.ptr(bd), 1e-5, 0.1, be.ptr(resd), elu, be.ptr(out),
                                be.ptr(cat), be.ptr(sm), be.ptr(si), be.ptr(rmd), be.ptr(rvd), _st()))
    torch.cuda.synchronize()
    assert float((out.double().cpu() - o.detach()).abs().max()) < 2e-5 * float(o.detach().abs().max())
    assert float(((cat[:, :Cc].float() + cat[:, Cc:].float()) - out).abs().max()) < 1e-5 * float(out.abs().max())
    assert torch.allclose(rmd.double().cpu(), rm, rtol=1e-5, atol=1e-6) and torch.allclose(rvd.double().cpu(), rv, rtol=1e-5)

    doutd = dout.cuda()
    dy = torch.empty(R, Cc, device="cuda")
    dycat = torch.empty(R, 2 * Cc, dtype=torch.bfloat16, device="cuda")
    dres = torch.empty(R, Cc, device="cuda")
    dg, db = torch.empty(Cc, device="cuda"), torch.empty(Cc, device="cuda")
    sc = torch.empty(16 * Cc + 256, dtype=torch.uint8, device="cuda")
    be.check(L.lagvae_bnact_bwd(be.ptr(doutd), None if from_cat else be.ptr(out), be.ptr(cat) if from_cat else None, be.ptr(yd), R, Cc,
                                be.ptr(gd), be.ptr(sm), be.ptr(si), elu, be.ptr(dy), be.ptr(dycat), be.ptr(dres), be.ptr(dg), be.ptr(db),
                                be.ptr(sc), _st()))
    torch.cuda.synchronize()
    scale = float(yr.grad.abs().max())
    assert float((dy.double().cpu() - yr.grad).abs().max()) < 5e-5 * scale
    assert float(((dycat[:, :Cc].float() + dycat[:, Cc:].float()) - dy).abs().max()) < 1e-5 * float(dy.abs().max())
    assert torch.allclose(dg.double().cpu(), gr.grad, rtol=2e-4, atol=1e-3) and torch.allclose(db.double().cpu(), br.grad, rtol=2e-4, atol=1e-3)
    if use_res:
        assert float((dres.double().cpu() - rr.grad).abs().max()) < 1e-5 * float(rr.grad.abs().max())
