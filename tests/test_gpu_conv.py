"""GPU unit tests of the im2col-free tcgen05 masked-convolution kernels (csrc/conv_tc.cu) through the C-ABI, against
torch's float64 CPU convolution (MaskedConv2d semantics: dec_pixelcnn_v2.py:12-30)."""
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


def _mask(k, mode):
    m = torch.ones(32, 32, k, k, dtype=torch.float64)
    if mode:
        m[:, :, k // 2, k // 2 + (1 if mode == 2 else 0):] = 0
        m[:, :, k // 2 + 1:] = 0
    return m


def _cat(be, x):
    cat = torch.empty(*x.shape[:-1], 64, dtype=torch.bfloat16, device="cuda")
    be.check(be.lib().lagvae_split_cat32(be.ptr(x), x.numel() // 32, be.ptr(cat), _st()))
    return cat


def test_split_cat32():
    be = _be()
    x = torch.randn(1000, 32, device="cuda") * 3
    cat = _cat(be, x)
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    assert torch.equal(cat[:, :32], hi) and torch.equal(cat[:, 32:], lo)


@pytest.mark.parametrize("k,mode", [(7, 2), (5, 2), (3, 2), (3, 0), (7, 1)])
@pytest.mark.parametrize("B", [3, 64])
def test_conv32_forward_dgrad_wgrad(k, mode, B):
    be = _be()
    L = be.lib()
    H = W = 28
    assert L.lagvae_conv32_supported(B, H, W, k, k) == 1
    g = torch.Generator().manual_seed(100 * k + mode + B)
    x = torch.randn(B, H, W, 32, generator=g)
    w = torch.randn(32, 32, k, k, generator=g) * 0.1
    dy = torch.randn(B, H, W, 32, generator=g)
    wm = w.double() * _mask(k, mode)
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
    wbuf = torch.empty(L.lagvae_conv32_wbuf_bytes(k, k), dtype=torch.uint8, device="cuda")
    be.check(L.lagvae_conv32_prepare_weights(be.ptr(wd), k, k, mode, be.ptr(wbuf), _st()))
    y = torch.full((B, H, W, 32), float("nan"), device="cuda")
    stats = torch.empty(64, dtype=torch.float64, device="cuda")
    be.check(L.lagvae_conv32_forward(be.ptr(xcat), be.ptr(wbuf), B, H, W, k, k, mode, be.ptr(y), be.ptr(stats), _st()))
    torch.cuda.synchronize()
    err = float((y.double().cpu() - y_ref).abs().max() / y_ref.abs().max())
    assert err < 2e-5, ("forward", err)
    yy = y.double().reshape(-1, 32)
    assert torch.allclose(stats[:32], yy.sum(0), rtol=1e-5, atol=1e-3) and torch.allclose(stats[32:], (yy * yy).sum(0), rtol=1e-5)

    dx = torch.full((B, H, W, 32), float("nan"), device="cuda")
    be.check(L.lagvae_conv32_dgrad(be.ptr(dycat), be.ptr(wbuf), B, H, W, k, k, mode, be.ptr(dx), _st()))
    torch.cuda.synchronize()
    err = float((dx.double().cpu() - dx_ref).abs().max() / dx_ref.abs().max())
    assert err < 2e-5, ("dgrad", err)

    dw = torch.full((32, 32, k, k), float("nan"), device="cuda")
    sc = torch.empty(L.lagvae_conv32_wgrad_scratch_bytes(k, k), dtype=torch.uint8, device="cuda")
    be.check(L.lagvae_conv32_wgrad(be.ptr(dycat), be.ptr(xcat), B, H, W, k, k, be.ptr(dw), be.ptr(sc), _st()))
    torch.cuda.synchronize()
    err = float((dw.double().cpu() - dw_ref).abs().max() / dw_ref.abs().max())
    assert err < 2e-5, ("wgrad", err)


def test_conv32_unsupported_geometry_is_refused():
    be = _be()
    assert be.lib().lagvae_conv32_supported(4, 28, 28, 9, 9) == 0      # > 49 taps
    assert be.lib().lagvae_conv32_supported(4, 28, 28, 4, 4) == 0      # even kernel
    assert be.lib().lagvae_conv32_supported(4, 7, 200, 3, 3) == 0      # a row does not fit a 128-row tile
