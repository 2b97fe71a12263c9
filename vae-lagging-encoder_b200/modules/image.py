"""Drop-in image modules: ResNetEncoderV2 / PixelCNNDecoderV2 with the reference's module tree (hence identical
`state_dict` keys: `encoder.main.0.main.1.conv1.weight`, `decoder.main.0.direct_connects.3.main.3.mask`, ...),
constructors and method contracts (reference modules/encoders/enc_resnet_v2.py:27-126,
modules/decoders/dec_pixelcnn_v2.py:12-195), computing on liblagvae.so kernels.

Layout: activations travel between layers as contiguous NHWC fp32 tensors [B, H, W, C] — the matrix
[B*H*W, C] — so 1x1 convolutions are plain GEMMs and k x k convolutions are im2col + GEMM (include/lagvae.h,
"Image path").  Every layer is a torch.autograd.Function whose forward and backward are C-ABI kernel launches;
torch provides the tape, device memory and the (tiny) weight re-layouts only.  No CPU fallback."""
import ctypes as C
import math
import os

import torch
import torch.nn as nn

from lagvae import LagvaeError
from lagvae import _backend as be

from .text import DecoderBase, GaussianEncoderBase

_SCRATCH = {}


def _scratch(nbytes, tag, device):
    key = (tag, device.index if device.index is not None else torch.cuda.current_device())
    t = _SCRATCH.get(key)
    if t is None or t.numel() < nbytes:
        from lagvae.graph import retire
        retire(t)                      # a live CUDA graph may still hold pointers into the outgrown buffer (lagvae/graph.py)
        t = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
        _SCRATCH[key] = t
    return t


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(t, what):
    if t.device.type != "cuda":
        raise LagvaeError("%s: modules.* run on a CUDA (B200) device only — no CPU fallback; got %s" % (what, t.device))


def _gemm(A, a_rs, a_cs, Bm, b_rs, b_cs, out, ldc, M, N, K, alpha=1.0, beta=0.0, bias=None):
    """out[M,N] = alpha * sum_k A(m,k) B(n,k) + beta*out + bias  (lagvae_gemm_auto)."""
    nb = be.lib().lagvae_gemm_auto_scratch_bytes(M, N, K)
    sc = _scratch(nb, "gemm", out.device)
    be.check(be.lib().lagvae_gemm_auto(be.ptr(A), a_rs, a_cs, be.ptr(Bm), b_rs, b_cs, be.ptr(out), ldc, M, N, K, float(alpha),
                                       float(beta), be.ptr(bias), be.ptr(sc), sc.numel(), _st()), "lagvae_gemm_auto")


# ------------------------------------------------------------------------------------------------------
# layer functions
# ------------------------------------------------------------------------------------------------------
class _ConvFn(torch.autograd.Function):
    """nn.Conv2d(bias=False) on NHWC activations: im2col (identity for 1x1/s1) + GEMM; backward = 2 GEMMs + col2im."""

    @staticmethod
    def forward(ctx, x, weight, stride, pad):
        _need_cuda(x, "conv2d")
        x = x.contiguous()
        B, H, W, Cin = x.shape
        Cout, _, kh, kw = weight.shape
        Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
        R, K = B * Ho * Wo, kh * kw * Cin
        wm = weight.detach().permute(0, 2, 3, 1).reshape(Cout, K).contiguous()      # [Cout, tap-major x channel]
        direct = kh == 1 and kw == 1 and stride == 1 and pad == 0
        if direct:
            col = x
        else:
            col = _scratch(R * K * 4, "col", x.device)[: R * K * 4].view(torch.float32)
            be.check(be.lib().lagvae_im2col(be.ptr(x), B, H, W, Cin, kh, kw, stride, pad, be.ptr(col), _st()), "lagvae_im2col")
        y = torch.empty(B, Ho, Wo, Cout, dtype=torch.float32, device=x.device)
        _gemm(col, K, 1, wm, K, 1, y, Cout, R, Cout, K)
        ctx.save_for_backward(x, wm)
        ctx.geom = (B, H, W, Cin, Cout, kh, kw, stride, pad, Ho, Wo, direct)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wm = ctx.saved_tensors
        B, H, W, Cin, Cout, kh, kw, stride, pad, Ho, Wo, direct = ctx.geom
        dy = dy.contiguous()
        R, K = B * Ho * Wo, kh * kw * Cin
        if direct:
            col = x
        else:   # recompute the patch matrix instead of keeping 24 x 315 MB of them alive
            col = _scratch(R * K * 4, "col", x.device)[: R * K * 4].view(torch.float32)
            be.check(be.lib().lagvae_im2col(be.ptr(x), B, H, W, Cin, kh, kw, stride, pad, be.ptr(col), _st()), "lagvae_im2col")
        dwm = torch.empty(Cout, K, dtype=torch.float32, device=x.device)
        _gemm(dy, 1, Cout, col, 1, K, dwm, K, Cout, K, R)                     # dW = dYᵀ · col
        dw = dwm.view(Cout, kh, kw, Cin).permute(0, 3, 1, 2).contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(B, H, W, Cin, dtype=torch.float32, device=x.device)
            if direct:
                _gemm(dy, Cout, 1, wm, 1, K, dx, Cin, R, K, Cout)             # dX = dY · W
            else:
                dcol = _scratch(R * K * 4, "dcol", x.device)[: R * K * 4].view(torch.float32)
                _gemm(dy, Cout, 1, wm, 1, K, dcol, K, R, K, Cout)
                be.check(be.lib().lagvae_col2im(be.ptr(dcol), B, H, W, Cin, kh, kw, stride, pad, be.ptr(dx), _st()), "lagvae_col2im")
        return dx, dw, None, None


def _conv32_ok(x, weight, stride, pad):
    """The tcgen05 im2col-free kernel covers the PixelCNN masked convolutions: 32 -> 32 channels, odd k <= 7, stride 1,
    pad k/2, tiles of whole image rows (csrc/conv_tc.cu)."""
    Cout, Cin, kh, kw = weight.shape
    return (os.environ.get("LAGVAE_CONV_TC", "1") != "0" and Cout == 32 and Cin == 32 and stride == 1 and kh == kw and pad == kh // 2
            and x.shape[-1] == 32 and bool(be.lib().lagvae_convtc_supported(x.shape[0], x.shape[1], x.shape[2], 32, 32, kh, kw)))


def _split_cat(x):
    """fp32 [..., C] -> bf16 [..., 2C] = [hi | lo] (the operand format of the convtc kernels), C in {32, 64}."""
    Cc = x.shape[-1]
    cat = torch.empty(*x.shape[:-1], 2 * Cc, dtype=torch.bfloat16, device=x.device)
    be.check(be.lib().lagvae_split_cat(be.ptr(x), x.numel() // Cc, Cc, be.ptr(cat), _st()), "lagvae_split_cat")
    return cat


class _Conv32Fn(torch.autograd.Function):
    """k x k (masked) convolution 32 -> 32 on NHWC activations through the im2col-free tcgen05 kernels: forward and dgrad
    multiply the live taps only; wgrad returns every tap (autograd of the reference does)."""

    @staticmethod
    def forward(ctx, x, weight, mask_mode):
        _need_cuda(x, "conv2d")
        x = x.contiguous()
        B, H, W, _ = x.shape
        k = weight.shape[2]
        w = weight.detach().contiguous()
        xcat = _split_cat(x)
        wbuf = torch.empty(be.lib().lagvae_convtc_wbuf_bytes(32, 32, k, k), dtype=torch.uint8, device=x.device)
        be.check(be.lib().lagvae_convtc_prepare_weights(be.ptr(w), 32, 32, k, k, mask_mode, be.ptr(wbuf), _st()), "lagvae_convtc_prepare_weights")
        y = torch.empty(B, H, W, 32, dtype=torch.float32, device=x.device)
        be.check(be.lib().lagvae_convtc_forward(be.ptr(xcat), be.ptr(wbuf), B, H, W, 32, 32, k, k, mask_mode, None, be.ptr(y), None, _st()),
                 "lagvae_convtc_forward")
        ctx.save_for_backward(xcat, wbuf)
        ctx.geom = (B, H, W, k, mask_mode)
        return y

    @staticmethod
    def backward(ctx, dy):
        xcat, wbuf = ctx.saved_tensors
        B, H, W, k, mask_mode = ctx.geom
        dycat = _split_cat(dy.contiguous())
        dw = torch.empty(32, 32, k, k, dtype=torch.float32, device=dy.device)
        sc = _scratch(be.lib().lagvae_convtc_wgrad_scratch_bytes(32, 32, k, k), "wgrad", dy.device)
        be.check(be.lib().lagvae_convtc_wgrad(be.ptr(dycat), be.ptr(xcat), B, H, W, 32, 32, k, k, be.ptr(dw), be.ptr(sc), _st()),
                 "lagvae_convtc_wgrad")
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(B, H, W, 32, dtype=torch.float32, device=dy.device)
            be.check(be.lib().lagvae_convtc_dgrad(be.ptr(dycat), be.ptr(wbuf), B, H, W, 32, 32, k, k, mask_mode, None, be.ptr(dx), _st()),
                     "lagvae_convtc_dgrad")
        return dx, dw, None


# ------------------------------------------------------------------------------------------------------
# fused PixelCNN block: 3 im2col-free tcgen05 convolutions with BatchNorm statistics in their epilogues + 3 fused
# BatchNorm/residual/ELU passes forward; 3 x (BN/ELU backward, wgrad, dgrad) backward (csrc/conv_tc.cu, image_fused.cu)
# ------------------------------------------------------------------------------------------------------
def _fused_ok(x, cin, cout, k, padded=False):
    if padded and cin < 32:
        cin = 32
    return (os.environ.get("LAGVAE_IMAGE_FUSED", "1") != "0" and x.is_cuda and x.dim() == 4 and (padded or x.shape[-1] == cin)
            and bool(be.lib().lagvae_convtc_supported(x.shape[0], x.shape[1], x.shape[2], cin, cout, k, k)))


def _cat_of(x):
    """The bf16 [hi | lo] operand copy of an fp32 NHWC activation: reuse the one its producer emitted, else split."""
    cat = getattr(x, "_lagvae_cat", None)
    if cat is not None and cat.shape[:-1] == x.shape[:-1] and cat.shape[-1] == 2 * x.shape[-1]:
        return cat
    return _split_cat(x.contiguous())


def _wprep(w, mask_mode):
    cout, cin, k, _ = w.shape
    wbuf = torch.empty(be.lib().lagvae_convtc_wbuf_bytes(cin, cout, k, k), dtype=torch.uint8, device=w.device)
    be.check(be.lib().lagvae_convtc_prepare_weights(be.ptr(w.detach().contiguous()), cout, cin, k, k, mask_mode, be.ptr(wbuf), _st()),
             "lagvae_convtc_prepare_weights")
    return wbuf


def _conv_fwd(xcat, wbuf, geom, cin, cout, k, mask_mode, stats):
    B, H, W = geom
    y = torch.empty(B, H, W, cout, dtype=torch.float32, device=xcat.device)
    be.check(be.lib().lagvae_convtc_forward(be.ptr(xcat), be.ptr(wbuf), B, H, W, cin, cout, k, k, mask_mode, None, be.ptr(y),
                                            be.ptr(stats) if stats is not None else None, _st()), "lagvae_convtc_forward")
    return y


def _conv_bwd(dycat, xcat, wbuf, geom, cin, cout, k, mask_mode, addend=None, need_dx=True):
    """(dx fp32 [B,H,W,cin] (+ addend), dw [cout,cin,k,k]) of a convtc convolution."""
    B, H, W = geom
    dw = torch.empty(cout, cin, k, k, dtype=torch.float32, device=dycat.device)
    sc = _scratch(be.lib().lagvae_convtc_wgrad_scratch_bytes(cin, cout, k, k), "wgrad", dycat.device)
    be.check(be.lib().lagvae_convtc_wgrad(be.ptr(dycat), be.ptr(xcat), B, H, W, cin, cout, k, k, be.ptr(dw), be.ptr(sc), _st()),
             "lagvae_convtc_wgrad")
    dx = None
    if need_dx:
        dx = torch.empty(B, H, W, cin, dtype=torch.float32, device=dycat.device)
        be.check(be.lib().lagvae_convtc_dgrad(be.ptr(dycat), be.ptr(wbuf), B, H, W, cin, cout, k, k, mask_mode, be.ptr(addend), be.ptr(dx),
                                              _st()), "lagvae_convtc_dgrad")
    return dx, dw


def _bnact_fwd(y, stats, bn, residual, want_f32, want_cat, training=True):
    """[ELU]((y - mean) * invstd * gamma + beta [+ residual]).  train(): batch statistics left by the convolution epilogue in
    `stats`, bn.running_* updated like nn.BatchNorm2d; eval(): the running statistics, nothing updated.
    Returns (out fp32 | None, cat | None, save_mean, save_invstd)."""
    Cc = y.shape[-1]
    R = y.numel() // Cc
    if not training:
        be.check(be.lib().lagvae_bn_eval_stats(be.ptr(bn.running_mean), be.ptr(bn.running_var), R, Cc, be.ptr(stats), _st()),
                 "lagvae_bn_eval_stats")
    out = torch.empty_like(y) if want_f32 else None
    cat = torch.empty(*y.shape[:-1], 2 * Cc, dtype=torch.bfloat16, device=y.device) if want_cat else None
    sm = torch.empty(Cc, dtype=torch.float32, device=y.device)
    si = torch.empty_like(sm)
    be.check(be.lib().lagvae_bnact_fwd(be.ptr(y), be.ptr(stats), R, Cc, be.ptr(bn.weight.detach()), be.ptr(bn.bias.detach()), float(bn.eps),
                                       float(bn.momentum), be.ptr(residual), 1, be.ptr(out), be.ptr(cat), be.ptr(sm), be.ptr(si),
                                       be.ptr(bn.running_mean) if training else None, be.ptr(bn.running_var) if training else None, _st()),
             "lagvae_bnact_fwd")
    return out, cat, sm, si


def _bnact_bwd(dout, out_f32, out_cat, y, gamma, sm, si, want_dres):
    """Backward of _bnact_fwd: (dycat, dres | None, dgamma, dbeta)."""
    Cc = y.shape[-1]
    R = y.numel() // Cc
    dycat = torch.empty(*y.shape[:-1], 2 * Cc, dtype=torch.bfloat16, device=y.device)
    dres = torch.empty_like(y) if want_dres else None
    dg = torch.empty(Cc, dtype=torch.float32, device=y.device)
    db = torch.empty_like(dg)
    sc = _scratch(16 * Cc + 256, "bn", y.device)
    be.check(be.lib().lagvae_bnact_bwd(be.ptr(dout), be.ptr(out_f32), be.ptr(out_cat), be.ptr(y), R, Cc, be.ptr(gamma), be.ptr(sm), be.ptr(si), 1,
                                       None, be.ptr(dycat), be.ptr(dres), be.ptr(dg), be.ptr(db), be.ptr(sc), _st()), "lagvae_bnact_bwd")
    return dycat, dres, dg, db


_BLOCK_SIZES = {}


def _block_sizes(d):
    key = (d.B, d.H, d.W, d.C, d.Cm, d.k)
    if key not in _BLOCK_SIZES:
        L = be.lib()
        _BLOCK_SIZES[key] = (int(L.lagvae_pixelblock_stash_bytes(C.byref(d))), int(L.lagvae_pixelblock_scratch_bytes(C.byref(d))))
    return _BLOCK_SIZES[key]


class _PixelBlockFn(torch.autograd.Function):
    """out = ELU(BN3(conv1x1(ELU(BN2(maskedconv_kxk(ELU(BN1(conv1x1(x))))))) + x) — PixelCNNBlock (dec_pixelcnn_v2.py:32-62):
    one C-ABI call per direction (csrc/image_plan.cu: 10 launches forward, 15 backward)."""

    @staticmethod
    def forward(ctx, x, w1, g1, b1, w2, g2, b2, w3, g3, b3, bns, holder):
        _need_cuda(x, "PixelCNNBlock")
        x = x.contiguous()
        bn1, bn2, bn3 = bns
        B, H, W, Cc = x.shape
        d = be.PixelBlockDims(B, H, W, Cc, w1.shape[0], w2.shape[2], float(bn1.eps), float(bn1.momentum), 0 if bn1.training else 1)
        prm = [t.detach() for t in (w1, g1, b1, w2, g2, b2, w3, g3, b3)]
        p = be.PixelBlockParams(*[t.data_ptr() for t in prm], bn1.running_mean.data_ptr(), bn1.running_var.data_ptr(),
                                bn2.running_mean.data_ptr(), bn2.running_var.data_ptr(), bn3.running_mean.data_ptr(),
                                bn3.running_var.data_ptr())
        stash = torch.empty(_block_sizes(d)[0], dtype=torch.uint8, device=x.device)
        xcat = getattr(x, "_lagvae_cat", None)
        if xcat is not None and tuple(xcat.shape) != (B, H, W, 2 * Cc):
            xcat = None
        out = torch.empty_like(x)
        outcat = torch.empty(B, H, W, 2 * Cc, dtype=torch.bfloat16, device=x.device)
        be.check(be.lib().lagvae_pixelblock_forward(C.byref(d), C.byref(p), be.ptr(x), be.ptr(xcat), be.ptr(stash), be.ptr(out),
                                                    be.ptr(outcat), _st()), "lagvae_pixelblock_forward")
        holder.append(outcat)
        ctx.save_for_backward(out, stash, *prm, *([xcat] if xcat is not None else []))
        ctx.d, ctx.p = d, p
        return out

    @staticmethod
    def backward(ctx, dout):
        saved = ctx.saved_tensors
        if ctx.d.eval:
            raise LagvaeError("PixelCNNBlock backward in eval() mode is not part of the hot path")
        out, stash, prm = saved[0], saved[1], saved[2:11]
        xcat = saved[11] if len(saved) > 11 else None
        d = ctx.d
        dout = dout.contiguous()
        shapes = [prm[0].shape, prm[1].shape, prm[2].shape, prm[3].shape, prm[4].shape, prm[5].shape, prm[6].shape, prm[7].shape,
                  prm[8].shape]
        sizes = [int(math.prod(sh)) for sh in shapes]
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + (n + 63) // 64 * 64)      # 256-B aligned segments
        flat = torch.empty(offs[-1], dtype=torch.float32, device=dout.device)
        views = [flat[offs[i]: offs[i] + sizes[i]].view(shapes[i]) for i in range(9)]
        g = be.PixelBlockGrads(*[v.data_ptr() for v in views])
        dx = torch.empty_like(out)
        sc = _scratch(_block_sizes(d)[1] + 256, "pixelblock", dout.device)
        scp = (sc.data_ptr() + 255) & ~255
        be.check(be.lib().lagvae_pixelblock_backward(C.byref(d), C.byref(ctx.p), be.ptr(dout), be.ptr(out), be.ptr(stash), be.ptr(xcat),
                                                     be.ptr(dx), C.byref(g), C.c_void_p(scp), _st()), "lagvae_pixelblock_backward")
        return (dx, *views, None, None)


class _ConvBnEluFn(torch.autograd.Function):
    """ELU(BN(conv_kxk(x))) on the convtc kernels: the head 1x1 64->64 of PixelCNNDecoderV2 (dec_pixelcnn_v2.py:145-147) and
    the mask-A 7x7 5->64 input block (:65-85; its 5 input channels are zero-padded to one 32-channel operand chunk — the
    mask lives in the weights, which MaskedConv2d zeroes in place before every forward)."""

    @staticmethod
    def forward(ctx, x, w, g, b, bn, holder):
        _need_cuda(x, "conv-bn-elu")
        B, H, W, C = x.shape
        Co, k = w.shape[0], w.shape[2]
        Cp = C if C in (32, 64) else 32
        wd = w.detach()
        if Cp != C:
            x = torch.nn.functional.pad(x, (0, Cp - C))
            wd = torch.nn.functional.pad(wd, (0, 0, 0, 0, 0, Cp - C))
        xcat = _cat_of(x.contiguous())
        wb = _wprep(wd, 0)
        stats = torch.empty(2 * Co, dtype=torch.float64, device=x.device)
        y = _conv_fwd(xcat, wb, (B, H, W), Cp, Co, k, 0, stats if bn.training else None)
        out, outcat, sm, si = _bnact_fwd(y, stats, bn, None, True, holder is not None, training=bn.training)
        if holder is not None:
            holder.append(outcat)
        ctx.save_for_backward(xcat, y, out, wb, sm, si, g.detach())
        ctx.geom = ((B, H, W), C, Cp, Co, k)
        ctx.training = bn.training
        return out

    @staticmethod
    def backward(ctx, dout):
        xcat, y, out, wb, sm, si, g = ctx.saved_tensors
        geom, C, Cp, Co, k = ctx.geom
        if not ctx.training:
            raise LagvaeError("BatchNorm backward in eval() mode is not part of the hot path")
        dycat, _, dg, db = _bnact_bwd(dout.contiguous(), out, None, y, g, sm, si, False)
        dx, dw = _conv_bwd(dycat, xcat, wb, geom, Cp, Co, k, 0, need_dx=ctx.needs_input_grad[0])
        if Cp != C:
            dx = dx[..., :C] if dx is not None else None
            dw = dw[:, :C].contiguous()
        return dx, dw, dg, db, None, None


class _BNFn(torch.autograd.Function):
    """nn.BatchNorm2d (affine) on NHWC rows; train(): batch statistics + running-stat update; eval(): running stats."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, eps, momentum, training):
        _need_cuda(x, "batch_norm")
        x = x.contiguous()
        Cc = x.shape[-1]
        R = x.numel() // Cc
        y = torch.empty_like(x)
        mean = torch.empty(Cc, dtype=torch.float32, device=x.device)
        invstd = torch.empty_like(mean)
        sc = _scratch(16 * Cc + 256, "bn", x.device)
        if training:
            be.check(be.lib().lagvae_bn_train_fwd(be.ptr(x), R, Cc, be.ptr(gamma.detach()), be.ptr(beta.detach()), float(eps),
                                                  float(momentum), be.ptr(y), be.ptr(mean), be.ptr(invstd), be.ptr(running_mean),
                                                  be.ptr(running_var), be.ptr(sc), _st()), "lagvae_bn_train_fwd")
        else:   # eval(): running statistics (image.py evaluates under torch.no_grad())
            mean = running_mean.detach().clone()
            invstd = torch.rsqrt(running_var.detach() + eps)
            be.check(be.lib().lagvae_bn_apply(be.ptr(x), R, Cc, be.ptr(mean), be.ptr(invstd), be.ptr(gamma.detach()),
                                              be.ptr(beta.detach()), be.ptr(y), _st()), "lagvae_bn_apply")
        ctx.training = training
        ctx.save_for_backward(x, gamma.detach(), mean, invstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        if not ctx.training:
            raise LagvaeError("BatchNorm backward in eval() mode is not part of the hot path")
        x, gamma, mean, invstd = ctx.saved_tensors
        dy = dy.contiguous()
        Cc = x.shape[-1]
        R = x.numel() // Cc
        dx = torch.empty_like(x)
        dgamma = torch.empty(Cc, dtype=torch.float32, device=x.device)
        dbeta = torch.empty_like(dgamma)
        sc = _scratch(16 * Cc + 256, "bn", x.device)
        be.check(be.lib().lagvae_bn_train_bwd(be.ptr(x), be.ptr(dy), R, Cc, be.ptr(gamma), be.ptr(mean), be.ptr(invstd), be.ptr(dx),
                                              be.ptr(dgamma), be.ptr(dbeta), be.ptr(sc), _st()), "lagvae_bn_train_bwd")
        return dx, dgamma, dbeta, None, None, None, None, None


class _EluFn(torch.autograd.Function):
    """y = ELU(a [+ b]) with alpha = 1; backward from y (shared by both addends)."""

    @staticmethod
    def forward(ctx, a, b):
        _need_cuda(a, "elu")
        a = a.contiguous()
        b = b.contiguous() if b is not None else None
        y = torch.empty_like(a)
        be.check(be.lib().lagvae_elu_fwd(be.ptr(a), be.ptr(b), be.ptr(y), a.numel(), _st()), "lagvae_elu_fwd")
        ctx.save_for_backward(y)
        ctx.two = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(y)
        be.check(be.lib().lagvae_elu_bwd(be.ptr(y), be.ptr(dy), be.ptr(dx), y.numel(), _st()), "lagvae_elu_bwd")
        return dx, (dx if ctx.two else None)


class _AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = a.contiguous(), b.contiguous()
        out = torch.empty_like(a)
        be.check(be.lib().lagvae_add(be.ptr(a), be.ptr(b), be.ptr(out), a.numel(), _st()), "lagvae_add")
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g


class _LinearFn(torch.autograd.Function):
    """nn.Linear with bias on [R, K] rows (enc_resnet_v2.py:106,123; dec_pixelcnn_v2.py:134-137)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        _need_cuda(x, "linear")
        x = x.contiguous()
        R, K = x.shape
        N = weight.shape[0]
        w = weight.detach().contiguous()
        y = torch.empty(R, N, dtype=torch.float32, device=x.device)
        _gemm(x, K, 1, w, K, 1, y, N, R, N, K, bias=bias.detach().contiguous())
        ctx.save_for_backward(x, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        R, K = x.shape
        N = w.shape[0]
        dx = torch.empty(R, K, dtype=torch.float32, device=x.device)
        _gemm(dy, N, 1, w, 1, K, dx, K, R, K, N)                               # dX = dY · W
        dw = torch.empty(N, K, dtype=torch.float32, device=x.device)
        _gemm(dy, 1, N, x, 1, K, dw, K, N, K, R)                               # dW = dYᵀ · X
        ones = torch.ones(1, R, dtype=torch.float32, device=x.device)
        db = torch.empty(N, dtype=torch.float32, device=x.device)
        _gemm(dy, 1, N, ones, R, 1, db, 1, N, 1, R)                            # db = dYᵀ · 1
        return dx, dw, db


class _BernoulliNLLFn(torch.autograd.Function):
    """sigmoid + Bernoulli NLL with eps = 1e-12 (dec_pixelcnn_v2.py:145-152,172-195): logits [B*ns, P], x [B, P]."""

    @staticmethod
    def forward(ctx, logits, x, ns):
        logits, x = logits.contiguous(), x.contiguous()
        Bd, P = logits.shape
        nll = torch.empty(Bd, dtype=torch.float32, device=logits.device)
        be.check(be.lib().lagvae_bernoulli_nll_fwd(be.ptr(logits), be.ptr(x), Bd // ns, ns, P, be.ptr(nll), _st()), "lagvae_bernoulli_nll_fwd")
        ctx.save_for_backward(logits, x)
        ctx.ns = ns
        return nll

    @staticmethod
    def backward(ctx, g):
        logits, x = ctx.saved_tensors
        Bd, P = logits.shape
        g = g.contiguous()
        dl = torch.empty_like(logits)
        be.check(be.lib().lagvae_bernoulli_nll_bwd(be.ptr(logits), be.ptr(x), be.ptr(g), Bd // ctx.ns, ctx.ns, P, be.ptr(dl), _st()),
                 "lagvae_bernoulli_nll_bwd")
        return dl, None, None


class _ReparamKLFn(torch.autograd.Function):
    """(z, KL) from (mu, logvar, eps) — encoder.py:40-79 (the image encoders produce mu/logvar with their own head)."""

    @staticmethod
    def forward(ctx, mu, logvar, eps):
        _need_cuda(mu, "reparameterize")
        mu, logvar, eps = mu.contiguous(), logvar.contiguous(), eps.contiguous()
        B, nz = mu.shape
        ns = eps.shape[1]
        z = torch.empty(B, ns, nz, dtype=torch.float32, device=mu.device)
        kl = torch.empty(B, dtype=torch.float32, device=mu.device)
        be.check(be.lib().lagvae_reparam_kl_fwd(be.ptr(mu), be.ptr(logvar), be.ptr(eps), B, nz, ns, be.ptr(z), be.ptr(kl), _st()),
                 "lagvae_reparam_kl_fwd")
        ctx.save_for_backward(mu, logvar, eps)
        return z, kl

    @staticmethod
    def backward(ctx, dz, dkl):
        mu, logvar, eps = ctx.saved_tensors
        B, nz = mu.shape
        ns = eps.shape[1]
        dml = torch.empty(B, 2 * nz, dtype=torch.float32, device=mu.device)
        dz = dz.contiguous() if dz is not None else None
        dkl = dkl.contiguous() if dkl is not None else None
        be.check(be.lib().lagvae_reparam_kl_bwd(be.ptr(dz), be.ptr(eps), be.ptr(mu), be.ptr(logvar), be.ptr(dkl), B, nz, ns, be.ptr(dml),
                                                _st()), "lagvae_reparam_kl_bwd")
        return dml[:, :nz], dml[:, nz:], None


# ------------------------------------------------------------------------------------------------------
# layers (parameter containers are the stock torch classes -> reference-compatible state_dict)
# ------------------------------------------------------------------------------------------------------
class Conv2dK(nn.Conv2d):
    def forward(self, x):   # x: NHWC
        return _ConvFn.apply(x, self.weight, self.stride[0], self.padding[0])


class MaskedConv2d(nn.Conv2d):
    """Reference dec_pixelcnn_v2.py:12-30 incl. the in-place `weight.data.mul_(mask)` on every forward (so masked taps
    of the checkpoint are zero while their gradients are not — SURVEY §7 quirk 6d)."""

    def __init__(self, mask_type, masked_channels, *args, **kwargs):
        super().__init__(*args, **kwargs)
        assert mask_type in {"A", "B"}
        self.mask_type, self.masked_channels = mask_type, masked_channels
        self.register_buffer("mask", self.weight.data.clone())
        _, _, kH, kW = self.weight.size()
        self.mask.fill_(1)
        self.mask[:, :masked_channels, kH // 2, kW // 2 + (mask_type == "B"):] = 0
        self.mask[:, :masked_channels, kH // 2 + 1:] = 0

    def reset_parameters(self):
        n = self.kernel_size[0] * self.kernel_size[1] * self.out_channels
        self.weight.data.normal_(0, math.sqrt(2.0 / n))
        if self.bias is not None:
            self.bias.data.zero_()

    def forward(self, x):
        self.weight.data.mul_(self.mask)
        if self.masked_channels == self.in_channels and _conv32_ok(x, self.weight, self.stride[0], self.padding[0]):
            return _Conv32Fn.apply(x, self.weight, 1 if self.mask_type == "A" else 2)
        return _ConvFn.apply(x, self.weight, self.stride[0], self.padding[0])


class BatchNorm2dK(nn.BatchNorm2d):
    def forward(self, x):   # x: NHWC
        if self.training and self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(1)
        return _BNFn.apply(x, self.weight, self.bias, self.running_mean, self.running_var, self.eps, self.momentum, self.training)


class ELUK(nn.ELU):
    def forward(self, x):
        return _EluFn.apply(x, None)


def _init_convs_and_bn(module):
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
            m.weight.data.normal_(0, math.sqrt(2.0 / n))
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.data.fill_(1)
            m.bias.data.zero_()


# ------------------------------------------------------------------------------------------------------
# encoder — enc_resnet_v2.py
# ------------------------------------------------------------------------------------------------------
class ResNetBlock(nn.Module):
    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        self.conv1 = Conv2dK(inplanes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn1 = BatchNorm2dK(planes)
        self.activation = ELUK()
        self.conv2 = Conv2dK(planes, planes, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = BatchNorm2dK(planes)
        downsample = None
        if stride != 1 or inplanes != planes:
            downsample = nn.Sequential(Conv2dK(inplanes, planes, kernel_size=1, stride=stride, bias=False), BatchNorm2dK(planes))
        self.downsample = downsample
        self.stride = stride
        _init_convs_and_bn(self)

    def forward(self, x):
        residual = x if self.downsample is None else self.downsample(x)
        out = self.activation(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return _EluFn.apply(out, residual)                       # ELU(out + residual)  enc_resnet_v2.py:69


class ResNet(nn.Module):
    def __init__(self, inplanes, planes, strides):
        super().__init__()
        assert len(planes) == len(strides)
        blocks = []
        for plane, stride in zip(planes, strides):
            blocks.append(ResNetBlock(inplanes, plane, stride=stride))
            inplanes = plane
        self.main = nn.Sequential(*blocks)

    def forward(self, x):
        return self.main(x)


class ResNetEncoderV2(GaussianEncoderBase):
    """Reference enc_resnet_v2.py:93-126; input x [B,1,28,28] (NCHW == NHWC for one channel)."""

    def __init__(self, args, ngpu=1):
        super().__init__()
        self.ngpu = ngpu
        self.nz = args.nz
        self.nc = 1
        hidden_units = 512
        self.main = nn.Sequential(
            ResNet(self.nc, [64, 64, 64], [2, 2, 2]),
            Conv2dK(64, hidden_units, 4, 1, 0, bias=False),
            BatchNorm2dK(hidden_units),
            ELUK(),
        )
        self.linear = nn.Linear(hidden_units, 2 * self.nz)
        self.reset_parameters()

    def reset_parameters(self):
        _init_convs_and_bn(self.main)
        nn.init.xavier_uniform_(self.linear.weight)
        nn.init.constant_(self.linear.bias, 0.0)

    def forward(self, input):
        _need_cuda(input, "ResNetEncoderV2")
        B = input.shape[0]
        x = input.reshape(B, 28, 28, self.nc) if self.nc == 1 else input.permute(0, 2, 3, 1).contiguous()
        out = self.main(x.float())                                # [B,1,1,512]
        out = _LinearFn.apply(out.reshape(B, -1), self.linear.weight, self.linear.bias)
        return out.chunk(2, 1)

    def encode(self, input, nsamples):
        mu, logvar = self.forward(input)
        eps = torch.empty(mu.shape[0], nsamples, mu.shape[1], dtype=torch.float32, device=mu.device).normal_()  # encoder.py:77
        return _ReparamKLFn.apply(mu, logvar, eps)

    def _engine_for(self, x):
        from .text import get_engine
        return get_engine(2, 1, 8, self.nz, x.device)               # only the MI kernel is used (dims are irrelevant to it)


# ------------------------------------------------------------------------------------------------------
# decoder — dec_pixelcnn_v2.py
# ------------------------------------------------------------------------------------------------------
class PixelCNNBlock(nn.Module):
    def __init__(self, in_channels, kernel_size):
        super().__init__()
        self.mask_type = "B"
        padding = kernel_size // 2
        out_channels = in_channels // 2
        self.main = nn.Sequential(
            Conv2dK(in_channels, out_channels, 1, bias=False),
            BatchNorm2dK(out_channels),
            ELUK(),
            MaskedConv2d(self.mask_type, out_channels, out_channels, out_channels, kernel_size, padding=padding, bias=False),
            BatchNorm2dK(out_channels),
            ELUK(),
            Conv2dK(out_channels, in_channels, 1, bias=False),
            BatchNorm2dK(in_channels),
        )
        self.activation = ELUK()
        _init_convs_and_bn(self)

    def forward(self, input):
        m = self.main
        k = m[3].kernel_size[0]
        if (self.training or not torch.is_grad_enabled()) and m[0].out_channels == 32 and _fused_ok(input, 64, 32, 1) and k <= 7:
            m[3].weight.data.mul_(m[3].mask)                        # dec_pixelcnn_v2.py:29
            if self.training:
                for bn in (m[1], m[4], m[7]):
                    bn.num_batches_tracked.add_(1)
            holder = []
            out = _PixelBlockFn.apply(input, m[0].weight, m[1].weight, m[1].bias, m[3].weight, m[4].weight, m[4].bias,
                                      m[6].weight, m[7].weight, m[7].bias, (m[1], m[4], m[7]), holder)
            out._lagvae_cat = holder[0]                             # operand copy for the next block's first convolution
            return out
        return _EluFn.apply(self.main(input), input)                # ELU(main(x) + x)  dec_pixelcnn_v2.py:61


class MaskABlock(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, masked_channels):
        super().__init__()
        self.mask_type = "A"
        padding = kernel_size // 2
        self.main = nn.Sequential(
            MaskedConv2d(self.mask_type, masked_channels, in_channels, out_channels, kernel_size, padding=padding, bias=False),
            BatchNorm2dK(out_channels),
            ELUK(),
        )
        m = self.main[1]
        m.weight.data.fill_(1)
        m.bias.data.zero_()

    def forward(self, input):
        conv, bn = self.main[0], self.main[1]
        if (self.training or not torch.is_grad_enabled()) and conv.out_channels == 64 and conv.in_channels <= 32 \
                and conv.kernel_size[0] <= 7 and _fused_ok(input, input.shape[-1], 64, conv.kernel_size[0], padded=True):
            conv.weight.data.mul_(conv.mask)                        # dec_pixelcnn_v2.py:29
            if self.training:
                bn.num_batches_tracked.add_(1)
            holder = []
            out = _ConvBnEluFn.apply(input, conv.weight, bn.weight, bn.bias, bn, holder)
            out._lagvae_cat = holder[0]
            return out
        return self.main(input)


class PixelCNN(nn.Module):
    def __init__(self, in_channels, out_channels, num_blocks, kernel_sizes, masked_channels):
        super().__init__()
        assert num_blocks == len(kernel_sizes)
        blocks = []
        for i in range(num_blocks):
            blocks.append(MaskABlock(in_channels, out_channels, kernel_sizes[i], masked_channels) if i == 0
                          else PixelCNNBlock(out_channels, kernel_sizes[i]))
        self.main = nn.ModuleList(blocks)
        self.direct_connects = nn.ModuleList([PixelCNNBlock(out_channels, kernel_sizes[i]) for i in range(1, num_blocks - 1)])

    def forward(self, input):
        direct_inputs = []
        for i, layer in enumerate(self.main):
            if i > 2:                                                # dec_pixelcnn_v2.py:112-115
                input = _AddFn.apply(input, self.direct_connects[i - 3](direct_inputs.pop(0)))
            input = layer(input)
            direct_inputs.append(input)
        assert len(direct_inputs) == 3, "architecture error: %d" % len(direct_inputs)
        return _AddFn.apply(input, self.direct_connects[-1](direct_inputs.pop(0)))


class PixelCNNDecoderV2(DecoderBase):
    """Reference dec_pixelcnn_v2.py:123-232 (training objective on the fused tcgen05 path; `decode` = ancestral sampling)."""

    def __init__(self, args, ngpu=1, mode="large"):
        super().__init__()
        self.ngpu = ngpu
        self.nz = args.nz
        self.nc = 1
        self.fm_latent = args.latent_feature_map
        self.img_latent = 28 * 28 * self.fm_latent
        if self.nz != 0:
            self.z_transform = nn.Sequential(nn.Linear(self.nz, self.img_latent))
        if mode == "small":
            kernal_sizes = [7, 7, 7, 5, 5, 3, 3]
        elif mode == "large":
            kernal_sizes = [7, 7, 7, 7, 7, 5, 5, 5, 5, 3, 3, 3, 3]
        else:
            raise ValueError("unknown mode: %s" % mode)
        hidden_channels = 64
        self.main = nn.Sequential(
            PixelCNN(self.nc + self.fm_latent, hidden_channels, len(kernal_sizes), kernal_sizes, self.nc),
            Conv2dK(hidden_channels, hidden_channels, 1, bias=False),
            BatchNorm2dK(hidden_channels),
            ELUK(),
            Conv2dK(hidden_channels, self.nc, 1, bias=False),
            nn.Sigmoid(),                                             # never called: fused into the Bernoulli NLL kernel
        )
        self.reset_parameters()

    def reset_parameters(self):
        if self.nz != 0:
            nn.init.xavier_uniform_(self.z_transform[0].weight)
            nn.init.constant_(self.z_transform[0].bias, 0)
        m = self.main[2]
        m.weight.data.fill_(1)
        m.bias.data.zero_()

    def _logits(self, img_nhwc):
        h = self.main[0](img_nhwc)
        conv, bn = self.main[1], self.main[2]
        if (self.training or not torch.is_grad_enabled()) and conv.out_channels == 64 and _fused_ok(h, 64, 64, 1):
            if self.training:
                bn.num_batches_tracked.add_(1)
            h = _ConvBnEluFn.apply(h, conv.weight, bn.weight, bn.bias, bn, None)
        else:
            h = self.main[3](bn(conv(h)))
        return self.main[4](h)                                        # [N,28,28,1]; the Sigmoid is fused into the NLL kernel

    def forward(self, input):
        """Probabilities for an NCHW input [N, nc+fm, 28, 28] (API compatibility; the loss path uses logits)."""
        _need_cuda(input, "PixelCNNDecoderV2")
        p = torch.sigmoid(self._logits(input.permute(0, 2, 3, 1).contiguous()))
        return p.permute(0, 3, 1, 2)

    def reconstruct_error(self, x, z):
        """Bernoulli NLL summed over the 784 pixels, [B, ns] (dec_pixelcnn_v2.py:172-195)."""
        _need_cuda(x, "PixelCNNDecoderV2")
        B, ns, _ = z.size()
        zl = self.z_transform[0]
        zt = _LinearFn.apply(z.reshape(B * ns, self.nz), zl.weight, zl.bias)          # [B*ns, fm*784] laid out (fm, H, W)
        zt = zt.view(B * ns, self.fm_latent, 28, 28).permute(0, 2, 3, 1)              # -> NHWC view
        xi = x.reshape(B, 1, 28, 28, self.nc).expand(B, ns, 28, 28, self.nc).reshape(B * ns, 28, 28, self.nc)
        img = torch.cat([xi.float(), zt], dim=3).contiguous()                        # channel concat (:187), data movement only
        logits = self._logits(img).reshape(B * ns, 28 * 28 * self.nc)
        nll = _BernoulliNLLFn.apply(logits, x.reshape(B, -1).float(), ns)
        return nll.view(B, ns)

    def log_probability(self, x, z):
        return -self.reconstruct_error(x, z)

    def decode(self, z, deterministic):
        """Ancestral sampling (dec_pixelcnn_v2.py:201-232): 784 sequential decoder forwards, pixel (i, j) set
        from its predicted probability (thresholded at 0.5 or Bernoulli-sampled); returns (image [B,nc,28,28], final
        probabilities).  Host-driven like the reference; every forward runs on the liblagvae.so kernels."""
        _need_cuda(z, "PixelCNNDecoderV2")
        H = W = 28
        B = z.size(0)
        zl = self.z_transform[0]
        with torch.no_grad():
            zt = _LinearFn.apply(z.detach().float().contiguous(), zl.weight, zl.bias).view(B, self.fm_latent, H, W)
            img = torch.cat([torch.zeros(B, self.nc, H, W, dtype=torch.float32, device=z.device), zt], dim=1)
            for i in range(H):
                for j in range(W):
                    probs = self.forward(img)
                    pij = probs[:, :, i, j]
                    img[:, :self.nc, i, j] = torch.ge(pij, 0.5).float() if deterministic else torch.bernoulli(pij)
            return img[:, :self.nc], self.forward(img)
