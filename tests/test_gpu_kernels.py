"""GPU unit tests of the building-block kernels, called through the C-ABI (ctypes)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _be():
    import lagvae._backend as be
    return be


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _split(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi, lo


@pytest.mark.parametrize("M,N,K", [(64, 64, 16), (37, 91, 53), (130, 257, 129), (5, 200, 50), (300, 7, 1000)])
@pytest.mark.parametrize("at,bt", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_f32(M, N, K, at, bt):
    be = _be()
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    A = torch.randn((K, M) if at else (M, K), generator=g, device="cuda")
    B = torch.randn((K, N) if bt else (N, K), generator=g, device="cuda")
    C0 = torch.randn(M, N + 3, generator=g, device="cuda")
    Cc = C0.clone()
    bias = torch.randn(N, generator=g, device="cuda")
    rows = torch.randn(4, N, generator=g, device="cuda")
    a_rs, a_cs = (1, M) if at else (K, 1)
    b_rs, b_cs = (1, N) if bt else (K, 1)
    be.check(be.lib().lagvae_gemm_f32(be.ptr(A), a_rs, a_cs, be.ptr(B), b_rs, b_cs, be.ptr(Cc), N + 3, M, N, K,
                                      0.5, 2.0, be.ptr(bias), be.ptr(rows), 4, _st()))
    Ad = (A.t() if at else A).double()
    Bd = (B.t() if bt else B).double()
    ref = 0.5 * Ad @ Bd.t() + 2.0 * C0[:, :N].double() + bias.double() + rows.double()[torch.arange(M, device="cuda") % 4]
    err = float((Cc[:, :N].double() - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, err
    assert torch.equal(Cc[:, N:], C0[:, N:])


def test_split_bf16():
    be = _be()
    x = torch.randn(33, 50, device="cuda") * 3
    hi = torch.zeros(33, 56, dtype=torch.bfloat16, device="cuda")
    lo = torch.zeros_like(hi)
    be.check(be.lib().lagvae_split_bf16(be.ptr(x), 50, 33, 50, be.ptr(hi), be.ptr(lo), 56, _st()))
    rec = hi.float() + lo.float()
    assert float((rec[:, :50] - x).abs().max() / x.abs().max()) < 2.0 ** -15
    assert float(rec[:, 50:].abs().max()) == 0.0
    h2, l2 = _split(x)
    assert torch.equal(hi[:, :50], h2) and torch.equal(lo[:, :50], l2)


TC_SHAPES = [(128, 128, 64), (256, 384, 192), (200, 130, 100), (129, 257, 72), (1000, 96, 520), (64, 520, 4096),
             (4096, 64, 160),
             # >= 2 x 148 tiles of 128x256: the wide-tile (BN=256, 2-stage) variant; ragged M, N and K
             (2048, 4864, 128), (2500, 4000, 200), (6368, 1536 + 40, 72)]


@pytest.mark.parametrize("M,N,K", TC_SHAPES)
@pytest.mark.parametrize("amn,bmn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("passes", [3, 1])
def test_gemm_tc(M, N, K, amn, bmn, passes):
    be = _be()
    g = torch.Generator(device="cuda").manual_seed(M + 3 * N + 5 * K)
    pad8 = lambda v: (v + 7) // 8 * 8
    A = torch.randn(M, K, generator=g, device="cuda")
    B = torch.randn(N, K, generator=g, device="cuda")
    Ast = torch.zeros((K, pad8(M)) if amn else (M, pad8(K)), device="cuda")
    Bst = torch.zeros((K, pad8(N)) if bmn else (N, pad8(K)), device="cuda")
    if amn:
        Ast[:, :M] = A.t()
    else:
        Ast[:, :K] = A
    if bmn:
        Bst[:, :N] = B.t()
    else:
        Bst[:, :K] = B
    ah, al = _split(Ast)
    bh, bl = _split(Bst)
    Cc = torch.full((M, N + 1), 7.0, device="cuda")
    bias = torch.randn(N, generator=g, device="cuda")
    be.check(be.lib().lagvae_gemm_tc(be.ptr(ah), be.ptr(al), Ast.shape[1], amn, be.ptr(bh), be.ptr(bl), Bst.shape[1],
                                     bmn, be.ptr(Cc), N + 1, M, N, K, passes, 1.0, 0.0, be.ptr(bias), None, 0, None,
                                     _st()))
    torch.cuda.synchronize()
    if passes == 3:
        ref = A.double() @ B.double().t() + bias.double()
        tol = 3e-5
    else:
        ref = A.to(torch.bfloat16).double() @ B.to(torch.bfloat16).double().t() + bias.double()
        tol = 1e-5
    scale = float((A.double().abs() @ B.double().abs().t()).max())
    err = float((Cc[:, :N].double() - ref).abs().max()) / scale
    assert err < tol, "rel err %.3e" % err
    assert float((Cc[:, N] - 7.0).abs().max()) == 0.0


def test_gemm_tc_beta_rows_and_map():
    be = _be()
    M, N, K = 260, 140, 200
    A, B = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
    ah, al = _split(A)
    bh, bl = _split(B)
    C0 = torch.randn(M, N, device="cuda")
    Cc = C0.clone()
    rows = torch.randn(13, N, device="cuda")
    perm = torch.randperm(M, device="cuda").to(torch.int32)
    be.check(be.lib().lagvae_gemm_tc(be.ptr(ah), be.ptr(al), K, 0, be.ptr(bh), be.ptr(bl), K, 0, be.ptr(Cc), N, M, N, K, 3,
                                     0.25, 1.0, None, be.ptr(rows), 13, be.ptr(perm), _st()))
    prod = 0.25 * A.double() @ B.double().t() + rows.double()[torch.arange(M, device="cuda") % 13]
    ref = C0.double().clone()
    ref[perm.long()] += prod
    assert float((Cc.double() - ref).abs().max() / ref.abs().max()) < 3e-5


def test_dropout_mask_statistics_and_determinism():
    be = _be()
    n = 1 << 20
    m1 = torch.empty(n, dtype=torch.uint8, device="cuda")
    m2 = torch.empty_like(m1)
    m3 = torch.empty_like(m1)
    be.check(be.lib().lagvae_dropout_mask(1234, 1, n, 0.5, be.ptr(m1), _st()))
    be.check(be.lib().lagvae_dropout_mask(1234, 1, n, 0.5, be.ptr(m2), _st()))
    be.check(be.lib().lagvae_dropout_mask(1234, 2, n, 0.5, be.ptr(m3), _st()))
    assert torch.equal(m1, m2)
    assert abs(float(m1.float().mean()) - 0.5) < 3e-3
    assert abs(float((m1 == m3).float().mean()) - 0.5) < 3e-3


def test_mi_estimate_matches_oracle():
    import lagging_oracle as O
    be = _be()
    for B, nz in [(32, 32), (5, 3), (64, 1), (100, 8)]:
        mu, lv, eps = torch.randn(B, nz), 0.3 * torch.randn(B, nz), torch.randn(B, 1, nz)
        want = O.calc_mi_from_stats(mu.double(), lv.double(), eps.double())
        out = torch.empty(1, device="cuda")
        mu_d, lv_d, eps_d = mu.cuda(), lv.cuda(), eps.cuda()      # keep the device tensors alive across the launch
        be.check(be.lib().lagvae_mi_estimate(be.ptr(mu_d), be.ptr(lv_d), be.ptr(eps_d), B, nz, be.ptr(out), _st()))
        assert abs(float(out) - want) <= 1e-4 * max(1.0, abs(want)), (float(out), want)


def test_clip_sgd_matches_torch():
    be = _be()
    for scale in (0.1, 30.0):      # below / above the clip threshold
        ps = [torch.randn(n, device="cuda") for n in (1000, 77, 5000, 3)]
        gs = [scale * torch.randn_like(p) for p in ps]
        ref_p = [p.clone().requires_grad_(True) for p in ps]
        for rp, g in zip(ref_p, gs):
            rp.grad = g.clone()
        norm_ref = torch.nn.utils.clip_grad_norm_(ref_p, 5.0)
        torch.optim.SGD(ref_p[:2], lr=1.0).step()
        n = len(ps)
        P = (C.c_void_p * n)(*[p.data_ptr() for p in ps])
        G = (C.c_void_p * n)(*[g.data_ptr() for g in gs])
        cnt = (C.c_int64 * n)(*[g.numel() for g in gs])
        norm = torch.empty(1, device="cuda")
        scratch = torch.empty(16384, dtype=torch.uint8, device="cuda")
        be.check(be.lib().lagvae_clip_sgd_step(P, G, cnt, n, 2, 5.0, 1.0, 1, be.ptr(norm), be.ptr(scratch), _st()))
        assert abs(float(norm) - float(norm_ref)) < 1e-5 * float(norm_ref)
        for p, rp, g in zip(ps, ref_p, gs):
            assert float((p - rp.detach()).abs().max()) < 1e-6 * max(1.0, float(rp.abs().max()))
            assert float((g - rp.grad).abs().max()) < 1e-6 * max(1.0, float(rp.grad.abs().max()))


@pytest.mark.parametrize("nh,Bd,Tn,init,dropout", [(128, 16, 5, True, True), (1024, 32, 6, False, False),
                                                    (1024, 32, 4, True, True), (192, 70, 4, True, False),
                                                    (64, 5, 3, False, True), (512, 130, 3, True, True),
                                                    (1024, 64, 3, True, True), (1024, 100, 3, True, False),
                                                    (256, 8, 4, False, True), (512, 32, 5, True, True)])
def test_lstm_persistent_tcgen05_matches_step_tier(nh, Bd, Tn, init, dropout):
    """Persistent tcgen05 recurrence (one cooperative launch) vs the launch-per-step fp32 tier."""
    be = _be()
    g = torch.Generator(device="cuda").manual_seed(nh + Bd)
    rn = lambda *s: torch.randn(*s, generator=g, device="cuda")
    w_hh = (torch.rand(4 * nh, nh, generator=g, device="cuda") * 2 - 1) * (3.0 / nh ** 0.5)
    pre = rn(Tn * Bd, 4 * nh)
    h0 = torch.tanh(rn(Bd, nh)) if init else None
    c0 = rn(Bd, nh) if init else None
    mask = (torch.rand(Bd, Tn, nh, generator=g, device="cuda") > 0.5).to(torch.uint8).contiguous()
    drop = be.Dropout()
    drop.mode, drop.p_in, drop.p_out, drop.mask_out = (1 if dropout else 0), 0.0, 0.5, mask.data_ptr()
    ws = torch.zeros(int(be.lib().lagvae_lstm_workspace_bytes(nh, Bd)), dtype=torch.uint8, device="cuda")
    dh_ext = rn(Tn * Bd, nh) * 0.1
    dh_last = rn(Bd, nh) * 0.1
    res = []
    for tier in (0, 1):
        gates = pre.clone()
        c_all, h_all, hd = torch.zeros(Tn * Bd, nh, device="cuda"), torch.zeros(Tn * Bd, nh, device="cuda"), torch.zeros(Tn * Bd, nh, device="cuda")
        be.check(be.lib().lagvae_lstm_forward(tier, nh, Tn, Bd, be.ptr(w_hh), be.ptr(h0), be.ptr(c0), be.ptr(gates), be.ptr(c_all),
                                              be.ptr(h_all), be.ptr(hd) if dropout else None, C.byref(drop), be.ptr(ws), ws.numel(), _st()),
                 "lstm_forward tier %d" % tier)
        torch.cuda.synchronize()
        dc, dhr, dg = torch.zeros(Bd, nh, device="cuda"), torch.zeros(Bd, nh, device="cuda"), torch.zeros(Tn * Bd, 4 * nh, device="cuda")
        be.check(be.lib().lagvae_lstm_backward(tier, nh, Tn, Bd, be.ptr(w_hh), be.ptr(c0), be.ptr(gates), be.ptr(c_all), be.ptr(dh_ext),
                                               be.ptr(dh_last), C.byref(drop), be.ptr(dc), be.ptr(dhr), be.ptr(dg), 1, be.ptr(ws),
                                               ws.numel(), _st()), "lstm_backward tier %d" % tier)
        torch.cuda.synchronize()
        res.append((gates, c_all, h_all, hd, dg, dc, dhr))
    names = ["gates", "c", "h", "hdrop", "dgates", "dc_init", "dh_init"]
    for n, a, b in zip(names, res[1], res[0]):
        err = float((a - b).abs().max() / (b.abs().max() + 1e-20))
        assert err < 2e-4, "%s: rel err %.3e" % (n, err)
    assert float(res[0][2].abs().max()) > 0.05   # the comparison is not vacuous


@pytest.mark.parametrize("scale", [0.05, 40.0], ids=["below-clip", "above-clip"])
def test_clip_adam_matches_torch(scale):
    """lagvae_clip_adam_step (image.py:312-314: clip_grad_norm_ over ALL parameters, Adam on the encoder's) against
    torch.nn.utils.clip_grad_norm_ + torch.optim.Adam over several steps; more tensors than fit a kernel-argument table."""
    import lagvae
    g = torch.Generator(device="cuda").manual_seed(5)
    shapes = [(64, 3, 3, 3), (64,), (64,), (5000,), (33, 7), (1,)] * 6 + [(20000,), (17,)]
    n_update = 20
    ps = [torch.randn(*s, generator=g, device="cuda") for s in shapes]
    ref_p = [p.clone().requires_grad_(True) for p in ps]
    opt = torch.optim.Adam(ref_p[:n_update], lr=1e-3)
    grads = [torch.zeros_like(p) for p in ps]
    ca = lagvae.ClipAdam(ps, grads, n_update, lr=1e-3, max_norm=5.0)
    for it in range(4):
        new = [scale * torch.randn(*s, generator=g, device="cuda") for s in shapes]
        for gr, nw, rp in zip(grads, new, ref_p):
            gr.copy_(nw)
            rp.grad = nw.clone()
        norm_ref = torch.nn.utils.clip_grad_norm_(ref_p, 5.0)
        opt.step()
        norm = ca.step()
        assert abs(float(norm) - float(norm_ref)) <= 1e-5 * float(norm_ref)
        for p, rp, gr in zip(ps, ref_p, grads):
            assert float((gr - rp.grad).abs().max()) <= 1e-6 * max(1.0, float(rp.grad.abs().max()))
            # Adam's update is lr * m / (sqrt(v) + eps): a relative rounding difference in sqrt/div moves p by ~lr * 1e-6
            assert float((p - rp.detach()).abs().max()) <= 2e-6 * max(1.0, float(rp.detach().abs().max())), it
    for p, rp in zip(ps[n_update:], ref_p[n_update:]):
        assert torch.equal(p, rp.detach())            # tensors beyond n_update are never stepped
