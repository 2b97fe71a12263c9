"""Parity of the configuration bench.py times (BASELINE.json configs[1]: Yahoo shape B=32, T=200, V=20001, ni=512,
nh=1024, nz=32, train() mode with dropout 0.5/0.5, kl_weight 0.1) against the unmodified reference / the oracle.

* mode 1: the reference's own dropout draws (fixture `yahoo_train.npz`, bit-packed keep-masks) are fed to the kernels:
  loss / rec / KL 1e-4 relative, every gradient's norm 2e-3 (fp32 accumulation-order noise over up to 6368-term sums on
  split-bf16 operands) and strided samples, the clip norm 1e-3, the post-step encoder parameters.
* mode 2 (in-kernel Philox = what bench.py runs): `lagvae_dropout_mask` materialises the masks the kernels used and the
  oracle is run on exactly those (CPU, a few seconds on the GPU box's host cores).
* the persistent cluster recurrence `k_lstm_v2` (nh=1024; Bd=32 and 128; T=200 / 64; dropout on) against float64
  autograd of `oracle.lstm_sequence` — not against the repo's own launch-per-step tier.
"""
import ctypes as C

import numpy as np
import pytest
import torch

import lagging_oracle as O
from util import assert_close

pytestmark = pytest.mark.gpu
OUT_TOL = 1e-4      # north_star: ELBO / KL / NLL / MI within 1e-4 relative
NORM_TOL = 2e-3     # per-tensor gradient norm
SLICE_TOL = 5e-3    # strided gradient samples, relative to max(|sample|, 5% of the tensor max)


def _yahoo(golden):
    g = golden("yahoo_train")
    V, ni, nh, nz, B, T, ns, train = [int(v) for v in g["meta"]]
    assert train == 1 and (V, ni, nh, nz, B, T) == (20001, 512, 1024, 32, 32, 200)
    p = O.scale_trained_like(O.init_text_params(V, ni, nh, nz, seed=0), 4.0)
    x = O.make_token_batch(B, T, V)
    m_in = np.unpackbits(g["mask_in_bits"])[: B * (T - 1) * ni].reshape(B, T - 1, ni)
    m_out = np.unpackbits(g["mask_out_bits"])[: B * (T - 1) * nh].reshape(B, T - 1, nh)
    return g, (V, ni, nh, nz, B, T), p, x, torch.from_numpy(m_in), torch.from_numpy(m_out)


def _check_grads(g, grads):
    tot, bad = 0.0, []
    for k, gr in zip(O.ALL_KEYS, grads):
        n = float(gr.double().norm())
        tot += n * n
        want = float(g["gnorm." + k])
        sl = gr.reshape(-1)[:: max(1, gr.numel() // 64)][:64].cpu().double()
        ws = torch.from_numpy(g["gslice." + k]).double()
        serr = float((sl - ws).abs().max()) / max(float(gr.abs().max()) * 0.05, float(ws.abs().max()), 1e-12)
        if abs(n - want) > NORM_TOL * max(want, 1e-6) or serr > SLICE_TOL:
            bad.append("%s: norm %.6g want %.6g, slice err %.2e" % (k, n, want, serr))
    assert not bad, "\n".join(bad)
    return tot ** 0.5


def test_yahoo_train_mode_reference_masks(golden):
    """mode 1 at the benchmarked shape: forward, all 13 gradients, clip, encoder SGD step vs the unmodified reference."""
    import lagvae
    g, (V, ni, nh, nz, B, T), p, x, m_in, m_out = _yahoo(golden)
    eng = lagvae.TextEngine(V, ni, nh, nz, "cuda")
    params = [p[k].cuda().contiguous() for k in O.ALL_KEYS]
    drop = lagvae.DropoutSpec(1, 0.5, 0.5, m_in.cuda().contiguous(), m_out.cuda().contiguous(), 0)
    xc, eps = x.cuda(), torch.from_numpy(g["eps"]).cuda()
    loss, rec, kl, mu, lv, z = eng.loss_forward(params, xc, eps, float(g["kl_weight"]), drop, want_stats=True)
    assert_close(loss, g["loss"], OUT_TOL, "loss")
    assert_close(rec, g["rec"], OUT_TOL, "rec")
    assert_close(kl, g["kl"], OUT_TOL, "kl", floor=1e-2)
    assert_close(mu, g["mu"], OUT_TOL, "mu", floor=1e-2)
    assert_close(lv, g["logvar"], OUT_TOL, "logvar", floor=1e-2)
    assert lagvae.lstm_variant()["forward"].startswith("v2"), lagvae.lstm_variant()     # the cluster kernel ran, not a fallback
    grads = eng.loss_backward(params, xc, torch.full((B,), 1.0 / B, device="cuda"), None, None)
    assert lagvae.lstm_variant()["backward"].startswith("v2"), lagvae.lstm_variant()
    tot = _check_grads(g, grads)
    assert abs(tot - float(g["grad_norm"])) <= 1e-3 * float(g["grad_norm"])
    assert float(g["grad_norm"]) > 5.0            # the clip is active in this fixture
    norm = eng.clip_sgd(params, grads, 6, 5.0, 1.0)                                      # text.py:385,387
    assert abs(float(norm) - float(g["grad_norm"])) <= 1e-3 * float(g["grad_norm"])
    coef = 5.0 / (float(g["grad_norm"]) + 1e-6)                                          # clip_grad_norm_ (A.5)
    for i, k in enumerate(O.ENC_KEYS):
        q = params[i]
        assert abs(float(q.double().norm()) - float(g["postnorm." + k])) <= 1e-5 * float(g["postnorm." + k]), k
        sl = q.reshape(-1)[:: max(1, q.numel() // 64)][:64].cpu()
        # post = p - coef * grad: the error budget is the gradient budget scaled by the clip coefficient (the embedding
        # gradient is sparse, so a per-tensor RMS would be the wrong yardstick) plus fp32 rounding of the parameter
        gscale = max(float(grads[i].abs().max()) * 0.05 / coef, float(np.abs(g["gslice." + k]).max()))
        err = float((sl - torch.from_numpy(g["postslice." + k])).abs().max())
        assert err <= SLICE_TOL * coef * gscale + 2e-7 * float(q.abs().max()), (k, err, gscale)
    # MI on the original parameters (eval-mode forward, encoder.py:111-145)
    params = [p[k].cuda().contiguous() for k in O.ALL_KEYS]
    m2, l2 = eng.encode_stats(params, xc)
    mi = float(eng.mi(m2, l2, torch.from_numpy(g["eps_mi"]).cuda()))
    assert abs(mi - float(g["mi"])) <= OUT_TOL * max(1.0, abs(float(g["mi"])))


def test_yahoo_fused_inner_step_train_mode(golden):
    """The call bench.py's `value` leg times (lagvae_text_inner_step; decoder weight gradients one bf16 pass, norm only)
    with the reference's masks: per-sentence loss, Σloss, clip norm, post-step encoder parameters."""
    import lagvae
    g, (V, ni, nh, nz, B, T), p, x, m_in, m_out = _yahoo(golden)
    eng = lagvae.TextEngine(V, ni, nh, nz, "cuda")
    params = [p[k].cuda().contiguous() for k in O.ALL_KEYS]
    drop = lagvae.DropoutSpec(1, 0.5, 0.5, m_in.cuda().contiguous(), m_out.cuda().contiguous(), 0)
    gw = eng.grad_workspace()
    out_loss, sc = torch.empty(B, device="cuda"), torch.empty(4, device="cuda")
    eng.inner_step(params, x.cuda(), torch.from_numpy(g["eps"]).cuda(), float(g["kl_weight"]), drop, gw, out_loss, sc)
    assert_close(out_loss, g["loss"], OUT_TOL, "loss")
    assert abs(float(sc[0]) - float(g["loss"].sum())) <= OUT_TOL * abs(float(g["loss"].sum()))
    assert abs(float(sc[3]) - float(g["grad_norm"])) <= 1e-3 * float(g["grad_norm"])
    for i, k in enumerate(O.ENC_KEYS):
        assert abs(float(params[i].double().norm()) - float(g["postnorm." + k])) <= 1e-5 * float(g["postnorm." + k]), k


def test_yahoo_philox_dropout_vs_oracle_on_the_same_masks(golden):
    """mode 2 (what bench.py runs) at the full shape: the oracle is fed the masks `lagvae_dropout_mask` materialises."""
    import lagvae
    import lagvae._backend as be
    g, (V, ni, nh, nz, B, T), p, x, _, _ = _yahoo(golden)
    seed = 0x5EED0BEEF
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    m_in = torch.empty(B, T - 1, ni, dtype=torch.uint8, device="cuda")
    m_out = torch.empty(B, T - 1, nh, dtype=torch.uint8, device="cuda")
    be.check(be.lib().lagvae_dropout_mask(seed, 1, m_in.numel(), 0.5, be.ptr(m_in), st))
    be.check(be.lib().lagvae_dropout_mask(seed, 2, m_out.numel(), 0.5, be.ptr(m_out), st))
    eps = torch.from_numpy(g["eps"])
    klw = float(g["kl_weight"])
    eng = lagvae.TextEngine(V, ni, nh, nz, "cuda")
    params = [p[k].cuda().contiguous() for k in O.ALL_KEYS]
    loss, rec, kl = eng.loss_forward(params, x.cuda(), eps.cuda(), klw, lagvae.DropoutSpec(2, 0.5, 0.5, None, None, seed))
    grads = eng.loss_backward(params, x.cuda(), torch.full((B,), 1.0 / B, device="cuda"), None, None)
    torch.cuda.synchronize()
    # oracle on the host (explicit cell loop + autograd), same masks
    torch.set_num_threads(max(1, min(32, torch.get_num_threads())))
    r = O.inner_step({k: v.clone() for k, v in p.items()}, x, klw, eps, m_in.cpu().float() * 2, m_out.cpu().float() * 2, update=False)
    assert_close(loss, r["loss"], OUT_TOL, "loss (philox)")
    assert_close(rec, r["rec"], OUT_TOL, "rec (philox)")
    assert_close(kl, r["kl"], OUT_TOL, "kl (philox)", floor=1e-2)
    tot = 0.0
    for k, gr in zip(O.ALL_KEYS, grads):
        want = r["grads"][k]
        n, wn = float(gr.double().norm()), float(want.double().norm())
        tot += n * n
        assert abs(n - wn) <= NORM_TOL * max(wn, 1e-6), (k, n, wn)
        assert_close(gr, want, 5e-3, "grad " + k, floor=1e-7)
    assert abs(tot ** 0.5 - r["grad_norm"]) <= 1e-3 * r["grad_norm"]


@pytest.mark.parametrize("Bd,Tn", [(32, 200), (128, 64)])
def test_lstm_v2_vs_float64_autograd(Bd, Tn):
    """k_lstm_v2 forward + backward (nh=1024, initial state, dropout on the emitted h, extra gradient on the last h)
    against float64 autograd of oracle.lstm_sequence on the host."""
    import lagvae
    import lagvae._backend as be
    nh = 1024
    gen = torch.Generator().manual_seed(Bd * 7 + Tn)
    w_hh = ((torch.rand(4 * nh, nh, generator=gen) * 2 - 1) * (2.0 / nh ** 0.5))
    pre = torch.randn(Tn, Bd, 4 * nh, generator=gen)
    h0 = torch.tanh(torch.randn(Bd, nh, generator=gen))
    c0 = torch.randn(Bd, nh, generator=gen)
    keep = (torch.rand(Bd, Tn, nh, generator=gen) > 0.5)
    dh_ext = torch.randn(Tn, Bd, nh, generator=gen) * 0.1          # gradient wrt the dropped-out h
    dh_last = torch.randn(Bd, nh, generator=gen) * 0.1
    # ---- float64 reference: lstm_sequence takes the input projection through an identity W_ih
    W = w_hh.double().requires_grad_(True)
    P = pre.double().transpose(0, 1).contiguous().requires_grad_(True)       # [B, T, 4nh]
    H0, C0 = h0.double().requires_grad_(True), c0.double().requires_grad_(True)
    eye = torch.eye(4 * nh, dtype=torch.float64)
    zb = torch.zeros(4 * nh, dtype=torch.float64)
    hs, h_last, _ = O.lstm_sequence(P, eye, W, zb, zb, H0, C0)
    hdrop = hs * keep.double() * 2.0
    obj = (hdrop * dh_ext.double().transpose(0, 1)).sum() + (h_last * dh_last.double()).sum()
    gP, gH0, gC0 = torch.autograd.grad(obj, [P, H0, C0])
    # ---- kernels
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    drop = be.Dropout()
    mask = keep.to(torch.uint8).cuda().contiguous()
    drop.mode, drop.p_in, drop.p_out, drop.mask_out = 1, 0.0, 0.5, mask.data_ptr()
    ws = torch.zeros(int(be.lib().lagvae_lstm_workspace_bytes(nh, Bd)), dtype=torch.uint8, device="cuda")
    gates = pre.reshape(Tn * Bd, 4 * nh).cuda().contiguous()
    c_all, h_all, hd = (torch.zeros(Tn * Bd, nh, device="cuda") for _ in range(3))
    w_d, h0_d, c0_d = w_hh.cuda(), h0.cuda(), c0.cuda()
    be.check(be.lib().lagvae_lstm_forward(1, nh, Tn, Bd, be.ptr(w_d), be.ptr(h0_d), be.ptr(c0_d), be.ptr(gates), be.ptr(c_all),
                                          be.ptr(h_all), be.ptr(hd), C.byref(drop), be.ptr(ws), ws.numel(), st), "lstm_forward")
    torch.cuda.synchronize()
    # Bd = 128: clusters of 4 do not fit (128 KB resident W_hh + 72 KB receive slots + ring > 227 KB), clusters of 2 do
    assert lagvae.lstm_variant()["forward"] == "v2/cs2", lagvae.lstm_variant()
    dc, dhr, dg = torch.zeros(Bd, nh, device="cuda"), torch.zeros(Bd, nh, device="cuda"), torch.zeros(Tn * Bd, 4 * nh, device="cuda")
    de, dl = dh_ext.reshape(Tn * Bd, nh).cuda().contiguous(), dh_last.cuda()
    be.check(be.lib().lagvae_lstm_backward(1, nh, Tn, Bd, be.ptr(w_d), be.ptr(c0_d), be.ptr(gates), be.ptr(c_all), be.ptr(de),
                                           be.ptr(dl), C.byref(drop), be.ptr(dc), be.ptr(dhr), be.ptr(dg), 1, be.ptr(ws),
                                           ws.numel(), st), "lstm_backward")
    torch.cuda.synchronize()
    assert lagvae.lstm_variant()["backward"].startswith("v2"), lagvae.lstm_variant()
    # outputs 1e-4 (the forward values feed loss/KL); gradients 1e-3 of the tensor max (200 dependent steps)
    assert_close(h_all.view(Tn, Bd, nh).transpose(0, 1), hs.detach(), 1e-4, "h")
    assert_close(hd.view(Tn, Bd, nh).transpose(0, 1), hdrop.detach(), 1e-4, "hdrop")
    assert_close(dg.view(Tn, Bd, 4 * nh).transpose(0, 1), gP, 1e-3, "dgates")
    assert_close(dhr, gH0, 1e-3, "dh_init")
    assert_close(dc, gC0, 1e-3, "dc_init")
