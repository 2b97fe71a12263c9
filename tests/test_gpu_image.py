"""Parity of the image path (ResNetEncoderV2 + PixelCNNDecoderV2, training-mode BatchNorm)
against the image oracle and the fixtures written from the unmodified reference."""
import types

import numpy as np
import pytest
import torch

import image_oracle as IO
from util import assert_close

pytestmark = pytest.mark.gpu


def _build(nz):
    import modules
    a = types.SimpleNamespace(nz=nz, latent_feature_map=4, device=torch.device("cuda"))
    vae = modules.VAE(modules.ResNetEncoderV2(a), modules.PixelCNNDecoderV2(a), a).to("cuda")
    p = IO.init_image_params(nz, seed=0)
    assert list(vae.state_dict().keys()) == list(p.keys())          # reference state_dict compatibility
    vae.load_state_dict(p)
    return vae, p


@pytest.mark.parametrize("name", ["omniglot_b8", "omniglot_b3_ns2"])
def test_image_loss_and_grads_vs_reference_golden(golden, name):
    from modules.image import _ReparamKLFn
    g = golden(name)
    B, nz, ns = [int(v) for v in g["meta"]]
    vae, p = _build(nz)
    vae.train()
    x = torch.from_numpy(g["x"]).cuda()
    eps = torch.from_numpy(g["eps"]).cuda()
    klw = float(g["kl_weight"])
    mu, logvar = vae.encoder(x)
    z, kl = _ReparamKLFn.apply(mu, logvar, eps)
    rec = vae.decoder.reconstruct_error(x, z).mean(dim=1)
    loss = rec + klw * kl
    assert_close(loss, g["loss"], 1e-4, "loss")
    assert_close(rec, g["rec"], 1e-4, "rec")
    assert_close(kl, g["kl"], 1e-4, "kl", floor=1e-2)
    loss.mean(dim=-1).backward()
    bad, tot = [], 0.0
    grads = dict(vae.named_parameters())
    for n in [str(s) for s in g["names"]]:
        gr = grads[n].grad
        nrm = float(gr.double().norm())
        tot += nrm * nrm
        want = float(g["gnorm." + n])
        sl = gr.reshape(-1)[:: max(1, gr.numel() // 32)][:32].cpu().double()
        ws = torch.from_numpy(g["gslice." + n]).double()
        serr = float((sl - ws).abs().max()) / max(float(gr.abs().max()) * 0.05, float(ws.abs().max()), 1e-12)
        if abs(nrm - want) > 3e-3 * max(want, 1e-6) or serr > 1e-2:
            bad.append("%s: norm %.6g want %.6g, slice err %.2e" % (n, nrm, want, serr))
    assert not bad, "\n".join(bad)
    assert abs(tot ** 0.5 - float(g["grad_norm"])) <= 2e-3 * float(g["grad_norm"])
    # running statistics were updated exactly once, with the unbiased variance
    sd = vae.state_dict()
    for k in g:
        if k.startswith("post."):
            assert_close(sd[k[5:]], g[k], 1e-4, k, floor=1e-3)
    # masked taps are zeroed in the weights by the forward (dec_pixelcnn_v2.py:29)
    w, m = sd["decoder.main.0.main.1.main.3.weight"], sd["decoder.main.0.main.1.main.3.mask"]
    assert float((w * (1 - m)).abs().max()) == 0.0


def test_image_module_api_inner_step_vs_oracle():
    """The statement sequence of image.py:300-314 (Adam encoder step) against the oracle on the same draws."""
    B, nz = 4, 16
    vae, p = _build(nz)
    vae.train()
    x = IO.make_image_batch(B, seed=5)
    torch.manual_seed(3)
    eps = torch.empty(B, 1, nz, device="cuda").normal_()
    leaves = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k and "mask" not in k) for k, v in p.items()}
    o_loss, o_rec, o_kl = IO.vae_loss(leaves, x, 0.3, eps.cpu())
    o_loss.mean().backward()
    enc_opt = torch.optim.Adam(vae.encoder.parameters(), lr=0.001)
    dec_opt = torch.optim.Adam(vae.decoder.parameters(), lr=0.001)
    enc_opt.zero_grad()
    dec_opt.zero_grad()
    torch.manual_seed(3)
    loss, loss_rc, loss_kl = vae.loss(x.cuda(), 0.3, nsamples=1)
    s = loss.sum().item()
    loss.mean(dim=-1).backward()
    torch.nn.utils.clip_grad_norm_(vae.parameters(), 5.0)
    enc_opt.step()
    assert abs(s - float(o_loss.sum())) <= 1e-4 * abs(float(o_loss.sum()))
    assert_close(loss_kl, o_kl.detach(), 1e-4, "kl", floor=1e-2)
    with torch.no_grad():
        vae.eval()
        l2, _, _ = vae.loss(x.cuda(), 1.0)
        assert bool(torch.isfinite(l2).all()) and not l2.requires_grad
        assert isinstance(vae.calc_mi_q(x.cuda()), float)


@pytest.mark.parametrize("B,nz", [(4, 8), (5, 32)])
def test_image_all_gradients_vs_oracle(B, nz):
    """Every one of the 248 gradients (masked taps included) against the image oracle on ragged batch
    sizes (28 / 35 pixel tiles: fewer tiles than SMs, uneven wgrad partitions)."""
    vae, p = _build(nz)
    vae.train()
    x = IO.make_image_batch(B, seed=5)
    torch.manual_seed(3)
    eps = torch.empty(B, 1, nz, device="cuda").normal_()
    leaves = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k and "mask" not in k) for k, v in p.items()}
    o_loss, _, _ = IO.vae_loss(leaves, x, 0.3, eps.cpu())
    o_loss.mean().backward()
    torch.manual_seed(3)
    loss, _, _ = vae.loss(x.cuda(), 0.3, nsamples=1)
    loss.mean(dim=-1).backward()
    assert_close(loss.detach(), o_loss.detach(), 1e-4, "loss")
    bad = []
    for n, q in vae.named_parameters():
        want = leaves[n].grad.double()
        d = float((q.grad.double().cpu() - want).norm())
        if d > 5e-3 * max(float(want.norm()), 1e-9):
            bad.append("%s: |diff| %.3g vs |grad| %.3g" % (n, d, float(want.norm())))
    assert not bad, "\n".join(bad[:10])


def test_pixelcnn_ancestral_sampling_is_causally_consistent():
    """PixelCNNDecoderV2.decode (dec_pixelcnn_v2.py:201-232) in eval() mode: because of the causal masks the
    probability of pixel (i, j) given the finished image equals the one it was decided from, so the deterministic decode
    must be a fixed point: x == (p >= 0.5) at every pixel (fails if a mask leaks, or the loop order / channel is wrong)."""
    vae, p = _build(8)
    vae.train()
    with torch.no_grad():                                     # warm the BatchNorm running statistics (momentum 0.1)
        for i in range(30):
            vae.loss(IO.make_image_batch(8, seed=40 + i).cuda(), 1.0)
    vae.eval()
    torch.manual_seed(4)
    z = torch.randn(3, 8, device="cuda")
    x, probs = vae.decoder.decode(z, True)
    assert x.shape == (3, 1, 28, 28) and probs.shape == (3, 1, 28, 28)
    margin = (probs - 0.5).abs() > 1e-4                       # ignore exact-threshold ties
    assert bool(((x > 0.5) == (probs >= 0.5))[margin].all())
    torch.manual_seed(5)
    xs, _ = vae.decoder.decode(z, False)
    assert set(xs.unique().tolist()) == {0.0, 1.0}                # Bernoulli draws: both values occur


def test_image_eval_mode_fused_matches_first_tier(monkeypatch):
    """eval() forward (BatchNorm running statistics; image.py test(), calc_mi, ancestral sampling) on the fused tcgen05
    path against the first-tier per-layer kernels on the same model."""
    vae, p = _build(8)
    vae.train()
    with torch.no_grad():
        for i in range(5):                                    # move the running statistics away from (0, 1)
            vae.loss(IO.make_image_batch(8, seed=60 + i).cuda(), 1.0)
    vae.eval()
    x = IO.make_image_batch(5, seed=7).cuda()
    sd_before = {k: v.clone() for k, v in vae.state_dict().items()}
    with torch.no_grad():
        torch.manual_seed(9)
        l_fused, r_fused, k_fused = vae.loss(x, 1.0)
        monkeypatch.setenv("LAGVAE_IMAGE_FUSED", "0")
        torch.manual_seed(9)
        l_tier1, r_tier1, k_tier1 = vae.loss(x, 1.0)
    assert_close(r_fused, r_tier1, 1e-4, "rec (eval)")
    assert_close(l_fused, l_tier1, 1e-4, "loss (eval)")
    for k, v in vae.state_dict().items():                    # eval() updates no statistics
        assert torch.equal(v, sd_before[k]), k


def test_image_full_batch_loss_and_grad_norm_vs_oracle():
    """BASELINE.json configs[3] at its full size (Omniglot shape, batch 64, nz 32): per-image loss / rec / KL and the
    clip norm of all gradients against the image oracle on the same draws (448 pixel tiles > #SMs: multi-wave tiles,
    multi-block wgrad partitions)."""
    B, nz = 64, 32
    vae, p = _build(nz)
    vae.train()
    x = IO.make_image_batch(B, seed=21)
    torch.manual_seed(6)
    eps = torch.empty(B, 1, nz, device="cuda").normal_()
    leaves = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k and "mask" not in k) for k, v in p.items()}
    o_loss, o_rec, o_kl = IO.vae_loss(leaves, x, 0.1, eps.cpu())
    o_loss.mean().backward()
    o_norm = sum(float(v.grad.double().norm()) ** 2 for v in leaves.values() if v.grad is not None) ** 0.5
    torch.manual_seed(6)
    loss, rec, kl = vae.loss(x.cuda(), 0.1, nsamples=1)
    loss.mean(dim=-1).backward()
    assert_close(loss.detach(), o_loss.detach(), 1e-4, "loss")
    assert_close(rec.detach(), o_rec.detach(), 1e-4, "rec")
    assert_close(kl.detach(), o_kl.detach(), 1e-4, "kl", floor=1e-2)
    norm = sum(float(q.grad.double().norm()) ** 2 for q in vae.parameters()) ** 0.5
    assert abs(norm - o_norm) <= 2e-3 * o_norm, (norm, o_norm)


def test_graphed_step_equals_eager_steps():
    """lagvae.GraphedStep (the image inner step of image.py:300-314 captured as ONE CUDA graph) against the same statements
    run eagerly: same Σloss per step, same encoder parameters and BatchNorm running statistics afterwards."""
    import lagvae
    vae_a, _ = _build(8)
    vae_b, _ = _build(8)
    vae_a.train()
    vae_b.train()
    # SGD, not Adam: Adam normalises noise-level gradient elements (the encoder's last BatchNorm reduces over 8 rows) to
    # +-lr steps whose sign depends on the fp64-atomic summation order, in eager mode as much as in the graph
    opt_a = torch.optim.SGD(vae_a.encoder.parameters(), lr=1e-3)
    opt_b = torch.optim.SGD(vae_b.encoder.parameters(), lr=1e-3)
    xs = [IO.make_image_batch(8, seed=80 + i).cuda() for i in range(3)]

    def body(vae, opt, x):
        opt.zero_grad(set_to_none=True)
        loss, _, _ = vae.loss(x, 0.3, nsamples=1)
        loss.mean(dim=-1).backward()
        torch.nn.utils.clip_grad_norm_(list(vae.parameters()), 5.0)
        opt.step()
        return loss.sum()

    torch.manual_seed(100)
    for _ in range(3):                                        # mirrors the helper's three eager warm-up steps
        body(vae_a, opt_a, xs[0])
    torch.manual_seed(100)
    step = lagvae.GraphedStep(lambda x: body(vae_b, opt_b, x), {"x": xs[0]}, warmup=3)
    for i in (1, 2):
        torch.manual_seed(200 + i)
        sa = float(body(vae_a, opt_a, xs[i]).detach())
        torch.manual_seed(200 + i)
        sb = float(step(x=xs[i]).detach())
        assert abs(sa - sb) <= 1e-5 * abs(sa), (i, sa, sb)
    sd_a, sd_b = vae_a.state_dict(), vae_b.state_dict()
    for k in sd_a:
        if sd_a[k].dtype.is_floating_point:
            assert_close(sd_b[k], sd_a[k], 1e-4, k, floor=1e-2)
        else:
            assert torch.equal(sd_a[k], sd_b[k]), k


def test_image_eval_mode_vs_reference_golden(golden):
    """eval() forward (running statistics; image.py test(), MI, sampling) on the fused tcgen05 path against the unmodified
    reference's loss / rec / KL (fixture written by oracle/validate_image_against_reference.py::run_eval_case)."""
    from modules.image import _ReparamKLFn
    g = golden("omniglot_eval_b5")
    B, nz, ns = [int(v) for v in g["meta"]]
    vae, p = _build(nz)
    sd = vae.state_dict()
    for k in g:
        if k.startswith("stat."):
            sd[k[5:]] = torch.from_numpy(g[k]).cuda()
    vae.load_state_dict(sd)
    vae.eval()
    x = torch.from_numpy(g["x"]).cuda()
    with torch.no_grad():
        mu, logvar = vae.encoder(x)
        z, kl = _ReparamKLFn.apply(mu, logvar, torch.from_numpy(g["eps"]).cuda())
        rec = vae.decoder.reconstruct_error(x, z).mean(dim=1)
    assert_close(rec, g["rec"], 1e-4, "rec (eval)")
    assert_close(kl, g["kl"], 1e-4, "kl (eval)", floor=1e-2)
    assert_close(rec + float(g["kl_weight"]) * kl, g["loss"], 1e-4, "loss (eval)")


def test_outgrown_buffers_are_retired_while_a_graph_is_alive():
    """a captured graph bakes raw pointers into the scratch buffers; a later, larger eager call must not free them."""
    import lagvae
    import lagvae.graph as G
    from modules import image as I
    dev = torch.device("cuda")
    x = torch.zeros(4, device=dev)
    step = lagvae.GraphedStep(lambda x: x * 2.0, {"x": x}, warmup=1)
    tag = "test-retire"
    a = I._scratch(1 << 20, tag, dev)
    ptr = a.data_ptr()
    del a
    n0 = len(G._RETIRED)
    b = I._scratch(8 << 20, tag, dev)                      # outgrows the buffer while `step` is alive
    assert len(G._RETIRED) == n0 + 1 and G._RETIRED[-1].data_ptr() == ptr and b.data_ptr() != ptr
    assert float(step(x=torch.ones(4, device=dev)).sum()) == 8.0
    del step
