"""Pins oracle/image_oracle.py against the UNMODIFIED reference image modules (CPU) and writes the fixtures
tests/golden/omniglot_*.npz.  Run in the authoring container only:  python oracle/validate_image_against_reference.py"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import image_oracle as IO  # noqa: E402
from validate_against_reference import load_reference_modules  # noqa: E402


def build(ref, nz, seed=0):
    torch.manual_seed(seed)
    args = types.SimpleNamespace(nz=nz, latent_feature_map=4, device=torch.device("cpu"))
    return ref.VAE(ref.ResNetEncoderV2(args), ref.PixelCNNDecoderV2(args), args)


def run_case(ref, name, B, nz, ns, klw, out_dir):
    vae = build(ref, nz)
    vae.train()                                        # BatchNorm uses batch statistics (image.py trains in train())
    # parameters come from the oracle's seeded generator so that the fixtures need not carry 2.3 M floats; the spec
    # (keys, order, shapes) is pinned against the reference state_dict here
    p0 = IO.init_image_params(nz, seed=0)
    ref_sd = vae.state_dict()
    assert list(ref_sd.keys()) == list(p0.keys()), "image_param_spec differs from the reference state_dict"
    assert all(tuple(ref_sd[k].shape) == tuple(p0[k].shape) for k in p0)
    vae.load_state_dict(p0)
    x = IO.make_image_batch(B)
    sd0 = {k: v.clone() for k, v in vae.state_dict().items()}
    torch.manual_seed(1)
    loss, rec, kl = vae.loss(x, klw, nsamples=ns)
    vae.zero_grad()
    loss.mean(dim=-1).backward()
    grads = {n: (q.grad.clone() if q.grad is not None else torch.zeros_like(q)) for n, q in vae.named_parameters()}
    gnorm = float(torch.nn.utils.clip_grad_norm_(vae.parameters(), 5.0))
    sd1 = vae.state_dict()
    torch.manual_seed(1)
    eps = torch.zeros(B, ns, nz).normal_()
    # --- oracle on the pre-forward parameters (masked taps are zeroed in place by the reference's forward; the oracle
    #     applies the mask to the value only, like the reference's `weight.data.mul_`)
    p = {k: v.clone().requires_grad_(v.dtype.is_floating_point and ("running" not in k) and ("mask" not in k))
         for k, v in sd0.items()}
    o_loss, o_rec, o_kl = IO.vae_loss(p, x, klw, eps)
    o_loss.mean().backward()
    err = float((o_loss - loss).abs().max() / loss.abs().max())
    assert err < 2e-5, (name, "loss", err)
    assert float((o_kl - kl).abs().max()) < 1e-5 * max(1.0, float(kl.abs().max()))
    gerr = 0.0
    for n, g_ref in grads.items():
        go = p[n].grad if p[n].grad is not None else torch.zeros_like(g_ref)
        # masked taps included: the oracle reproduces the reference's non-zero gradients there
        e = float((go - g_ref).abs().max() / (g_ref.abs().max() + 1e-30))
        gerr = max(gerr, e)
        assert e < 2e-3, (name, n, e)
    print("[%s] loss.sum=%.6f rec.sum=%.6f KL.sum=%.6e gnorm=%.6f (oracle rel err loss %.1e, grads %.1e)" %
          (name, float(loss.sum()), float(rec.sum()), float(kl.sum()), gnorm, err, gerr))
    out = {"meta": np.array([B, nz, ns], dtype=np.int64), "kl_weight": np.float64(klw), "x": x.numpy(), "eps": eps.numpy(),
           "loss": loss.detach().numpy(), "rec": rec.detach().numpy(), "kl": kl.detach().numpy(), "grad_norm": np.float64(gnorm)}
    names = []
    for n, gq in grads.items():
        names.append(n)
        out["gnorm." + n] = np.float64(gq.double().norm())
        out["gslice." + n] = gq.reshape(-1)[:: max(1, gq.numel() // 32)][:32].numpy()
    out["names"] = np.array(names)
    for k in sd1:
        if "running_" in k:
            out["post." + k] = sd1[k].numpy()
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **out)


def run_eval_case(ref, name, B, nz, ns, klw, out_dir):
    """eval() forward (image.py test(), MI, ancestral sampling): BatchNorm normalises with its running statistics.  The
    statistics are moved away from (0, 1) by a few train()-mode forwards of the reference, stored in the fixture, and the
    reference's eval() loss is compared with the oracle's (training=False)."""
    vae = build(ref, nz)
    vae.load_state_dict(IO.init_image_params(nz, seed=0))
    vae.train()
    with torch.no_grad():
        for i in range(4):
            vae.loss(IO.make_image_batch(8, seed=300 + i), 1.0)
    vae.eval()
    sd = {k: v.clone() for k, v in vae.state_dict().items()}
    x = IO.make_image_batch(B, seed=77)
    torch.manual_seed(1)
    with torch.no_grad():
        loss, rec, kl = vae.loss(x, klw, nsamples=ns)
    torch.manual_seed(1)
    eps = torch.zeros(B, ns, nz).normal_()
    with torch.no_grad():
        o_loss, o_rec, o_kl = IO.vae_loss(sd, x, klw, eps, training=False)
    err = float((o_loss - loss).abs().max() / loss.abs().max())
    assert err < 2e-5, (name, err)
    assert float((o_kl - kl).abs().max()) < 1e-5 * max(1.0, float(kl.abs().max()))
    print("[%s] eval loss.sum=%.6f rec.sum=%.6f KL.sum=%.6e (oracle rel err %.1e)" % (name, float(loss.sum()), float(rec.sum()), float(kl.sum()), err))
    out = {"meta": np.array([B, nz, ns], dtype=np.int64), "kl_weight": np.float64(klw), "x": x.numpy(), "eps": eps.numpy(),
           "loss": loss.numpy(), "rec": rec.numpy(), "kl": kl.numpy()}
    for k, v in sd.items():
        if "running_" in k or "num_batches" in k:
            out["stat." + k] = v.numpy()
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **out)


def main():
    torch.set_num_threads(os.cpu_count())
    ref = load_reference_modules()
    out_dir = os.path.join(ROOT, "tests", "golden")
    run_case(ref, "omniglot_b8", 8, 32, 1, 0.1, out_dir)
    run_case(ref, "omniglot_b3_ns2", 3, 8, 2, 1.0, out_dir)
    run_eval_case(ref, "omniglot_eval_b5", 5, 8, 1, 1.0, out_dir)


if __name__ == "__main__":
    main()
