#!/usr/bin/env python
"""Image inner step (BASELINE.json configs[3]: Omniglot ResNet-enc + PixelCNN-dec VAE, aggressive=1, batch 64, 1 GPU):
one iteration of image.py:300-318 — zero_grad, vae.loss, Σloss.item(), backward, clip_grad_norm_(all, 5.0), Adam step
on the encoder — through the drop-in `modules` API (eager = what unmodified image.py drives; graph = the same statement
sequence captured in ONE CUDA graph, only Σloss read back), next to the same model expressed with torch's stock cuDNN ops
(oracle/image_oracle.py functional port, informational).  `python scripts/bench_image.py` prints one JSON line; bench.py
embeds `run()` as its `image` object.  Not the headline metric."""
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "vae-lagging-encoder_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch

FLOPS_LIVE_TAP_STEP = 143e9      # image inner step, live taps only, B = 64


def run(B=64, steps=10, warm=3, graph=True, torch_port=True, dev=None):
    import image_oracle as IO
    import lagvae
    import modules
    NZ = 32
    dev = dev or torch.device("cuda")
    a = types.SimpleNamespace(nz=NZ, latent_feature_map=4, device=dev)
    xs = [IO.make_image_batch(B, seed=10 + i).to(dev) for i in range(8)]

    def build(capturable):
        torch.manual_seed(0)
        vae = modules.VAE(modules.ResNetEncoderV2(a), modules.PixelCNNDecoderV2(a), a).to(dev).train()
        eo = torch.optim.Adam(vae.encoder.parameters(), lr=0.001, capturable=capturable)
        do = torch.optim.Adam(vae.decoder.parameters(), lr=0.001, capturable=capturable)
        return vae, eo, do

    # ---- eager: the statement sequence of image.py:300-314 ----
    vae, enc_opt, dec_opt = build(False)

    def step(i):
        enc_opt.zero_grad()
        dec_opt.zero_grad()
        loss, rc, kl = vae.loss(xs[i % 8], 0.1, nsamples=1)
        s = loss.sum().item()
        loss.mean(dim=-1).backward()
        torch.nn.utils.clip_grad_norm_(vae.parameters(), 5.0)
        enc_opt.step()
        return s

    for i in range(warm):
        step(i)
    torch.cuda.synchronize()
    l0 = lagvae.launch_count()
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    torch.cuda.synchronize()
    ours = steps / (time.perf_counter() - t0)
    launches = (lagvae.launch_count() - l0) / steps
    out = {"metric": "aggressive inner-loop encoder steps/sec (Omniglot ResNet+PixelCNN VAE, batch %d)" % B, "value": ours,
           "unit": "steps/s", "ms_per_step": 1e3 / ours, "lagvae_launches_per_step": launches,
           "api": "modules.VAE.loss -> backward -> clip_grad_norm_ -> Adam.step (image.py:300-314 sequence), eager"}
    del vae, enc_opt, dec_opt

    # ---- the same step captured in ONE CUDA graph (forward + backward + clip + Adam; only Σloss is read back) ----
    if graph:
        try:
            vae_g, e_opt, d_opt = build(True)
            allp = list(vae_g.parameters())

            def body_x(x):
                e_opt.zero_grad(set_to_none=True)
                d_opt.zero_grad(set_to_none=True)
                loss, rc, kl = vae_g.loss(x, 0.1, nsamples=1)
                loss.mean(dim=-1).backward()
                torch.nn.utils.clip_grad_norm_(allp, 5.0)
                e_opt.step()
                return loss.sum()

            lb = lagvae.launch_count()
            step_g = lagvae.GraphedStep(lambda x: body_x(x), {"x": xs[0]}, warmup=3)
            nodes = (lagvae.launch_count() - lb) // 4      # 3 eager warm-up runs + the captured one
            for i in range(warm):
                step_g(x=xs[i % 8]).item()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(steps):
                s = step_g(x=xs[i % 8]).item()            # next batch (image.py:316-318) + Σloss readback every step (:306)
            torch.cuda.synchronize()
            gv = steps / (time.perf_counter() - t0)
            out["cuda_graph"] = {"value": gv, "ms_per_step": 1e3 / gv, "lagvae_kernels_in_graph": nodes, "loss_sum": s,
                                 "algorithmic_tflops_live_tap": FLOPS_LIVE_TAP_STEP * (B / 64.0) * gv / 1e12}
            del step_g, vae_g, e_opt, d_opt
        except Exception as ex:   # reported, never hidden
            out["cuda_graph"] = {"error": repr(ex)[:400]}
            torch.cuda.synchronize()

    # ---- torch stock ops on the same GPU (functional port; cuDNN conv / batch_norm) ----
    if torch_port:
        p = {k: (v.to(dev).requires_grad_(True) if v.dtype.is_floating_point and "running" not in k and "mask" not in k else v.to(dev))
             for k, v in IO.init_image_params(NZ, seed=0).items()}
        enc_leaves = [v for k, v in p.items() if k.startswith("encoder.") and v.requires_grad]
        all_leaves = [v for v in p.values() if v.requires_grad]
        opt = torch.optim.Adam(enc_leaves, lr=0.001)

        def step_ref(i):
            for v in all_leaves:
                v.grad = None
            eps = torch.empty(B, 1, NZ, device=dev).normal_()
            loss, _, _ = IO.vae_loss(p, xs[i % 8], 0.1, eps)
            s = loss.sum().item()
            loss.mean().backward()
            torch.nn.utils.clip_grad_norm_(all_leaves, 5.0)
            opt.step()
            return s

        for i in range(warm):
            step_ref(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(steps):
            step_ref(i)
        torch.cuda.synchronize()
        ref = steps / (time.perf_counter() - t0)
        out["torch_gpu_port"] = {"value": ref, "ms_per_step": 1e3 / ref,
                                 "what": "oracle/image_oracle.py functional port on torch cuDNN/cuBLAS, same GPU"}
        del p, opt
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    print(json.dumps(run(B=int(os.environ.get("IMG_B", "64")), steps=int(os.environ.get("IMG_STEPS", "10")),
                         graph=os.environ.get("IMG_GRAPH", "1") == "1")))
