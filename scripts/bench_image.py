#!/usr/bin/env python
"""Image inner step (BASELINE.json configs[3]: Omniglot ResNet-enc + PixelCNN-dec VAE, aggressive=1, batch 64, 1 GPU):
one iteration of image.py:300-318 — zero_grad, vae.loss, Σloss.item(), backward, clip_grad_norm_(all, 5.0), Adam step
on the encoder — through the drop-in `modules` API, next to the same model expressed with torch's stock cuDNN ops
(oracle/image_oracle.py functional port, informational).  Prints one JSON line.  Not the headline metric."""
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "vae-lagging-encoder_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch
import image_oracle as IO
import modules
import lagvae

B, NZ, STEPS, WARM = int(os.environ.get("IMG_B", "64")), 32, int(os.environ.get("IMG_STEPS", "10")), 3
dev = torch.device("cuda")
a = types.SimpleNamespace(nz=NZ, latent_feature_map=4, device=dev)
torch.manual_seed(0)
vae = modules.VAE(modules.ResNetEncoderV2(a), modules.PixelCNNDecoderV2(a), a).to(dev).train()
enc_opt = torch.optim.Adam(vae.encoder.parameters(), lr=0.001)
dec_opt = torch.optim.Adam(vae.decoder.parameters(), lr=0.001)
xs = [IO.make_image_batch(B, seed=10 + i).to(dev) for i in range(8)]


def step(i):
    enc_opt.zero_grad()
    dec_opt.zero_grad()
    loss, rc, kl = vae.loss(xs[i % 8], 0.1, nsamples=1)
    s = loss.sum().item()
    loss.mean(dim=-1).backward()
    torch.nn.utils.clip_grad_norm_(vae.parameters(), 5.0)
    enc_opt.step()
    return s


for i in range(WARM):
    step(i)
torch.cuda.synchronize()
l0 = lagvae.launch_count()
t0 = time.perf_counter()
for i in range(STEPS):
    step(i)
torch.cuda.synchronize()
ours = STEPS / (time.perf_counter() - t0)
launches = (lagvae.launch_count() - l0) / STEPS

# ---- the same step captured in ONE CUDA graph (forward + backward + clip + Adam; only Σloss is read back) ----------------
def graph_mode(n_steps):
    torch.manual_seed(0)
    vae_g = modules.VAE(modules.ResNetEncoderV2(a), modules.PixelCNNDecoderV2(a), a).to(dev).train()
    e_opt = torch.optim.Adam(vae_g.encoder.parameters(), lr=0.001, capturable=True)
    d_opt = torch.optim.Adam(vae_g.decoder.parameters(), lr=0.001, capturable=True)
    allp = list(vae_g.parameters())
    x_static = xs[0].clone()
    s_static = torch.zeros((), device=dev)

    def body():
        e_opt.zero_grad(set_to_none=True)
        d_opt.zero_grad(set_to_none=True)
        loss, rc, kl = vae_g.loss(x_static, 0.1, nsamples=1)
        s_static.copy_(loss.sum())
        loss.mean(dim=-1).backward()
        torch.nn.utils.clip_grad_norm_(allp, 5.0)
        e_opt.step()

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                     # warm-up on a side stream (torch.cuda.graph recipe)
        for _ in range(3):
            body()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    l_before = lagvae.launch_count()
    with torch.cuda.graph(g):
        body()
    nodes = lagvae.launch_count() - l_before
    for i in range(WARM):
        x_static.copy_(xs[i % 8])
        g.replay()
        s_static.item()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n_steps):
        x_static.copy_(xs[i % 8])                     # next batch (image.py:316-318)
        g.replay()
        s = s_static.item()                           # Σloss readback every step (image.py:306)
    torch.cuda.synchronize()
    return n_steps / (time.perf_counter() - t0), nodes, s


graph = None
if os.environ.get("IMG_GRAPH", "1") == "1":
    try:
        gv, gnodes, gs = graph_mode(STEPS)
        graph = {"value": gv, "ms_per_step": 1e3 / gv, "lagvae_kernels_in_graph": gnodes, "loss_sum": gs}
    except Exception as ex:
        graph = {"error": repr(ex)[:400]}
        torch.cuda.synchronize()


# torch stock ops on the same GPU (functional port; cuDNN conv / batch_norm)
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


for i in range(WARM):
    step_ref(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(STEPS):
    step_ref(i)
torch.cuda.synchronize()
ref = STEPS / (time.perf_counter() - t0)
print(json.dumps({"metric": "aggressive inner-loop encoder steps/sec (Omniglot ResNet+PixelCNN VAE, batch %d)" % B, "value": ours,
                  "unit": "steps/s", "ms_per_step": 1e3 / ours, "kernel_launches_per_step": launches, "cuda_graph": graph,
                  "torch_gpu_port": {"value": ref, "what": "oracle/image_oracle.py functional port on torch cuDNN/cuBLAS, same GPU"},
                  "note": "correctness-first tier of the image rows: one autograd node + 1-4 kernel launches per layer"}))
