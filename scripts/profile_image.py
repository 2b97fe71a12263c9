#!/usr/bin/env python
"""Kernel-time breakdown of one image inner step (image.py:300-314 sequence through the drop-in modules) with
torch.profiler (CUPTI sees the liblagvae.so launches too).  Prints kernels aggregated by name, sorted by device time."""
import os
import re
import sys
import types
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "vae-lagging-encoder_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch
import image_oracle as IO
import modules

B, NZ = int(os.environ.get("IMG_B", "64")), 32
dev = torch.device("cuda")
a = types.SimpleNamespace(nz=NZ, latent_feature_map=4, device=dev)
torch.manual_seed(0)
vae = modules.VAE(modules.ResNetEncoderV2(a), modules.PixelCNNDecoderV2(a), a).to(dev).train()
enc_opt = torch.optim.Adam(vae.encoder.parameters(), lr=0.001)
xs = [IO.make_image_batch(B, seed=10 + i).to(dev) for i in range(4)]


def step(i):
    vae.zero_grad(set_to_none=True)
    loss, rc, kl = vae.loss(xs[i % 4], 0.1, nsamples=1)
    s = loss.sum().item()
    loss.mean(dim=-1).backward()
    torch.nn.utils.clip_grad_norm_(vae.parameters(), 5.0)
    enc_opt.step()
    return s


for i in range(int(os.environ.get("IMG_WARM", "3"))):
    step(i)
torch.cuda.synchronize()
if os.environ.get("IMG_NCU") == "1":      # under ncu: one more step is the profiled range (scripts/gpu_imgprof.sh)
    step(0)
    torch.cuda.synchronize()
    sys.exit(0)
N = 3
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for i in range(N):
        step(i)
    torch.cuda.synchronize()
agg = defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        m = re.search(r"(k_\w+(<[^(]*>)?)", ev.name)
        name = m.group(1) if m else ev.name.split("(")[0][:90]
        agg[name][0] += 1
        agg[name][1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(v[1] for v in agg.values())
print("total device time per step: %.3f ms over %d kernels/step" % (tot / N / 1e3, sum(v[0] for v in agg.values()) // N))
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%7.3f ms %5.1f%% %5d x %8.1f us  %s" % (t / N / 1e3, 100 * t / tot, n // N, t / n, name))
