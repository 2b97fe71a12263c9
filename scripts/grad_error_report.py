#!/usr/bin/env python
"""Measured gradient errors of the CUDA path against the reference's golden gradients (tests/golden/*.npz): per tensor
max |got - want| / max |want|, for both kernel tiers.  The numbers justify the tolerances the parity tests assert
(outputs 1e-4 as north_star asks; gradients have their own, measured bar).  Writes gpurun_out/grad_errors.md."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "vae-lagging-encoder_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import lagvae
import lagging_oracle as O
from util import FULL_CASES, case_inputs, case_params

rows = ["| fixture | tier | worst tensor | max abs err / tensor max | loss rel err |", "|---|---|---|---:|---:|"]
for name in FULL_CASES:
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")))
    c = case_inputs(g)
    for simt in (True, False):
        eng = lagvae.TextEngine(c["V"], c["ni"], c["nh"], c["nz"], "cuda", force_simt=simt)
        params = [case_params(g)[k].cuda().contiguous() for k in O.ALL_KEYS]
        drop = lagvae.DropoutSpec()
        if c["train"]:
            drop = lagvae.DropoutSpec(1, 0.5, 0.5, torch.from_numpy(g["mask_in"]).to(torch.uint8).cuda().contiguous(),
                                      torch.from_numpy(g["mask_out"]).to(torch.uint8).cuda().contiguous(), 0)
        loss, rec, kl = eng.loss_forward(params, c["x"].cuda(), c["eps"].cuda(), c["klw"], drop)
        grads = eng.loss_backward(params, c["x"].cuda(), torch.full((c["B"],), 1.0 / c["B"], device="cuda"), None, None)
        worst, wk = 0.0, ""
        for k, gr in zip(O.ALL_KEYS, grads):
            want = torch.from_numpy(g["g." + k]).double()
            e = float((gr.cpu().double() - want).abs().max()) / max(float(want.abs().max()), 1e-30)
            if float(want.abs().max()) > 0 and e > worst:
                worst, wk = e, k
        lerr = float((loss.cpu().double() - torch.from_numpy(g["loss"]).double()).abs().max() / torch.from_numpy(g["loss"]).double().abs().max())
        rows.append("| %s | %s | %s | %.2e | %.2e |" % (name, "simt" if simt else "tcgen05", wk, worst, lerr))
txt = "\n".join(rows)
print(txt)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "grad_errors.md"), "w").write(txt + "\n")
