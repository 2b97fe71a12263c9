#!/usr/bin/env python
"""One forward + one backward launch of the persistent LSTM recurrence at the Yahoo shape (B=32, nh=1024, T=200) — the target of
`ncu --replay-mode application -k regex:k_lstm_v2` (kernel replay cannot re-run a cooperative cluster launch; application
replay re-runs this whole script once per metric pass)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vae-lagging-encoder_b200"))
import torch
import lagvae
import lagvae._backend as be

nh, Bd, Tn = 1024, int(os.environ.get("LSTM_BD", "32")), 200
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
w_hh = (torch.rand(4 * nh, nh, generator=g, device=dev) * 2 - 1) * (3.0 / nh ** 0.5)
gates = torch.randn(Tn * Bd, 4 * nh, generator=g, device=dev)
ws = torch.zeros(int(be.lib().lagvae_lstm_workspace_bytes(nh, Bd)), dtype=torch.uint8, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
drop = be.Dropout()
c_all, h_all = torch.zeros(Tn * Bd, nh, device=dev), torch.zeros(Tn * Bd, nh, device=dev)
be.check(be.lib().lagvae_lstm_forward(1, nh, Tn, Bd, be.ptr(w_hh), None, None, be.ptr(gates), be.ptr(c_all), be.ptr(h_all), None,
                                      C.byref(drop), be.ptr(ws), ws.numel(), st))
dh_ext = torch.randn(Tn * Bd, nh, generator=g, device=dev) * 0.01
dc, dhr, dg = torch.zeros(Bd, nh, device=dev), torch.zeros(Bd, nh, device=dev), torch.zeros(Tn * Bd, 4 * nh, device=dev)
be.check(be.lib().lagvae_lstm_backward(1, nh, Tn, Bd, be.ptr(w_hh), None, be.ptr(gates), be.ptr(c_all), be.ptr(dh_ext), None,
                                       C.byref(drop), be.ptr(dc), be.ptr(dhr), be.ptr(dg), 1, be.ptr(ws), ws.numel(), st))
torch.cuda.synchronize()
print("lstm_only:", lagvae.lstm_variant(), float(h_all.abs().mean()), float(dg.abs().mean()))
