#!/usr/bin/env python
"""Probe for DESIGN §9 item 3: does a 20-CTA tcgen05 GEMM on a side stream run CONCURRENTLY with the persistent
128-CTA cooperative-cluster LSTM recurrence (which leaves 20 of the 148 SMs idle)?  Times the LSTM forward (clusters of
2) and backward (clusters of 4) alone, the GEMM alone, and both (LSTM launched first / GEMM launched first) with CUDA
events; overlap works when t(both) ~= max(t_lstm, t_gemm) instead of the sum.  r1i: forward overlaps fully in both
orders; the backward case is the open question behind the slower LAGVAE_SIDE_WGRAD=1 step (profiles/README.md)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "vae-lagging-encoder_b200"))
import torch
import lagvae._backend as be

L = be.lib()
dev = torch.device("cuda")
nh, Bd, Tn = 1024, 32, 200
w_hh = (torch.rand(4 * nh, nh, device=dev) * 2 - 1) / nh ** 0.5
pre = torch.randn(Tn * Bd, 4 * nh, device=dev)
gates = pre.clone()
c_all, h_all = torch.zeros(Tn * Bd, nh, device=dev), torch.zeros(Tn * Bd, nh, device=dev)
ws = torch.zeros(int(L.lagvae_lstm_workspace_bytes(nh, Bd)), dtype=torch.uint8, device=dev)
drop = be.Dropout()
# 20 output tiles of 128 x 128, long K: ~20 CTAs busy for about as long as the recurrence
M, N, K = 128 * 20, 128, int(os.environ.get("PROBE_K", "98304"))
A = torch.randn(M, K, device=dev).to(torch.bfloat16)
Bm = torch.randn(N, K, device=dev).to(torch.bfloat16)
out = torch.empty(M, N, device=dev)
s_main, s_side = torch.cuda.Stream(), torch.cuda.Stream()


def lstm(stream):
    be.check(L.lagvae_lstm_forward(1, nh, Tn, Bd, be.ptr(w_hh), None, None, be.ptr(gates), be.ptr(c_all), be.ptr(h_all), None,
                                   C.byref(drop), be.ptr(ws), ws.numel(), C.c_void_p(stream.cuda_stream)))


dh = torch.randn(Tn * Bd, nh, device=dev) * 0.01
dc, dhr, dg = torch.zeros(Bd, nh, device=dev), torch.zeros(Bd, nh, device=dev), torch.zeros(Tn * Bd, 4 * nh, device=dev)


def lstm_bwd(stream):
    be.check(L.lagvae_lstm_backward(1, nh, Tn, Bd, be.ptr(w_hh), None, be.ptr(gates), be.ptr(c_all), be.ptr(dh), None, C.byref(drop),
                                    be.ptr(dc), be.ptr(dhr), be.ptr(dg), 1, be.ptr(ws), ws.numel(), C.c_void_p(stream.cuda_stream)))


def gemm(stream):
    be.check(L.lagvae_gemm_tc(be.ptr(A), be.ptr(A), K, 0, be.ptr(Bm), be.ptr(Bm), K, 0, be.ptr(out), N, M, N, K, 1, 1.0, 0.0,
                              None, None, 0, None, C.c_void_p(stream.cuda_stream)))


def timed(fn, reps=5, restore=True):
    ts = []
    for _ in range(reps):
        if restore:
            gates.copy_(pre)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()                       # default stream; both work streams wait on it
        s_main.wait_event(e0)
        s_side.wait_event(e0)
        fn()
        d0, d1 = torch.cuda.Event(), torch.cuda.Event()
        d0.record(s_main)
        d1.record(s_side)
        torch.cuda.current_stream().wait_event(d0)
        torch.cuda.current_stream().wait_event(d1)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for f in (lambda: lstm(s_main), lambda: gemm(s_side)):
    f()
torch.cuda.synchronize()
t_l = timed(lambda: lstm(s_main))
t_g = timed(lambda: gemm(s_side))
t_lg = timed(lambda: (lstm(s_main), gemm(s_side)))
t_gl = timed(lambda: (gemm(s_side), lstm(s_main)))
print("lstm alone %.3f ms | 20-CTA gemm alone %.3f ms | lstm then gemm %.3f ms | gemm then lstm %.3f ms | sum %.3f max %.3f"
      % (t_l, t_g, t_lg, t_gl, t_l + t_g, max(t_l, t_g)))

# backward recurrence (clusters of 4) on the activated gates the forward left behind
gates.copy_(pre)
lstm(s_main)
torch.cuda.synchronize()
lstm_bwd(s_main)
torch.cuda.synchronize()
t_b = timed(lambda: lstm_bwd(s_main), restore=False)
t_bg = timed(lambda: (lstm_bwd(s_main), gemm(s_side)), restore=False)
t_gb = timed(lambda: (gemm(s_side), lstm_bwd(s_main)), restore=False)
print("lstm BWD alone %.3f ms | 20-CTA gemm alone %.3f ms | bwd then gemm %.3f ms | gemm then bwd %.3f ms | sum %.3f max %.3f"
      % (t_b, t_g, t_bg, t_gb, t_b + t_g, max(t_b, t_g)))
