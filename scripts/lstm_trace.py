#!/usr/bin/env python
"""Per-phase clock64 trace of the persistent LSTM kernels (CTA 0) at the Yahoo recurrence shape.
Writes gpurun_out/lstm_trace.txt."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vae-lagging-encoder_b200"))
import torch
import lagvae._backend as be

nh, Bd, Tn = 1024, int(os.environ.get("TRACE_BD", "32")), 200
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
w_hh = (torch.rand(4 * nh, nh, generator=g, device=dev) * 2 - 1) * (3.0 / nh ** 0.5)
pre = torch.randn(Tn * Bd, 4 * nh, generator=g, device=dev)
ws = torch.zeros(int(be.lib().lagvae_lstm_workspace_bytes(nh, Bd)), dtype=torch.uint8, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
dbg = torch.zeros(2 * (Tn + 1) * 8, dtype=torch.int64, device=dev)
drop = be.Dropout()
out = []


def summarize(name, steps):
    d2 = dbg.view(-1, 8)[Tn + 1:Tn + 1 + steps].cpu().double()
    d = dbg.view(-1, 8)[:steps].cpu().double()
    t0 = d[:, 0]
    per_step = (t0[1:] - t0[:-1])
    sel = slice(5, steps - 1)
    names = ["first stage landed", "all MMAs issued", "accumulators complete", "partials in shared memory",
             "cluster barrier passed", "cell done, stores issued", "proxy fence done"]
    cols = [d[:, i] - d[:, 0] for i in (1, 2, 3, 5, 6, 7, 4)]
    out.append("%s: Bd=%d nh=%d  mean step = %.0f cycles" % (name, Bd, nh, float(per_step[sel].mean())))
    for n, c in zip(names, cols):
        if float(d[sel, 5].abs().max()) == 0.0 and n in names[3:6]:
            continue                      # v1 kernels do not record the cluster phases
        out.append("   %-28s +%7.0f cycles (mean offset from step start)" % (n, float(c[sel].mean())))
    out.append("   %-28s +%7.0f cycles" % ("next step start", float(per_step[sel].mean())))
    if float(d2.abs().max()) > 0:
        for i, n in enumerate(["copy: epilogue at go-wait", "copy: go seen", "copy: all stages issued", "copy: all stages landed"]):
            out.append("   %-32s %+7.0f cycles (vs step start)" % (n, float((d2[:, i] - d[:, 0])[sel].mean())))


for rep in range(2):
    gates = pre.clone()
    c_all = torch.zeros(Tn * Bd, nh, device=dev)
    h_all = torch.zeros(Tn * Bd, nh, device=dev)
    be.lib().lagvae_debug_trace_buffer(be.ptr(dbg), dbg.numel())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    be.check(be.lib().lagvae_lstm_forward(1, nh, Tn, Bd, be.ptr(w_hh), None, None, be.ptr(gates), be.ptr(c_all), be.ptr(h_all),
                                          None, C.byref(drop), be.ptr(ws), ws.numel(), st))
    e1.record()
    torch.cuda.synchronize()
    if rep == 1:
        out.append("forward kernel: %.3f ms for %d steps" % (e0.elapsed_time(e1), Tn))
        summarize("forward", Tn)
    dh_ext = torch.randn(Tn * Bd, nh, generator=g, device=dev) * 0.01
    dc, dhr, dg = torch.zeros(Bd, nh, device=dev), torch.zeros(Bd, nh, device=dev), torch.zeros(Tn * Bd, 4 * nh, device=dev)
    dbg.zero_()
    e0.record()
    be.check(be.lib().lagvae_lstm_backward(1, nh, Tn, Bd, be.ptr(w_hh), None, be.ptr(gates), be.ptr(c_all), be.ptr(dh_ext), None,
                                           C.byref(drop), be.ptr(dc), be.ptr(dhr), be.ptr(dg), 1, be.ptr(ws), ws.numel(), st))
    e1.record()
    torch.cuda.synchronize()
    if rep == 1:
        out.append("backward kernel: %.3f ms for %d steps" % (e0.elapsed_time(e1), Tn + 1))
        summarize("backward", Tn + 1)
be.lib().lagvae_debug_trace_buffer(None, 0)
txt = "\n".join(out)
print(txt)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "lstm_trace_bd%d.txt" % Bd), "w").write(txt + "\n")
