#!/usr/bin/env python
"""bench.py — aggressive inner-loop encoder steps/s (BASELINE.json metric; Yahoo LSTM-VAE, batch 32, seq 200).

One *step* = one iteration of the reference loop text.py:371-391: zero_grad, VAE.loss forward, Σloss,
backward of mean(loss), clip_grad_norm_(all 13 tensors, 5.0), SGD step on the 6 encoder tensors, next batch.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

ours:       `value`  fused lagvae_text_inner_step, inputs resident in HBM, CUDA events, max over ranks
            `e2e`    the drop-in `modules.VAE` API exactly as text.py:373-387 drives it, token ids in PINNED
                     HOST memory copied H2D every step, Σloss read back D2H every step
reference:  the reference's own modules (a staged copy: $VAE_REF_PATH / baseline/_ref; else oracle.FastPort, the same
            torch layer types) on the host cores, FULL 32-sentence steps, the sample bounded in the number of steps.
N > 1:      one process per GPU (torchrun), batch-sharded data parallel: every rank runs the same step on
            its own 32-sentence shard, ONE NCCL all-reduce of the flat 53.8 M-float gradient per step, then
            clip + SGD on the averaged gradient.  Weak scaling: global batch = 32 N.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "vae-lagging-encoder_b200"), os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np
import torch

METRIC = "aggressive inner-loop encoder steps/sec (Yahoo LSTM-VAE, batch32 seq200)"
CFG = dict(V=20001, ni=512, nh=1024, nz=32, B=32, T=200)   # BASELINE.json configs[1], SURVEY §8(d2)
KL_WEIGHT = 0.1                                           # --kl_start 0.1 (text.py:337-338)
POOL = 64


def flops_step(B, T, V, ni, nh, nz):
    """F_step = 3 * F_fwd."""
    Td = T - 1
    f = 2 * B * (T * ni * 4 * nh + T * nh * 4 * nh + nh * 2 * nz) + \
        2 * B * (Td * (ni + nz) * 4 * nh + Td * nh * 4 * nh + nz * nh + Td * nh * V)
    return 3 * f


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx, self.rows, self.stop_flag = gpu_index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------------------------------------------
def _reference_stepper(device):
    """The reference's OWN modules ($VAE_REF_PATH -> baseline/_ref -> /root/reference) driven by the statement sequence
    of text.py:373-387; the oracle's port (the same torch layer types, oracle.FastPort) only when no copy of the reference
    is present.  Returns (stepper, kind) with kind "reference" | "port"."""
    import lagging_oracle as O
    import reference_loader as R
    c = CFG
    ref = R.load_reference_modules()
    if ref is not None:
        vae = R.build_reference_vae(ref, c["V"], c["ni"], c["nh"], c["nz"], device)
        return R.ReferenceStepper(vae), "reference"
    torch.manual_seed(0)
    return O.FastPort(c["V"], c["ni"], c["nh"], c["nz"]).to(device).train(), "port"


def _host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def _calibrate_threads(m, budget_s=30.0):
    """oneDNN's LSTM with 200 dependent time steps can scale NEGATIVELY with the thread count on many-core hosts
    (measured in round 1: 128 threads 30x slower than 8), so "all the host threads it can use" is found by timing an
    8-sentence step at every core and at a few smaller counts and keeping the fastest.  Full-batch steps are then timed
    at that count; nothing is extrapolated."""
    import lagging_oracle as O
    c = CFG
    ncpu = _host_threads()
    x8 = O.make_token_batch(8, c["T"], c["V"], seed=77)
    cands = sorted({min(ncpu, t) for t in (ncpu, ncpu // 2, 32, 16)} - {0}, reverse=True)
    best_t, best_th, t_start = None, ncpu, time.perf_counter()
    for th in cands:
        torch.set_num_threads(th)
        m.inner_step(x8, KL_WEIGHT)                   # warm (allocator, oneDNN primitive cache)
        t0 = time.perf_counter()
        m.inner_step(x8, KL_WEIGHT)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best_t, best_th = dt, th
        if time.perf_counter() - t_start > budget_s:
            break
    torch.set_num_threads(best_th)
    return best_th, ncpu


def _cpu_reference_rate(n_steps, n_warm, budget_s):
    """Time FULL 32-sentence steps of the reference path on the host cores.  The sample is bounded in the number of steps
    (as many of the requested K as fit `budget_s`, at least 2), never in the batch: a full step is the unit of the metric."""
    import lagging_oracle as O
    c = CFG
    m, kind = _reference_stepper("cpu")
    th, ncpu = _calibrate_threads(m)
    xs = [O.make_token_batch(c["B"], c["T"], c["V"], seed=1234 + i) for i in range(4)]
    t0 = time.perf_counter()
    m.inner_step(xs[0], KL_WEIGHT)                   # warm-up step, also the estimate of a step's cost
    t1 = time.perf_counter() - t0
    n_w = max(0, min(n_warm, 3) - 1)
    n_t = int(max(2, min(n_steps, (budget_s - t1 * (1 + n_w)) / max(t1, 1e-3))))
    for i in range(n_w):
        m.inner_step(xs[(1 + i) % 4], KL_WEIGHT)
    t0 = time.perf_counter()
    for i in range(n_t):
        m.inner_step(xs[i % 4], KL_WEIGHT)
    dt = time.perf_counter() - t0
    rate = n_t / dt
    what = "the reference's own modules (VAE.loss -> backward -> clip_grad_norm_ -> SGD.step, text.py:373-387)" if kind == "reference" \
        else "oracle.FastPort (the reference's torch layer types; no copy of the reference on this box)"
    sample = ("%d timed full steps (32 sentences x T=%d, train mode) after %d warm-up steps of %s on the host CPU; "
              "%d of %d host threads (fastest of the calibrated counts); no extrapolation" % (n_t, c["T"], 1 + n_w, what, th, ncpu))
    return rate, th, kind, sample


def run_reference(args, rank, world):
    """Reference arm: the reference path on the host cores (rank 0 only), bounded to a few minutes whatever K and W."""
    if rank != 0:
        return
    rate, th, kind, sample = _cpu_reference_rate(args.steps, args.warmup, budget_s=240.0)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / rate, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: Yahoo LSTM-VAE aggressive inner step, B=32 T=200 V=20001 (reference path, host cores)"},
            "cpu_baseline": {"value": rate, "unit": "steps/s", "cores": th, "kind": kind, "sample": sample},
            "e2e": {"value": rate, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_quick():
    """~10-30 s of CPU work: a bounded number of FULL steps of the reference path (rank 0, N=1 only)."""
    rate, th, kind, sample = _cpu_reference_rate(3, 1, budget_s=25.0)
    return {"value": rate, "unit": "steps/s", "cores": th, "kind": kind, "sample": sample}


def reference_same_gpu(steps=20):
    """The reference's own modules on THIS GPU through torch's stock cuDNN/cuBLAS path (cudnn.deterministic as
    text.py:102, default TF32 flags recorded) — the denominator of north_star's ">= 10x the reference single-GPU PyTorch
    step time".  >= 20 timed steps, CUDA events, host inputs resident on the device as in the reference script."""
    import lagging_oracle as O
    c = CFG
    torch.backends.cudnn.deterministic = True         # text.py:102
    m, kind = _reference_stepper("cuda")
    xs = [O.make_token_batch(c["B"], c["T"], c["V"], seed=99 + i).cuda() for i in range(4)]
    for i in range(3):
        m.inner_step(xs[i % 4], KL_WEIGHT)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        m.inner_step(xs[i % 4], KL_WEIGHT)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del m
    torch.cuda.empty_cache()
    return {"value": 1e3 / ms, "unit": "steps/s", "ms_per_step": ms, "steps": steps, "kind": kind,
            "what": ("the reference's own modules" if kind == "reference" else "oracle.FastPort") +
                    " (nn.LSTM / nn.Linear / CrossEntropyLoss -> cuDNN / cuBLAS fp32, text.py:373-387 sequence) on this GPU",
            "tf32_matmul": bool(torch.backends.cuda.matmul.allow_tf32), "tf32_cudnn": bool(torch.backends.cudnn.allow_tf32)}


def time_vocab_gemm(eng, reps=5):
    """Dominant tensor kernel timed alone with CUDA events on the launching stream: the vocabulary
    projection [B*(T-1), nh] x [nh, V] (dec_lstm.py:109), split-bf16 3-pass."""
    import ctypes as C
    import lagvae._backend as be
    c = CFG
    M, N, K = c["B"] * (c["T"] - 1), c["V"], c["nh"]
    A = torch.randn(M, K, device="cuda")
    Bm = torch.randn(N, K, device="cuda") * 0.03
    ah, al = A.to(torch.bfloat16), (A - A.to(torch.bfloat16).float()).to(torch.bfloat16)
    bh, bl = Bm.to(torch.bfloat16), (Bm - Bm.to(torch.bfloat16).float()).to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    call = lambda: be.check(be.lib().lagvae_gemm_tc(be.ptr(ah), be.ptr(al), K, 0, be.ptr(bh), be.ptr(bl), K, 0, be.ptr(out), N,
                                                    M, N, K, 3, 1.0, 0.0, None, None, 0, None, st))
    for _ in range(2):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return 2.0 * M * N * K, ms


def time_lstm_kernels(reps=3):
    """The dominant kernels of the step: the persistent tcgen05 LSTM recurrences (one launch = all T time steps),
    timed alone with CUDA events on the launching stream at the Yahoo recurrence shape."""
    import ctypes as C
    import lagvae._backend as be
    c = CFG
    nh, Bd, Tn = c["nh"], c["B"], c["T"]
    w_hh = (torch.rand(4 * nh, nh, device="cuda") * 2 - 1) * (1.0 / nh ** 0.5)
    pre = torch.randn(Tn * Bd, 4 * nh, device="cuda")
    ws = torch.zeros(int(be.lib().lagvae_lstm_workspace_bytes(nh, Bd)), dtype=torch.uint8, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    drop = be.Dropout()
    gates = pre.clone()
    c_all, h_all = torch.zeros(Tn * Bd, nh, device="cuda"), torch.zeros(Tn * Bd, nh, device="cuda")
    dh = torch.randn(Tn * Bd, nh, device="cuda") * 0.01
    dc, dhr, dg = torch.zeros(Bd, nh, device="cuda"), torch.zeros(Bd, nh, device="cuda"), torch.zeros(Tn * Bd, 4 * nh, device="cuda")
    fwd = lambda: be.check(be.lib().lagvae_lstm_forward(1, nh, Tn, Bd, be.ptr(w_hh), None, None, be.ptr(gates), be.ptr(c_all),
                                                        be.ptr(h_all), None, C.byref(drop), be.ptr(ws), ws.numel(), st))
    bwd = lambda: be.check(be.lib().lagvae_lstm_backward(1, nh, Tn, Bd, be.ptr(w_hh), None, be.ptr(gates), be.ptr(c_all), be.ptr(dh),
                                                         None, C.byref(drop), be.ptr(dc), be.ptr(dhr), be.ptr(dg), 1, be.ptr(ws), ws.numel(), st))
    out = {}
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        gates.copy_(pre) if name == "fwd" else None
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            if name == "fwd":
                gates.copy_(pre)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        out[name] = float(np.median(ms))
    return out, 2.0 * Bd * nh * 4 * nh * Tn


def run_ours(args, rank, world, local_rank):
    import lagvae
    import lagging_oracle as O
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    c = CFG
    B, T, V = c["B"], c["T"], c["V"]
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    # parameters: the reference initialisers (text.py:265-266), identical on every rank
    p = O.init_text_params(V, c["ni"], c["nh"], c["nz"], seed=0)
    params = [p[k].to(dev).contiguous() for k in O.ALL_KEYS]
    eng = lagvae.TextEngine(V, c["ni"], c["nh"], c["nz"], dev)
    # pool of 64 distinct device-resident batches (per-rank shard: different sentences on each rank)
    pool = [O.make_token_batch(B, T, V, seed=1234 + 1000 * rank + i).to(dev) for i in range(POOL)]
    rng = np.random.RandomState(783435)               # text.py:73 seed; picks identical on all ranks
    picks = [int(rng.randint(0, POOL)) for _ in range(args.warmup + args.steps + 8)]
    gw = eng.grad_workspace()
    grads = eng.split_grads(gw)
    out_loss = torch.empty(B, device=dev)
    sc = torch.empty(4, device=dev)
    gen = torch.Generator(device=dev).manual_seed(783435 + rank)
    seed_ctr = [0]

    def drop():
        seed_ctr[0] += 1
        return lagvae.DropoutSpec(2, 0.5, 0.5, None, None, 783435 * 1000003 + seed_ctr[0] * 7919 + rank)

    from lagvae.dp import EngineBackend, dp_inner_step
    backend = EngineBackend(eng, KL_WEIGHT, lambda b: torch.empty(b, 1, c["nz"], device=dev).normal_(generator=gen), drop)

    def step_fused(i):
        x = pool[picks[i]]
        eps = torch.empty(B, 1, c["nz"], device=dev).normal_(generator=gen) if world == 1 else None   # encoder.py:77
        if world == 1:
            eng.inner_step(params, x, eps, KL_WEIGHT, drop(), gw, out_loss, sc)   # text.py:373-387 fused
        else:
            # tested host logic (tests/test_dp_gloo.py): shard-local fwd+bwd with 1/(B*N) upstream gradient,
            # ONE all-reduce of the flat gradient, clip + SGD on the averaged gradient
            dp_inner_step(backend, params, x, gw, presharded=True, global_rows=B * world, read_loss=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: device-resident inputs ----------------
    for i in range(args.warmup):
        step_fused(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # three repeats of EXACTLY K steps, each bracketed by barrier + synchronize on both sides and timed with CUDA events on
    # the launching stream (max over ranks); `value` is the median repeat, all three are reported
    rep_ms, launches, n_timed_samples = [], 0, 0
    for rep in range(3):
        barrier()
        l0 = lagvae.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            step_fused(args.warmup + (rep * args.steps + i) % (len(picks) - args.warmup))
        e1.record()
        barrier()
        launches = lagvae.launch_count() - l0
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        rep_ms.append(float(ms))
        if rep == 0:
            n_timed_samples = len(sampler.rows)        # clock samples taken inside the first timed region proper
    ms_total = float(np.median(rep_ms))
    ms_per_step = ms_total / args.steps
    value = world * args.steps / (ms_total / 1e3)     # 32-sentence encoder steps per second, all ranks
    variant = lagvae.lstm_variant()
    if c["nh"] >= 256 and not (variant["forward"].startswith("v") and variant["backward"].startswith("v")):
        raise RuntimeError("bench: the persistent recurrence kernels did not run (%r)" % (variant,))

    # ---------------- e2e: drop-in modules API, host token ids, per-step readback ----------------
    e2e = None
    if not args.no_e2e:
        import types
        import modules

        class Vocab(dict):
            def __len__(self):
                return V

            def id2word(self, i):
                return str(i)
        a = types.SimpleNamespace(ni=c["ni"], enc_nh=c["nh"], dec_nh=c["nh"], nz=c["nz"], dec_dropout_in=0.5,
                                  dec_dropout_out=0.5, device=dev)
        mi_ = lambda t: torch.nn.init.uniform_(t, -0.01, 0.01)
        ei_ = lambda t: torch.nn.init.uniform_(t, -0.1, 0.1)
        torch.manual_seed(783435)
        vae = modules.VAE(modules.LSTMEncoder(a, V, mi_, ei_), modules.LSTMDecoder(a, Vocab(), mi_, ei_), a).to(dev)
        vae.train()
        enc_opt = torch.optim.SGD(vae.encoder.parameters(), lr=1.0, momentum=0)
        dec_opt = torch.optim.SGD(vae.decoder.parameters(), lr=1.0, momentum=0)
        # N > 1: the SPMD contract of SURVEY §8 b3 — every rank runs the SAME statements on the SAME global batch
        # (32 sentences per GPU: rows [32 r, 32 r + 32) are rank r's shard); `VAE.loss` shards it by rank internally, returns
        # the full vectors and all-reduces the flat gradient bucket in its backward (lagvae_allreduce_bucket)
        if world == 1:
            host_pool = [t.cpu().pin_memory() for t in pool[:16]]
        else:
            host_pool = [torch.cat([O.make_token_batch(B, T, V, seed=1234 + 1000 * r + i) for r in range(world)]).pin_memory()
                         for i in range(16)]
        xdev = torch.empty(B * world, T, dtype=torch.int64, device=dev)
        allp = list(vae.parameters())

        def step_api(i):
            xdev.copy_(host_pool[picks[i] % 16], non_blocking=True)               # H2D of this step's input
            enc_opt.zero_grad()
            dec_opt.zero_grad()
            loss, loss_rc, loss_kl = vae.loss(xdev, KL_WEIGHT, nsamples=1)        # text.py:379 (sharded over the ranks inside)
            s = loss.sum().item()                                                 # text.py:381 (D2H sync)
            loss = loss.mean(dim=-1)
            loss.backward()                                                       # text.py:384 (+ the bucket all-reduce)
            torch.nn.utils.clip_grad_norm_(allp, 5.0)                             # text.py:385
            enc_opt.step()                                                        # text.py:387
            return s

        n_e2e = max(3, min(args.steps, 30))
        for i in range(2):
            step_api(i)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(n_e2e):
            step_api(2 + i)
        f1.record()
        barrier()
        ems = torch.tensor([f0.elapsed_time(f1)], device=dev)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        e2e = {"value": world * n_e2e / (float(ems) / 1e3), "unit": "steps/s", "h2d_bytes_per_step": B * world * T * 8,
               "d2h_bytes_per_step": 4, "steps": n_e2e,
               "api": "modules.VAE.loss -> backward -> clip_grad_norm_ -> SGD.step (text.py:373-387 sequence)" +
                      ("; SPMD: every rank is handed the global batch of %d sentences, VAE.loss shards it by rank and all-reduces "
                       "the flat gradient bucket (lagvae_allreduce_bucket) in its backward" % (B * world) if world > 1 else "")}
        del vae, enc_opt, dec_opt
        torch.cuda.empty_cache()

    # ---------------- graph: a 15-step window of the inner loop as ONE CUDA graph ----------------
    # text.py:371-400 reads only the ACCUMULATED Σloss, once per 15 iterations (text.py:389-398), so a window of 15 fused
    # inner steps is one graph launch: H2D of the window's 15 batches from pinned host memory, replay, D2H of one float.
    # Philox masks are fresh per step (device seed word advanced inside the graph), the decoder weights are re-split once
    # per replay (lagvae/graph.py); tests/test_gpu_text_graph.py checks the replay against the eager steps.
    graph_leg = None
    if world == 1 and not args.no_e2e:
        try:
            from lagvae import graph as G
            Wn = 15
            word = G.philox_word(dev)
            host_x = [torch.stack([pool[(j * Wn + i) % POOL] for i in range(Wn)]).cpu().pin_memory() for j in range(4)]
            xs_static = torch.empty(Wn, B, T, dtype=torch.int64, device=dev)
            eps_static = torch.empty(Wn, B, 1, c["nz"], device=dev)
            out_l = torch.empty(Wn, B, device=dev)
            out_s = torch.empty(Wn, 4, device=dev)

            def window(xs):
                eps_static.normal_()                               # graph-safe generator: fresh eps per replay
                for i in range(Wn):
                    G.bump_philox_word(word)
                    eng.inner_step(params, xs[i], eps_static[i], KL_WEIGHT,
                                   lagvae.DropoutSpec(2, 0.5, 0.5, None, None, 783435 * 1000003 + rank, word), gw, out_l[i], out_s[i])
                return out_s[:, 0].sum()                           # burn_cur_loss of the window (text.py:389)

            gs = lagvae.GraphedStep(window, {"xs": xs_static}, warmup=1)
            nwin = max(2, min(6, args.steps // Wn))
            for j in range(2):
                float(gs(xs=host_x[j % 4]))
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for j in range(nwin):
                burn = float(gs(xs=host_x[j % 4]))                 # H2D (pinned) -> replay -> D2H of the window's Σloss
            g1.record()
            barrier()
            gms_total = g0.elapsed_time(g1)
            graph_leg = {"value": nwin * Wn / (gms_total / 1e3), "unit": "steps/s", "steps_per_graph": Wn, "windows": nwin,
                         "ms_per_step": gms_total / (nwin * Wn), "h2d_bytes_per_step": B * T * 8, "d2h_bytes_per_step": 4.0 / Wn,
                         "last_window_loss_sum": burn,
                         "api": "lagvae.GraphedStep(15 x TextEngine.inner_step = lagvae_text_inner_step): host token ids in pinned "
                                "memory, one graph launch and one 4-byte readback per 15-step window (text.py:389-398 reads only "
                                "the window sum)"}
            del gs
        except Exception as ex:
            graph_leg = {"error": repr(ex)[:300]}

    if rank != 0:
        sampler.stop_flag = True
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    F = flops_step(B, T, V, c["ni"], c["nh"], c["nz"])
    gflop, gms = time_vocab_gemm(eng)
    lstm_ms, lstm_flop = time_lstm_kernels()
    # the sampler ran through the timed region AND the e2e / kernel-timing legs that follow it (all under the same load)
    sampler.stop_flag = True
    sampler.join(timeout=3)
    clocks = sampler.summary()
    clocks["samples_in_timed_region"] = n_timed_samples
    peak = peaks["bf16_tflops"]
    roof = {"bound": "tensor", "kernel": "k_lstm_v2<backward> persistent tcgen05 LSTM recurrence (dominant kernel: 4 LSTM launches = "
            "%.0f%% of the step; B=32 nh=1024, 200 dependent time steps per launch)" % (100.0 * 2 * (lstm_ms["fwd"] + lstm_ms["bwd"]) / ms_per_step),
            "achieved": lstm_flop / (lstm_ms["bwd"] * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
            "frac": lstm_flop / (lstm_ms["bwd"] * 1e-3) / 1e12 / peak, "traffic": None,
            "peak_source": peak_src + " bf16 burst (kernel timed alone with CUDA events on the launching stream)",
            "ms_per_launch": lstm_ms["bwd"], "us_per_time_step": 1e3 * lstm_ms["bwd"] / (c["T"] + 1),
            "note": "sequential-dependency latency bound, not a throughput kernel: each of the 200 steps is all-gather -> "
                    "64 tcgen05.mma -> cluster DSMEM reduce -> cell -> grid barrier (profiles/README.md); algorithmic FLOPs "
                    "2*B*nh*4nh*T counted once (3 split-bf16 passes issued)",
            "forward": {"ms_per_launch": lstm_ms["fwd"], "achieved": lstm_flop / (lstm_ms["fwd"] * 1e-3) / 1e12},
            "gemm": {"kernel": "k_gemm_tc<K-major,K-major> vocab projection 6368x20001x1024 (split-bf16, 3 MMA passes)",
                     "bound": "tensor", "achieved": gflop / (gms * 1e-3) / 1e12, "peak": peak,
                     "frac": gflop / (gms * 1e-3) / 1e12 / peak, "ms_per_launch": gms,
                     "mma_rate_frac": 3 * gflop / (gms * 1e-3) / 1e12 / peak},
            "step_tensor_frac": F / (ms_per_step * 1e-3) / 1e12 / peaks.get("bf16_tflops_sustained", peak)}
    line = {"metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16x3-split operands (hi*hi + hi*lo + lo*hi), f32 accumulate/state; in `value` and `graph` the four decoder "
                     "weight-gradient GEMMs (dW_pred, decoder dW_ih, dW_hh and the dX that feeds only the decoder embedding "
                     "gradient: never applied in the inner loop, they only enter the clip norm) run ONE bf16 pass; `e2e` runs "
                     "all GEMMs 3-pass",
            "data": "synthetic", "repeats_ms": rep_ms, "lstm_variant": variant,
            "config": {"workload": "configs[1]: Yahoo LSTM-VAE aggressive inner step (text.py:371-391), B=32/GPU T=200 V=20001 ni=512 nh=1024 nz=32, "
                                   "train-mode dropout 0.5/0.5, kl_weight 0.1, SGD lr 1.0, clip 5.0",
                       "global_batch": B * world, "parallelism": "dp%d" % world,
                       "l2": "per-step working set (509 MB logits + 420 MB gate stashes) exceeds the 126 MB L2; no explicit flush",
                       "pool": "%d device-resident batches, np.random.seed(783435) picks" % POOL},
            "clocks": clocks, "e2e": e2e, "graph": graph_leg, "gpu_launches": int(launches), "roofline": roof,
            "flops_per_step": F, "algorithmic_tflops": F / (ms_per_step * 1e-3) / 1e12}
    if world == 1 and not args.no_cpu:
        try:
            line["reference_same_gpu"] = reference_same_gpu()
            line["speedup_vs_reference_same_gpu"] = {"value": value / line["reference_same_gpu"]["value"],
                                                     "e2e": (e2e["value"] / line["reference_same_gpu"]["value"]) if e2e else None,
                                                     "target": 10.0}
        except Exception as ex:
            line["reference_same_gpu"] = {"error": repr(ex)[:200]}
        line["cpu_baseline"] = cpu_baseline_quick()
    if world == 1 and not args.no_image:
        # informational: configs[3] (Omniglot ResNet encoder + PixelCNN decoder, batch 64) through the same drop-in API
        try:
            sys.path.insert(0, os.path.join(ROOT, "scripts"))
            import bench_image
            line["image"] = bench_image.run(B=64, steps=10, warm=3, dev=dev)
        except Exception as ex:
            line["image"] = {"error": repr(ex)[:300]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", dest="no_cpu", action="store_true", help="skip the CPU / torch-GPU side baselines")
    ap.add_argument("--no-e2e", dest="no_e2e", action="store_true", help="skip the module-API leg (profiling runs only)")
    ap.add_argument("--no-image", dest="no_image", action="store_true", help="skip the informational image-path (configs[3]) numbers")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        print(json.dumps({"error": "launch with torchrun for --gpus > 1"}))
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
