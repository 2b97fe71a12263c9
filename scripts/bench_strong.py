#!/usr/bin/env python
"""BASELINE.json configs[2]: Yelp text LSTM-VAE (V=19997, T=100 — SURVEY §8 d2(3)), aggressive inner step, kl_weight 1.0,
GLOBAL batch 32 sharded over 1 -> 8 B200s (per-rank 32 / 16 / 8 / 4 sentences), one NCCL all-reduce of the flat gradient bucket
per inner step = STRONG scaling.  Two legs, as in bench.py:
  value  fused kernels per rank (lagvae.dp.dp_inner_step: shard-local fwd+bwd, bucket all-reduce with the decoder part overlapped,
         clip + SGD), inputs resident in HBM, CUDA events, max over ranks;
  e2e    the drop-in `modules.VAE` driven by the statement sequence of text.py:373-387 SPMD — every rank gets the full 32-sentence
         batch from pinned host memory, `VAE.loss` shards it inside (lagvae.dp.ShardedTextLoss).
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/bench_strong.py --steps K --warmup W
"""
import argparse
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (os.path.join(ROOT, "vae-lagging-encoder_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, _p)
import numpy as np
import torch
import torch.distributed as dist

CFG = dict(V=19997, ni=512, nh=1024, nz=32, B=32, T=100)
KLW = 1.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    import lagvae
    import lagging_oracle as O
    import modules
    from lagvae.dp import EngineBackend, dp_inner_step
    rank, world, lr_ = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", lr_)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    c = CFG
    B, T, V = c["B"], c["T"], c["V"]
    p = O.init_text_params(V, c["ni"], c["nh"], c["nz"], seed=0)
    params = [p[k].to(dev).contiguous() for k in O.ALL_KEYS]
    eng = lagvae.TextEngine(V, c["ni"], c["nh"], c["nz"], dev)
    pool = [O.make_token_batch(B, T, V, seed=1234 + i).to(dev) for i in range(32)]       # identical on every rank
    rng = np.random.RandomState(783435)
    picks = [int(rng.randint(0, 32)) for _ in range(a.warmup + 3 * a.steps + 8)]
    gw = eng.grad_workspace()
    gen = torch.Generator(device=dev).manual_seed(783435 + rank)
    ctr = [0]

    def drop():
        ctr[0] += 1
        return lagvae.DropoutSpec(2, 0.5, 0.5, None, None, 783435 * 1000003 + ctr[0] * 7919 + rank)
    backend = EngineBackend(eng, KLW, lambda b: torch.empty(b, 1, c["nz"], device=dev).normal_(generator=gen), drop)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    step_fused = lambda i: dp_inner_step(backend, params, pool[picks[i]], gw)             # shards the 32 rows by rank
    for i in range(a.warmup):
        step_fused(i)
    reps = [timed(lambda i, r=r: step_fused(a.warmup + r * a.steps + i), a.steps) for r in range(3)]
    ms_step = float(np.median(reps)) / a.steps

    # e2e: module API, SPMD
    class Vocab(dict):
        def __len__(self):
            return V

        def id2word(self, i):
            return str(i)
    ns_ = types.SimpleNamespace(ni=c["ni"], enc_nh=c["nh"], dec_nh=c["nh"], nz=c["nz"], dec_dropout_in=0.5, dec_dropout_out=0.5, device=dev)
    mi_ = lambda t: torch.nn.init.uniform_(t, -0.01, 0.01)
    ei_ = lambda t: torch.nn.init.uniform_(t, -0.1, 0.1)
    torch.manual_seed(783435)
    vae = modules.VAE(modules.LSTMEncoder(ns_, V, mi_, ei_), modules.LSTMDecoder(ns_, Vocab(), mi_, ei_), ns_).to(dev).train()
    enc_opt = torch.optim.SGD(vae.encoder.parameters(), lr=1.0, momentum=0)
    dec_opt = torch.optim.SGD(vae.decoder.parameters(), lr=1.0, momentum=0)
    host_pool = [t.cpu().pin_memory() for t in pool[:16]]
    xdev = torch.empty(B, T, dtype=torch.int64, device=dev)
    allp = list(vae.parameters())

    def step_api(i):
        xdev.copy_(host_pool[picks[i] % 16], non_blocking=True)
        enc_opt.zero_grad()
        dec_opt.zero_grad()
        loss, _, _ = vae.loss(xdev, KLW, nsamples=1)
        s = loss.sum().item()
        loss.mean(dim=-1).backward()
        torch.nn.utils.clip_grad_norm_(allp, 5.0)
        enc_opt.step()
        return s
    for i in range(2):
        step_api(i)
    n_e2e = max(3, min(a.steps, 10))
    ems = timed(lambda i: step_api(2 + i), n_e2e)
    if rank == 0:
        line = {"metric": "aggressive inner-loop encoder steps/sec (Yelp LSTM-VAE, GLOBAL batch32 seq100)", "value": 1e3 / ms_step,
                "unit": "steps/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "data": "synthetic", "repeats_ms": reps,
                "dtype": "bf16x3-split operands, f32 accumulate/state (decoder weight gradients one bf16 pass in `value`)",
                "config": {"workload": "configs[2]: Yelp LSTM-VAE aggressive inner step, GLOBAL B=32 T=100 V=19997 ni=512 nh=1024 nz=32, "
                                       "kl_weight 1.0, per-rank rows %d" % (-(-B // world)), "global_batch": B, "parallelism": "dp%d" % world},
                "e2e": {"value": n_e2e / (ems / 1e3), "unit": "steps/s", "h2d_bytes_per_step": B * T * 8, "d2h_bytes_per_step": 4, "steps": n_e2e,
                        "api": "modules.VAE.loss (sharded inside) -> backward -> clip_grad_norm_ -> SGD.step, SPMD"},
                "lstm_variant": lagvae.lstm_variant()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
