"""N>1 host logic on CPU: world_size-2 gloo run of lagvae.dp.dp_inner_step with an oracle back-end must equal
the single-process step on the full batch (exact reference semantics under batch sharding)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup_path():
    for p in (os.path.join(ROOT, "vae-lagging-encoder_b200"), os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)


class OracleBackend:
    """CPU stand-in for EngineBackend: gradients from the oracle (test-only)."""

    def __init__(self, eps_global, klw, lo, decoder_offset=None):
        self.eps, self.klw, self.lo = eps_global, klw, lo
        if decoder_offset is not None:          # two-bucket collective ([decoder | encoder]) as EngineBackend exposes it
            self.decoder_offset = decoder_offset

    def forward_backward(self, params, x, g_scale, flat):
        import lagging_oracle as O
        leaves = {k: p.detach().clone().requires_grad_(True) for k, p in zip(O.ALL_KEYS, params)}
        eps = self.eps[self.lo:self.lo + x.shape[0]]
        loss, _, _ = O.vae_loss(leaves, x, self.klw, eps)
        (loss * g_scale).sum().backward()
        off = 0
        for k in O.ALL_KEYS:
            g = leaves[k].grad if leaves[k].grad is not None else torch.zeros_like(leaves[k])
            if k == "decoder.embed.weight":
                g[-1].zero_()
            flat[off:off + g.numel()] = g.reshape(-1)
            off += g.numel()
        return loss.detach()

    def clip_sgd(self, params, flat, max_norm, lr):
        import lagging_oracle as O
        norm = float(flat.double().norm())
        coef = O.clip_coef(norm, max_norm)
        off = 0
        for i, p in enumerate(params):
            n = p.numel()
            if i < 6:
                p -= lr * coef * flat[off:off + n].view_as(p)
            off += n
        return norm


def _worker(rank, world, port, B, out_q, two_buckets=False):
    _setup_path()
    import lagging_oracle as O
    from lagvae.dp import dp_inner_step, shard_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    V, ni, nh, nz, T = 60, 6, 8, 2, 5
    p = O.scale_trained_like(O.init_text_params(V, ni, nh, nz, seed=4))
    params = [p[k] for k in O.ALL_KEYS]
    x = O.make_token_batch(B, T, V)
    eps = torch.randn(B, 1, nz, generator=torch.Generator().manual_seed(9))
    flat = torch.zeros(sum(q.numel() for q in params))
    lo, hi = shard_bounds(B, rank, world)
    dec_off = sum(q.numel() for q in params[:6]) if two_buckets else None
    backend = OracleBackend(eps, 0.5, lo, dec_off)
    if two_buckets == "single":       # the default of the product back-end: bucket split known, ONE all-reduce after the backward,
        backend.single_bucket = True  # and Σloss left on the device (read_loss=False)
        loss_sum, norm = dp_inner_step(backend, params, x, flat, max_norm=0.05, read_loss=False)
        assert isinstance(loss_sum, torch.Tensor) and loss_sum.shape == (1,)
        loss_sum = float(loss_sum)
    else:
        loss_sum, norm = dp_inner_step(backend, params, x, flat, max_norm=0.05)
    out_q.put((rank, loss_sum, norm, [q.numpy().copy() for q in params[:6]]))   # numpy: pickled by value (a torch tensor
    # travels as a shared-memory fd that dies with this process)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B,two_buckets", [(6, False), (5, False), (1, False), (5, True), (1, True), (5, "single"), (1, "single")])
def test_two_rank_gloo_equals_single_process(B, two_buckets):
    """even, ragged, and an empty shard on rank 1; single bucket and the [decoder | encoder] split of the overlap path"""
    _setup_path()
    import lagging_oracle as O
    world, port = 2, 29000 + os.getpid() % 2000 + B + (20 if two_buckets == "single" else (10 if two_buckets else 0))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q, two_buckets)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    res = [(a, b, c, [torch.from_numpy(e) for e in d]) for a, b, c, d in res]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    # single-process reference on the full batch
    V, ni, nh, nz, T = 60, 6, 8, 2, 5
    p = O.scale_trained_like(O.init_text_params(V, ni, nh, nz, seed=4))
    x = O.make_token_batch(B, T, V)
    eps = torch.randn(B, 1, nz, generator=torch.Generator().manual_seed(9))
    r = O.inner_step(p, x, 0.5, eps, max_norm=0.05, update=True)
    for rank, loss_sum, norm, enc in res:
        assert abs(loss_sum - r["loss_sum"]) <= 1e-5 * abs(r["loss_sum"])
        assert abs(norm - r["grad_norm"]) <= 1e-5 * r["grad_norm"]
        assert r["coef"] < 1.0          # the clip is active: the norm of the AVERAGED gradient matters
        for k, got in zip(O.ENC_KEYS, enc):
            assert float((got - p[k]).abs().max()) <= 1e-6 * max(1.0, float(p[k].abs().max())), k
    # replicas stay bit-identical without a parameter broadcast
    for a, b in zip(res[0][3], res[1][3]):
        assert torch.equal(a, b)


def test_shard_bounds_cover_and_partition():
    _setup_path()
    from lagvae.dp import shard_bounds
    for n in (0, 1, 5, 32, 33):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("B,chunk", [(6, 4), (5, 2), (4, 4)])
def test_micro_batch_accumulation_equals_full_batch(B, chunk):
    """accumulated_inner_step (batches beyond the persistent LSTM kernels' row limit run as micro-batches) against the
    single full-batch step of the oracle: same Σloss, same clip norm of the mean gradient, same encoder update."""
    _setup_path()
    import lagging_oracle as O
    from lagvae.dp import accumulated_inner_step
    V, ni, nh, nz, T = 60, 6, 8, 2, 5
    p = O.scale_trained_like(O.init_text_params(V, ni, nh, nz, seed=4))
    params = [p[k].clone() for k in O.ALL_KEYS]
    x = O.make_token_batch(B, T, V)
    eps = torch.randn(B, 1, nz, generator=torch.Generator().manual_seed(9))
    n = sum(q.numel() for q in params)
    flat, tmp = torch.zeros(n), torch.zeros(n)
    backend = OracleBackend(eps, 0.5, 0)
    loss_sum, norm = accumulated_inner_step(backend, params, x, flat, tmp, chunk, max_norm=0.05,
                                            on_chunk=lambda lo: setattr(backend, "lo", lo))
    r = O.inner_step(p, x, 0.5, eps, max_norm=0.05, update=True)
    assert abs(loss_sum - r["loss_sum"]) <= 1e-5 * abs(r["loss_sum"])
    assert abs(norm - r["grad_norm"]) <= 1e-5 * r["grad_norm"] and r["coef"] < 1.0
    for k, got in zip(O.ENC_KEYS, params[:6]):
        assert float((got - p[k]).abs().max()) <= 1e-6 * max(1.0, float(p[k].abs().max())), k


# ----------------------------------------------------------------------------------------------------------------------
# b3: VAE.loss sharded inside the boundary (lagvae.dp.ShardedTextLoss) — the statement sequence of text.py:379-387 run SPMD
class OracleEngine:
    """CPU stand-in for lagvae.TextEngine (test-only): same loss_forward / loss_backward contract, oracle arithmetic."""

    def __init__(self, two_buckets):
        self.generation = 0
        if two_buckets:
            self.decoder_offset = None      # set by the worker once the parameter sizes are known

    def loss_forward(self, params, x, eps, kl_weight, drop):
        import lagging_oracle as O
        self.generation += 1
        with torch.enable_grad():          # called from inside an autograd.Function.forward (grad mode off there)
            self._leaves = {k: p.detach().clone().requires_grad_(True) for k, p in zip(O.ALL_KEYS, params)}
            self._out = O.vae_loss(self._leaves, x, kl_weight, eps)
        return tuple(t.detach() for t in self._out)

    def loss_backward(self, params, x, g_loss, g_rec, g_kl, generation=None, grads_out=None):
        import lagging_oracle as O
        assert generation == self.generation
        with torch.enable_grad():
            obj = sum((o * g).sum() for o, g in zip(self._out, (g_loss, g_rec, g_kl)) if g is not None)
        obj.backward()
        for k, dst in zip(O.ALL_KEYS, grads_out):
            g = self._leaves[k].grad
            g = torch.zeros_like(self._leaves[k]) if g is None else g.clone()
            if k == "decoder.embed.weight":
                g[-1].zero_()
            dst.copy_(g)
        return grads_out


def _sharded_worker(rank, world, port, B, out_q, two_buckets):
    _setup_path()
    import lagging_oracle as O
    from lagvae.dp import ShardedTextLoss
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    V, ni, nh, nz, T = 60, 6, 8, 2, 5
    p = O.scale_trained_like(O.init_text_params(V, ni, nh, nz, seed=4))
    params = [p[k].clone().requires_grad_(True) for k in O.ALL_KEYS]
    x = O.make_token_batch(B, T, V)                       # SPMD: every rank holds the full batch and the same eps
    eps = torch.randn(B, 1, nz, generator=torch.Generator().manual_seed(9))
    eng = OracleEngine(two_buckets)
    if two_buckets:
        eng.decoder_offset = sum(q.numel() for q in params[:6])
    loss, rec, kl = ShardedTextLoss.apply(eng, dist.group.WORLD, x, 0.5, eps, lambda lo, hi: None, *params)
    s = loss.sum().item()                                 # text.py:381
    loss.mean(dim=-1).backward()                          # text.py:382-384
    norm = float(torch.nn.utils.clip_grad_norm_(params, 0.05))    # text.py:385 (threshold scaled to the toy gradients)
    torch.optim.SGD(params[:6], lr=1.0).step()            # text.py:387
    out_q.put((rank, s, norm, loss.detach().numpy().copy(), rec.detach().numpy().copy(), kl.detach().numpy().copy(),
               [q.detach().numpy().copy() for q in params[:6]], [q.grad.numpy().copy() for q in params]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B,two_buckets", [(6, False), (5, True), (1, True)])
def test_sharded_vae_loss_equals_single_process(B, two_buckets):
    """The driver's statement sequence over ShardedTextLoss on 2 gloo ranks (even, ragged, empty shard) == the single-process
    oracle step: full [B] loss vectors on every rank, gradient of the global mean, clip norm, encoder update."""
    _setup_path()
    import lagging_oracle as O
    world, port = 2, 31000 + os.getpid() % 2000 + B + (10 if two_buckets else 0)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, B, q, two_buckets)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    V, ni, nh, nz, T = 60, 6, 8, 2, 5
    p = O.scale_trained_like(O.init_text_params(V, ni, nh, nz, seed=4))
    x = O.make_token_batch(B, T, V)
    eps = torch.randn(B, 1, nz, generator=torch.Generator().manual_seed(9))
    r = O.inner_step(p, x, 0.5, eps, max_norm=0.05, update=True)
    for rank, s, norm, loss, rec, kl, enc, grads in res:
        assert abs(s - r["loss_sum"]) <= 1e-5 * abs(r["loss_sum"])
        assert np.allclose(loss, r["loss"].numpy(), rtol=1e-5, atol=1e-6) and loss.shape == (B,)
        assert np.allclose(rec, r["rec"].numpy(), rtol=1e-5, atol=1e-6) and np.allclose(kl, r["kl"].numpy(), rtol=1e-4, atol=1e-6)
        assert abs(norm - r["grad_norm"]) <= 1e-5 * r["grad_norm"] and r["coef"] < 1.0
        for k, got in zip(O.ENC_KEYS, enc):
            assert float(np.abs(got - p[k].numpy()).max()) <= 1e-6 * max(1.0, float(p[k].abs().max())), k
    for a, b in zip(res[0][7], res[1][7]):                 # identical gradients on both ranks: replicas stay in lock step
        assert np.array_equal(a, b)
