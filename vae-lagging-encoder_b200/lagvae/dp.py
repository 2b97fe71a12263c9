"""Batch-sharded data parallelism for the aggressive inner step.

One process per GPU; every rank holds the full parameters, runs forward+backward on its own rows with the
upstream gradient of the GLOBAL mean (1 / B_global), then ONE all-reduce (sum) of the flat gradient bucket per
inner step; clip_grad_norm_ (text.py:385) is evaluated on the averaged full gradient — identical on all ranks —
and the encoder-only SGD step (text.py:387) keeps the replicas bit-identical without any parameter broadcast.

The arithmetic back-end is injected (`LocalBackend` protocol): the product uses `EngineBackend` (CUDA kernels,
NCCL); the CPU tests drive the same host logic over gloo with an oracle back-end."""
from typing import List, Protocol, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row shard [lo, hi) of a batch of n_rows; ragged batches (text_data.py:241-247 yields
    batches smaller than batch_size) give the first n_rows % world ranks one extra row; shards may be empty."""
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class LocalBackend(Protocol):
    def forward_backward(self, params: Sequence[torch.Tensor], x: torch.Tensor, g_scale: float,
                         flat_grads: torch.Tensor) -> torch.Tensor:
        """Run VAE.loss forward + backward on local rows with upstream gradient g_scale per row, write the 13
        gradients into flat_grads (reference parameter order) and return the local per-row loss [b_local]."""

    def clip_sgd(self, params: Sequence[torch.Tensor], flat_grads: torch.Tensor, max_norm: float, lr: float) -> float:
        """clip_grad_norm_ over the whole bucket + SGD on the encoder tensors; returns the pre-clip norm."""


def dp_inner_step(backend: LocalBackend, params: Sequence[torch.Tensor], x_global: torch.Tensor,
                  flat_grads: torch.Tensor, group=None, max_norm: float = 5.0, lr: float = 1.0,
                  presharded: bool = False, global_rows: int = None):
    """One aggressive inner step under DP.  Returns (Σloss over the global batch, pre-clip grad norm).
    x_global is the full batch (every rank runs the unmodified SPMD driver with identical seeds)
    unless presharded=True, in which case it already is this rank's shard and global_rows must be given."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if presharded:
        x_local, n_global = x_global, int(global_rows)
    else:
        lo, hi = shard_bounds(x_global.shape[0], rank, world)
        x_local, n_global = x_global[lo:hi], x_global.shape[0]
    if x_local.shape[0] > 0:
        loss_local = backend.forward_backward(params, x_local, 1.0 / n_global, flat_grads)   # loss.mean() over the GLOBAL batch
        loss_sum = loss_local.sum().reshape(1).to(torch.float32)
    else:                                   # empty shard: contributes zeros
        flat_grads.zero_()
        loss_sum = torch.zeros(1, dtype=torch.float32, device=flat_grads.device)
    if world > 1:
        # the data-path collective: one sum of the flat gradient.  Back-ends that expose `decoder_offset` get it as two
        # buckets — [decoder | encoder] — and, when they also provide `decoder_grads_stream()`, the decoder bucket
        # (70% of the bytes, final before the encoder LSTM backward starts) is reduced on a side stream underneath
        # the still-running encoder backward.
        off = getattr(backend, "decoder_offset", None)
        if off is None:
            dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
        else:                                   # the bucket split depends on the back-end only: identical on every rank
            side = (backend.decoder_grads_stream(x_local)
                    if x_local.shape[0] > 0 and hasattr(backend, "decoder_grads_stream") else None)
            if side is not None:
                with torch.cuda.stream(side):
                    dist.all_reduce(flat_grads[off:], op=dist.ReduceOp.SUM, group=group)
                dist.all_reduce(flat_grads[:off], op=dist.ReduceOp.SUM, group=group)
                torch.cuda.current_stream().wait_stream(side)
            else:
                dist.all_reduce(flat_grads[off:], op=dist.ReduceOp.SUM, group=group)
                dist.all_reduce(flat_grads[:off], op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(loss_sum, op=dist.ReduceOp.SUM, group=group)            # 4 bytes: Σloss for text.py:381
    norm = backend.clip_sgd(params, flat_grads, max_norm, lr)
    return float(loss_sum), norm


def accumulated_inner_step(backend: LocalBackend, params: Sequence[torch.Tensor], x: torch.Tensor, flat_grads: torch.Tensor,
                           scratch_grads: torch.Tensor, max_rows: int, max_norm: float = 5.0, lr: float = 1.0, on_chunk=None):
    """One aggressive inner step on ONE device with the batch run as micro-batches of <= max_rows rows (the persistent
    LSTM kernels keep <= 256 batch rows on chip): every micro-batch back-propagates the upstream gradient of the GLOBAL
    mean (1 / B), the flat gradients accumulate, then ONE clip_grad_norm_ + encoder SGD — the arithmetic of
    dp_inner_step with the shards executed back to back instead of on different GPUs.  `on_chunk(lo)` (optional) is
    called before each micro-batch (row offset).  Returns (Σloss, pre-clip grad norm)."""
    n = int(x.shape[0])
    loss_sum = None
    for i, lo in enumerate(range(0, n, int(max_rows))):
        if on_chunk is not None:
            on_chunk(lo)
        buf = flat_grads if i == 0 else scratch_grads
        loss = backend.forward_backward(params, x[lo:lo + int(max_rows)], 1.0 / n, buf)
        if i:
            flat_grads.add_(scratch_grads)
        part = loss.sum().reshape(1).to(torch.float32)
        loss_sum = part if loss_sum is None else loss_sum + part
    norm = backend.clip_sgd(params, flat_grads, max_norm, lr)
    return float(loss_sum), norm


class EngineBackend:
    """Product back-end: lagvae.TextEngine kernels; flat_grads is the engine's flat gradient workspace."""

    def __init__(self, engine, kl_weight, eps_fn, drop_fn, overlap=True):
        self.engine, self.kl_weight, self.eps_fn, self.drop_fn = engine, kl_weight, eps_fn, drop_fn
        self._views = None
        self.decoder_offset = engine.decoder_offset
        self._overlap = overlap
        self._side = None
        self._hooked = set()

    def decoder_grads_stream(self, x_local):
        """Side stream that already waits for the decoder gradients of the backward just enqueued (None = no overlap)."""
        if not self._overlap:
            return None
        B, T = int(x_local.shape[0]), int(x_local.shape[1])
        if self._side is None:
            self._side = torch.cuda.Stream(device=x_local.device)
        self.engine.wait_decoder_grads(B, T, 1, self._side)
        return self._side

    def forward_backward(self, params, x, g_scale, flat_grads):
        eng = self.engine
        if self._views is None or self._views[0].data_ptr() != flat_grads.data_ptr():
            self._views = eng.split_grads(flat_grads)
        B = x.shape[0]
        if self._overlap and (B, x.shape[1]) not in self._hooked:
            eng.enable_decoder_grads_event(B, x.shape[1], 1)
            self._hooked.add((B, x.shape[1]))
        loss, _, _ = eng.loss_forward(params, x, self.eps_fn(B), self.kl_weight, self.drop_fn())
        gl = torch.full((B,), g_scale, dtype=torch.float32, device=x.device)
        # aggressive loop: decoder weights are not stepped, their gradients only enter the clip norm
        eng.loss_backward(params, x, gl, None, None, grads_out=self._views, decoder_wgrad_norm_only=True)
        return loss

    def clip_sgd(self, params, flat_grads, max_norm, lr):
        if self._views is None or self._views[0].data_ptr() != flat_grads.data_ptr():   # e.g. after a scratch-buffer micro-batch
            self._views = self.engine.split_grads(flat_grads)
        return self.engine.clip_sgd(params, self._views, 6, max_norm, lr, scale_all=False)
