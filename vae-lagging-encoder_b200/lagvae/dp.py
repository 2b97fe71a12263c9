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
                  presharded: bool = False, global_rows: int = None, read_loss: bool = True):
    """One aggressive inner step under DP.  Returns (Σloss over the global batch, pre-clip grad norm); read_loss=False
    leaves Σloss on the device (a [1] tensor) instead of synchronising on it (text.py:381 reads it every step, a fused
    caller only needs the 15-step window sum, text.py:389-398).
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
        # lagvae_allreduce_bucket (communicator owned by liblagvae.so) on CUDA, torch.distributed otherwise (gloo tests)
        comm = _bucket_comm(group if group is not None else dist.group.WORLD, flat_grads.device)
        reduce = (lambda t, st=None: comm.all_reduce(t, st)) if comm is not None else \
                 (lambda t, st=None: dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group))
        side = None
        if off is not None and x_local.shape[0] > 0 and hasattr(backend, "decoder_grads_stream"):
            side = backend.decoder_grads_stream(x_local)
        if off is None or (side is None and getattr(backend, "single_bucket", False)):
            reduce(flat_grads)                  # one bucket
        elif side is not None:                  # the bucket split depends on the back-end only: identical on every rank
            with torch.cuda.stream(side):
                reduce(flat_grads[off:], side)
            reduce(flat_grads[:off])
            torch.cuda.current_stream().wait_stream(side)
        else:
            reduce(flat_grads[off:])
            reduce(flat_grads[:off])
        dist.all_reduce(loss_sum, op=dist.ReduceOp.SUM, group=group)            # 4 bytes: Σloss for text.py:381
    norm = backend.clip_sgd(params, flat_grads, max_norm, lr)
    return (float(loss_sum) if read_loss else loss_sum), norm


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


# ----------------------------------------------------------------------------------------------------------------------
# b3: the UNMODIFIED driver runs SPMD — every rank executes the same script with the same seeds on the same
# full batch; `VAE.loss` shards the batch by rank internally and returns the FULL [B] vectors, so `loss.sum().item()`,
# the break rule (text.py:393-398) and every later batch pick agree on all ranks.  The exchange is in the autograd node:
#   forward : local rows -> (loss, rec, KL)[lo:hi] ; one all-reduce (sum) of a zero-padded [3, B] tensor = the all-gather
#   backward: upstream gradients sliced to [lo:hi] -> local backward into a flat bucket -> ONE all-reduce (sum) of the
#             bucket (decoder part first, on a side stream under the encoder backward when the engine exposes the hook)
# After it every rank holds the gradient of the global mean: clip_grad_norm_ / SGD run by the driver keep the replicas
# identical without a parameter broadcast.
class ShardedTextLoss(torch.autograd.Function):
    """(loss, rec, KL), each the full [B], for a batch sharded over the ranks of `group`.

    `engine` needs: loss_forward(params, x, eps, kl_weight, drop) -> 3 x [b], generation, loss_backward(params, x, g_loss,
    g_rec, g_kl, generation=, grads_out=) and (optionally) decoder_offset / enable_decoder_grads_event / wait_decoder_grads —
    lagvae.TextEngine on the GPU, an oracle stand-in in the gloo CPU test."""

    @staticmethod
    def forward(ctx, engine, group, x, kl_weight, eps, drop_fn, *params):
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        B = int(x.shape[0])
        lo, hi = shard_bounds(B, rank, world)
        p = [q.detach() for q in params]
        out = torch.zeros(3, B, dtype=torch.float32, device=x.device)
        ctx.engine, ctx.group, ctx.p, ctx.span, ctx.B = engine, group, p, (lo, hi), B
        ctx.x_local = x[lo:hi].contiguous() if hi > lo else None
        ctx.gen = None
        if hi > lo:
            loss, rec, kl = engine.loss_forward(p, ctx.x_local, eps[lo:hi].contiguous(), kl_weight, drop_fn(lo, hi))
            ctx.gen = engine.generation
            out[0, lo:hi], out[1, lo:hi], out[2, lo:hi] = loss, rec, kl
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)          # 3*B floats: the all-gather of SURVEY §8 b3
        return out[0], out[1], out[2]

    @staticmethod
    def backward(ctx, g_loss, g_rec, g_kl):
        eng, (lo, hi) = ctx.engine, ctx.span
        dev = g_loss.device
        # a fresh bucket per backward: the returned gradients are views of it and may become the parameters' .grad
        flat = torch.empty(sum(int(q.numel()) for q in ctx.p), dtype=torch.float32, device=dev)
        views, off = [], 0
        for q in ctx.p:
            views.append(flat[off:off + q.numel()].view(q.shape))
            off += q.numel()
        dec_off = getattr(eng, "decoder_offset", None)
        side = None
        if hi > lo:
            Bl, T = int(ctx.x_local.shape[0]), int(ctx.x_local.shape[1])
            hook = dec_off is not None and hasattr(eng, "enable_decoder_grads_event") and dev.type == "cuda" and _overlap_default()
            if hook:
                eng.enable_decoder_grads_event(Bl, T, getattr(eng, "last_ns", 1))
            sl = lambda g: None if g is None else g[lo:hi].contiguous()
            eng.loss_backward(ctx.p, ctx.x_local, sl(g_loss), sl(g_rec), sl(g_kl), generation=ctx.gen, grads_out=views)
            if hook:
                side = _side_stream(dev)
                eng.wait_decoder_grads(Bl, T, getattr(eng, "last_ns", 1), side)
        else:
            flat.zero_()                                                  # empty shard (B < world): contributes zeros
        # the collective itself: lagvae_allreduce_bucket (NCCL communicator owned by liblagvae.so) on CUDA, torch.distributed
        # otherwise (gloo in the CPU tests, or LAGVAE_DP_COMM=torch)
        comm = _bucket_comm(ctx.group, dev)
        reduce = (lambda t, st=None: comm.all_reduce(t, st)) if comm is not None else \
                 (lambda t, st=None: dist.all_reduce(t, op=dist.ReduceOp.SUM, group=ctx.group))
        if dec_off is None or not _overlap_default():
            reduce(flat)                 # one bucket (LAGVAE_DP_OVERLAP=0: identical decision on every rank)
        elif side is not None:
            flat.record_stream(side)
            with torch.cuda.stream(side):
                reduce(flat[dec_off:], side)
            reduce(flat[:dec_off])
            torch.cuda.current_stream().wait_stream(side)
        else:                    # same two buckets in the same order on every rank (an empty-shard rank has no event to wait on)
            reduce(flat[dec_off:])
            reduce(flat[:dec_off])
        return (None, None, None, None, None, None, *views)


_SIDE = {}


def _overlap_default():
    """LAGVAE_DP_OVERLAP=1: all-reduce the decoder part of the bucket on a side stream as soon as the decoder gradients are
    final (under the encoder backward).  0 (default): ONE all-reduce of the whole bucket after the backward — the backward
    then keeps its own side-stream work (dW_pred under the recurrences, csrc/text_plan.cu), which the early hand-over
    excludes, and no NCCL CTAs compete with the recurrence clusters for SMs.  Measured on 2 B200s (profiles/README.md r2v):
    8.29 ms/step without the early hand-over against 8.69 ms with it."""
    import os
    return os.environ.get("LAGVAE_DP_OVERLAP", "0") != "0"


_COMMS = {}


def _bucket_comm(group, dev):
    """lagvae.BucketComm for `group` (created on first use; collective), or None when the bucket is not on a CUDA device or
    LAGVAE_DP_COMM=torch asks for torch.distributed's all_reduce."""
    import os
    if dev.type != "cuda" or os.environ.get("LAGVAE_DP_COMM", "lagvae") == "torch":
        return None
    key = id(group)
    if key not in _COMMS:
        from .optim import BucketComm
        with torch.cuda.device(dev):
            _COMMS[key] = BucketComm(group)
    return _COMMS[key]



def _side_stream(dev):
    key = (dev.type, dev.index)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _SIDE[key]


def dp_group():
    """The process group `VAE.loss` shards over: the default group when torch.distributed is initialised with more than
    one rank and LAGVAE_DP != 0; None otherwise (single-process semantics)."""
    import os
    if os.environ.get("LAGVAE_DP", "1") == "0" or not dist.is_available() or not dist.is_initialized():
        return None
    return dist.group.WORLD if dist.get_world_size() > 1 else None


class EngineBackend:
    """Product back-end: lagvae.TextEngine kernels; flat_grads is the engine's flat gradient workspace."""

    def __init__(self, engine, kl_weight, eps_fn, drop_fn, overlap=None):
        overlap = _overlap_default() if overlap is None else overlap
        self.engine, self.kl_weight, self.eps_fn, self.drop_fn = engine, kl_weight, eps_fn, drop_fn
        self._views = None
        self.decoder_offset = engine.decoder_offset
        self._overlap = overlap
        self.single_bucket = not overlap       # no early hand-over of the decoder gradients -> one all-reduce of the whole bucket
        self._side = None

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
        if self._overlap:
            eng.enable_decoder_grads_event(B, x.shape[1], 1)     # idempotent; the engine re-applies it when plans are re-created
        loss, _, _ = eng.loss_forward(params, x, self.eps_fn(B), self.kl_weight, self.drop_fn())
        gl = torch.full((B,), g_scale, dtype=torch.float32, device=x.device)
        # aggressive loop: decoder weights are not stepped, their gradients only enter the clip norm
        eng.loss_backward(params, x, gl, None, None, grads_out=self._views, decoder_wgrad_norm_only=True)
        return loss

    def clip_sgd(self, params, flat_grads, max_norm, lr):
        if self._views is None or self._views[0].data_ptr() != flat_grads.data_ptr():   # e.g. after a scratch-buffer micro-batch
            self._views = self.engine.split_grads(flat_grads)
        return self.engine.clip_sgd(params, self._views, 6, max_norm, lr, scale_all=False)
