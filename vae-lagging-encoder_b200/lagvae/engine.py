"""Host-side engine: owns the device workspace and the per-shape plans, and turns torch tensors into
the raw pointers the C-ABI takes.  PyTorch is plumbing here (device memory, streams); every
arithmetic stage runs in liblagvae.so kernels.  No CPU path exists — CPU tensors raise."""
import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _backend as be
from . import graph as graph_mod

PARAM_NAMES = [
    "encoder.embed.weight", "encoder.lstm.weight_ih_l0", "encoder.lstm.weight_hh_l0",
    "encoder.lstm.bias_ih_l0", "encoder.lstm.bias_hh_l0", "encoder.linear.weight",
    "decoder.embed.weight", "decoder.trans_linear.weight", "decoder.lstm.weight_ih_l0",
    "decoder.lstm.weight_hh_l0", "decoder.lstm.bias_ih_l0", "decoder.lstm.bias_hh_l0",
    "decoder.pred_linear.weight",
]
N_ENC = 6


@dataclass
class DropoutSpec:
    """mode 0 = off (eval), 1 = explicit uint8 keep-masks, 2 = in-kernel Philox (include/lagvae.h).  seed_dev: optional
    int64 device tensor [1] added to `seed` when the kernels run (CUDA-graph replays bump it, lagvae/graph.py)."""
    mode: int = 0
    p_in: float = 0.0
    p_out: float = 0.0
    mask_in: Optional[torch.Tensor] = None   # uint8 [B, T-1, ni]
    mask_out: Optional[torch.Tensor] = None  # uint8 [B*ns, T-1, nh]
    seed: int = 0
    seed_dev: Optional[torch.Tensor] = None

    def to_c(self):
        d = be.Dropout()
        d.mode, d.p_in, d.p_out, d.seed = int(self.mode), float(self.p_in), float(self.p_out), int(self.seed) & (2 ** 64 - 1)
        d.mask_in = self.mask_in.data_ptr() if self.mask_in is not None else None
        d.mask_out = self.mask_out.data_ptr() if self.mask_out is not None else None
        d.seed_dev = self.seed_dev.data_ptr() if self.seed_dev is not None else None
        return d


def param_shapes(V, ni, nh, nz):
    return [(V, ni), (4 * nh, ni), (4 * nh, nh), (4 * nh,), (4 * nh,), (2 * nz, nh),
            (V, ni), (nh, nz), (4 * nh, ni + nz), (4 * nh, nh), (4 * nh,), (4 * nh,), (V, nh)]


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


import itertools
_ENGINE_UID = itertools.count(1)


class TextEngine:
    """One engine per (V, ni, nh, nz, device).  Plans (per B, T, ns) share one workspace that grows
    to the largest shape seen; only the most recent forward's stash is alive."""

    def __init__(self, V, ni, nh, nz, device, force_simt=None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise be.LagvaeError("lagvae kernels need a CUDA (B200, sm_100) device; got %r — "
                                 "there is no CPU fallback" % (device,))
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.V, self.ni, self.nh, self.nz = int(V), int(ni), int(nh), int(nz)
        if force_simt is None:
            force_simt = os.environ.get("LAGVAE_FORCE_SIMT", "0") == "1"
        self.flags = be.PLAN_FORCE_SIMT if force_simt else be.PLAN_DEFAULT
        self.shapes = param_shapes(self.V, self.ni, self.nh, self.nz)
        self.counts = [int(torch.Size(s).numel()) for s in self.shapes]
        self._ws = None
        self._plans = {}
        self._uid = next(_ENGINE_UID)
        self._ws_generation = 0   # bumped whenever the workspace is re-allocated: the cache inside it is gone
        self._dec_bump = 0     # decoder-weight updates made by liblagvae.so itself (outer_step): part of the weight epoch
        self._hooked = set()   # (B, T, ns) keys whose plans record the decoder-gradient event: re-applied when plans are re-created
        self.last_ns = 1
        self.generation = 0  # bumped by every forward; backward must match
        with torch.cuda.device(self.device):
            be.check(be.lib().lagvae_device_check(), "lagvae_device_check")

    # ---- plans / workspace -------------------------------------------------------------------
    def _dims(self, B, T, ns):
        return be.TextDims(B, T, ns, self.V, self.ni, self.nh, self.nz)

    def _drop_plans(self):
        for h in self._plans.values():
            be.lib().lagvae_text_plan_destroy(h)
        self._plans = {}

    def __del__(self):
        try:
            self._drop_plans()
        except Exception:
            pass

    def plan(self, B, T, ns):
        key = (int(B), int(T), int(ns))
        h = self._plans.get(key)
        if h is not None:
            return h
        d = self._dims(*key)
        need = be.lib().lagvae_text_workspace_bytes(C.byref(d), self.flags)
        if need == 0:
            raise be.LagvaeError("unsupported text dims %r" % (key,))
        if self._ws is None or self._ws.numel() < need:
            self._drop_plans()
            from .graph import retire
            retire(self._ws)           # a live CUDA graph may still hold pointers into the outgrown workspace
            self._ws = None
            self._ws = torch.empty(int(need), dtype=torch.uint8, device=self.device)
            self._ws_generation += 1
        out = C.c_void_p()
        with torch.cuda.device(self.device):
            be.check(be.lib().lagvae_text_plan_create(C.byref(d), self.flags, C.c_void_p(self._ws.data_ptr()),
                                                      self._ws.numel(), C.byref(out)), "lagvae_text_plan_create")
        self._plans[key] = out
        if key in self._hooked:          # the workspace grew and every plan was dropped: restore the event hook
            with torch.cuda.device(self.device):
                be.check(be.lib().lagvae_text_decoder_grads_event(out, 1), "lagvae_text_decoder_grads_event")
        return out

    # ---- argument marshalling ----------------------------------------------------------------
    def _params(self, params: Sequence[torch.Tensor], what="params"):
        if len(params) != be.NPARAM:
            raise be.LagvaeError("%s: expected %d tensors" % (what, be.NPARAM))
        tp = be.TextParams()
        for i, (t, shp) in enumerate(zip(params, self.shapes)):
            if t is None:       # slot not used by the entry point being called
                tp.p[i] = None
                continue
            if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != tuple(shp):
                raise be.LagvaeError("%s[%d] (%s): need contiguous fp32 %r on %s, got %s %r on %s" % (
                    what, i, PARAM_NAMES[i], tuple(shp), self.device, t.dtype, tuple(t.shape), t.device))
            tp.p[i] = t.data_ptr()
        return tp

    def _declare_decoder_epoch(self, plan, params):
        """Tell the plan whether the decoder weights changed since its last call (include/lagvae.h,
        lagvae_text_decoder_weights_epoch): the epoch is derived from the data pointers and torch's version counters of the
        7 decoder tensors — every in-place update (optimizer step, load_state_dict, copy_) bumps a counter — plus a counter
        for the updates liblagvae.so makes itself (outer_step), which torch cannot see."""
        dec = params[N_ENC:]
        capturing = torch.cuda.is_current_stream_capturing()
        serial = graph_mod.capture_serial() if capturing else None
        if any(t is None for t in dec) or (capturing and serial is None):
            epoch = 0        # bare torch.cuda.graph: the decision below would be frozen into the graph -> never cache
        elif capturing:
            # one epoch per GraphedStep capture: the first captured step re-splits at every replay, the rest reuse (graph.py)
            epoch = ((serial * 2654435761 + self._uid * 998244353 + 0x5bd1e995) & (2 ** 63 - 1)) | 1
        else:
            # engine uid + workspace generation: a NEW engine / workspace must never hit a cache entry that an earlier one
            # left behind at a recycled address (the allocator hands freed workspaces and tensors out again)
            h = self._dec_bump * 1000003 + self._uid * 998244353 + self._ws_generation * 7919
            for t in dec:
                h = (h * 1000003 + t.data_ptr() * 31 + t._version) & (2 ** 63 - 1)
            epoch = h | 1
        be.check(be.lib().lagvae_text_decoder_weights_epoch(plan, C.c_uint64(epoch)), "lagvae_text_decoder_weights_epoch")

    def _x(self, x):
        if x.device != self.device or x.dtype != torch.int64 or x.dim() != 2:
            raise be.LagvaeError("x must be an int64 [B,T] tensor on %s (got %s %r on %s); there is no CPU path"
                                 % (self.device, x.dtype, tuple(x.shape), x.device))
        return x if x.is_contiguous() else x.contiguous()

    def _f32(self, t, shape, name):
        if t is None:
            return None
        if t.device != self.device or t.dtype != torch.float32 or tuple(t.shape) != tuple(shape):
            raise be.LagvaeError("%s must be fp32 %r on %s" % (name, tuple(shape), self.device))
        return t if t.is_contiguous() else t.contiguous()

    def _check_drop(self, drop, B, T, ns):
        if drop is None:
            return DropoutSpec()
        if drop.mode == 1:
            for m, shp, p in ((drop.mask_in, (B, T - 1, self.ni), drop.p_in), (drop.mask_out, (B * ns, T - 1, self.nh), drop.p_out)):
                if p > 0 and (m is None or m.dtype != torch.uint8 or tuple(m.shape) != shp or m.device != self.device
                              or not m.is_contiguous()):
                    raise be.LagvaeError("dropout mode 1 needs contiguous uint8 mask of shape %r" % (shp,))
        return drop

    # ---- entry points --------------------------------------------------------------------------
    def loss_forward(self, params, x, eps, kl_weight, drop=None, want_stats=False):
        x = self._x(x)
        B, T = x.shape
        ns = int(eps.shape[1])
        eps = self._f32(eps, (B, ns, self.nz), "eps")
        drop = self._check_drop(drop, B, T, ns)
        tp = self._params(params)
        h = self.plan(B, T, ns)
        self._declare_decoder_epoch(h, params)
        out = [torch.empty(B, dtype=torch.float32, device=self.device) for _ in range(3)]
        mu = logvar = z = None
        if want_stats:
            mu = torch.empty(B, self.nz, dtype=torch.float32, device=self.device)
            logvar = torch.empty_like(mu)
            z = torch.empty(B, ns, self.nz, dtype=torch.float32, device=self.device)
        dc = drop.to_c()
        self.generation += 1
        self.last_ns = ns
        with torch.cuda.device(self.device):
            be.check(be.lib().lagvae_text_loss_forward(
                h, C.byref(tp), be.ptr(x), be.ptr(eps), float(kl_weight), C.byref(dc), be.ptr(out[0]),
                be.ptr(out[1]), be.ptr(out[2]), be.ptr(mu), be.ptr(logvar), be.ptr(z), _stream()),
                "lagvae_text_loss_forward")
        self._last = (x, eps, drop)  # keep inputs alive until backward
        if want_stats:
            return out[0], out[1], out[2], mu, logvar, z
        return out[0], out[1], out[2]

    def loss_backward(self, params, x, g_loss, g_rec, g_kl, generation=None, grads_out=None,
                      decoder_wgrad_norm_only=False) -> List[torch.Tensor]:
        if generation is not None and generation != self.generation:
            raise be.LagvaeError("backward called after another forward on the same engine: the stash of "
                                 "this loss has been overwritten (one live VAE.loss graph per model)")
        x = self._x(x)
        B, T = x.shape
        ns = self._last[1].shape[1]
        tp = self._params(params)
        gl, gr, gk = (self._f32(g, (B,), n) for g, n in ((g_loss, "g_loss"), (g_rec, "g_rec"), (g_kl, "g_kl")))
        grads = grads_out if grads_out is not None else [torch.empty(s, dtype=torch.float32, device=self.device) for s in self.shapes]
        tg = self._params(grads, "grads")
        self._declare_decoder_epoch(self.plan(B, T, ns), params)
        with torch.cuda.device(self.device):
            be.check(be.lib().lagvae_text_loss_backward(self.plan(B, T, ns), C.byref(tp), be.ptr(x), be.ptr(gl),
                                                        be.ptr(gr), be.ptr(gk), C.byref(tg),
                                                        be.BWD_DECODER_WGRAD_NORM_ONLY if decoder_wgrad_norm_only else be.BWD_DEFAULT,
                                                        _stream()),
                     "lagvae_text_loss_backward")
        return grads

    def encode_stats(self, params, x):
        x = self._x(x)
        B, T = x.shape
        tp = self._params(params)
        mu = torch.empty(B, self.nz, dtype=torch.float32, device=self.device)
        logvar = torch.empty_like(mu)
        self.generation += 1
        with torch.cuda.device(self.device):
            be.check(be.lib().lagvae_text_encode_stats(self.plan(B, T, 1), C.byref(tp), be.ptr(x), be.ptr(mu),
                                                       be.ptr(logvar), _stream()), "lagvae_text_encode_stats")
        return mu, logvar

    def reconstruct_error(self, params, x, z, drop=None):
        x = self._x(x)
        B, T = x.shape
        ns = int(z.shape[1])
        z = self._f32(z, (B, ns, self.nz), "z")
        drop = self._check_drop(drop, B, T, ns)
        tp = self._params(params)
        out = torch.empty(B, ns, dtype=torch.float32, device=self.device)
        dc = drop.to_c()
        self.generation += 1
        self._declare_decoder_epoch(self.plan(B, T, ns), params)
        with torch.cuda.device(self.device):
            be.check(be.lib().lagvae_text_reconstruct_error(self.plan(B, T, ns), C.byref(tp), be.ptr(x), be.ptr(z),
                                                            C.byref(dc), be.ptr(out), _stream()),
                     "lagvae_text_reconstruct_error")
        return out

    def decode_logits(self, params, src, z, drop=None):
        """LSTMDecoder.decode (dec_lstm.py:66-111): src int64 [B, T'] (decoder input tokens), z [B, ns, nz] ->
        logits [B*ns, T', V].  Forward only."""
        src = self._x(src)
        B, Td = src.shape
        ns = int(z.shape[1])
        z = self._f32(z, (B, ns, self.nz), "z")
        drop = self._check_drop(drop, B, Td + 1, ns)
        tp = self._params(params)
        out = torch.empty(B * ns, Td, self.V, dtype=torch.float32, device=self.device)
        dc = drop.to_c()
        self.generation += 1
        self._declare_decoder_epoch(self.plan(B, Td + 1, ns), params)
        with torch.cuda.device(self.device):
            be.check(be.lib().lagvae_text_decode_logits(self.plan(B, Td + 1, ns), C.byref(tp), be.ptr(src), be.ptr(z),
                                                        C.byref(dc), be.ptr(out), _stream()), "lagvae_text_decode_logits")
        return out

    def mi(self, mu, logvar, eps):
        B, nz = mu.shape
        mu, logvar = self._f32(mu, (B, nz), "mu"), self._f32(logvar, (B, nz), "logvar")
        eps = self._f32(eps.reshape(B, nz), (B, nz), "eps")
        out = torch.empty(1, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            be.check(be.lib().lagvae_mi_estimate(be.ptr(mu), be.ptr(logvar), be.ptr(eps), B, nz, be.ptr(out),
                                                 _stream()), "lagvae_mi_estimate")
        return out

    def clip_sgd(self, params, grads, n_update, max_norm, lr, scale_all=True):
        n = len(grads)
        P = (C.c_void_p * n)(*[p.data_ptr() for p in params])
        G = (C.c_void_p * n)(*[g.data_ptr() for g in grads])
        cnt = (C.c_int64 * n)(*[g.numel() for g in grads])
        norm = torch.empty(1, dtype=torch.float32, device=self.device)
        scratch = torch.empty(16384, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            be.check(be.lib().lagvae_clip_sgd_step(P, G, cnt, n, n_update, float(max_norm), float(lr),
                                                   1 if scale_all else 0, be.ptr(norm), be.ptr(scratch), _stream()),
                     "lagvae_clip_sgd_step")
        return norm

    @property
    def decoder_offset(self):
        """Index of the first decoder gradient in the flat bucket (encoder tensors come first, as in vae.parameters())."""
        return sum(self.counts[:6])

    def enable_decoder_grads_event(self, B, T, ns=1, enable=True):
        """Data-parallel overlap hook (lagvae.h): loss_backward of the (B, T, ns) plan records an event once the 7 decoder
        gradients are final."""
        key = (int(B), int(T), int(ns))
        if enable and key in self._hooked and key in self._plans:
            return
        (self._hooked.add if enable else self._hooked.discard)(key)
        with torch.cuda.device(self.device):
            be.check(be.lib().lagvae_text_decoder_grads_event(self.plan(B, T, ns), 1 if enable else 0),
                     "lagvae_text_decoder_grads_event")

    def wait_decoder_grads(self, B, T, ns, stream):
        """Make `stream` (torch.cuda.Stream) wait for the decoder-gradient event of the last loss_backward."""
        with torch.cuda.device(self.device):
            be.check(be.lib().lagvae_text_wait_decoder_grads(self.plan(B, T, ns), C.c_void_p(stream.cuda_stream)),
                     "lagvae_text_wait_decoder_grads")

    def grad_workspace(self):
        return torch.empty(sum(self.counts), dtype=torch.float32, device=self.device)

    def split_grads(self, flat):
        out, off = [], 0
        for s, n in zip(self.shapes, self.counts):
            out.append(flat[off:off + n].view(s))
            off += n
        return out

    def inner_step(self, params, x, eps, kl_weight, drop, grad_ws, out_loss, out_scalars, max_norm=5.0, lr=1.0):
        """One fused iteration of text.py:371-391 (device side).  out_loss [B], out_scalars [4] =
        {Σloss, Σrec, ΣKL, grad-norm} are caller-owned device tensors."""
        x = self._x(x)
        B, T = x.shape
        ns = int(eps.shape[1])
        drop = self._check_drop(drop, B, T, ns)
        tp = self._params(params)
        dc = drop.to_c()
        self.generation += 1
        self._declare_decoder_epoch(self.plan(B, T, ns), params)
        with torch.cuda.device(self.device):
            be.check(be.lib().lagvae_text_inner_step(self.plan(B, T, ns), C.byref(tp), be.ptr(x), be.ptr(eps),
                                                     float(kl_weight), C.byref(dc), float(max_norm), float(lr),
                                                     be.ptr(grad_ws), be.ptr(out_loss), be.ptr(out_scalars),
                                                     _stream()), "lagvae_text_inner_step")
        self._last = (x, eps, drop)

    def outer_step(self, params, x, eps, kl_weight, drop, grad_ws, out_loss, out_scalars, update_encoder, max_norm=5.0, lr=1.0):
        """The fused decoder-update step of text.py:407-424 (device side): decoder SGD step, plus the encoder's when
        `update_encoder` (aggressive phase over).  Same buffers as inner_step."""
        x = self._x(x)
        B, T = x.shape
        ns = int(eps.shape[1])
        drop = self._check_drop(drop, B, T, ns)
        tp = self._params(params)
        dc = drop.to_c()
        self.generation += 1
        self._declare_decoder_epoch(self.plan(B, T, ns), params)
        self._dec_bump += 1                       # this call steps the decoder weights in place
        with torch.cuda.device(self.device):
            be.check(be.lib().lagvae_text_outer_step(self.plan(B, T, ns), C.byref(tp), be.ptr(x), be.ptr(eps),
                                                     float(kl_weight), C.byref(dc), float(max_norm), float(lr),
                                                     1 if update_encoder else 0, be.ptr(grad_ws), be.ptr(out_loss),
                                                     be.ptr(out_scalars), _stream()), "lagvae_text_outer_step")
        self._last = (x, eps, drop)
