"""Drop-in `VAE` (reference modules/vae.py:9-303): same constructor, methods and return contracts; the
training objective `loss()` is ONE differentiable node whose forward and backward are the fused B200
launch sequences in liblagvae.so."""
import math

import torch
import torch.nn as nn

from lagvae import LagvaeError
from lagvae.dp import ShardedTextLoss, dp_group

from .text import LSTMDecoder, LSTMEncoder, _dec_params, _enc_params, dropout_spec, get_engine
from .utils import log_sum_exp


class _TextLossFn(torch.autograd.Function):
    """(loss, rec, KL) = VAE.loss(x)  — vae.py:79-98.  backward = the autograd walk of text.py:382-384
    implemented by lagvae_text_loss_backward (all 13 gradients; decoder ones included because
    clip_grad_norm_ at text.py:385 reads them)."""

    @staticmethod
    def forward(ctx, engine, x, kl_weight, eps, drop, *params):
        p = [q.detach() for q in params]
        loss, rec, kl = engine.loss_forward(p, x, eps, kl_weight, drop)
        ctx.engine, ctx.x, ctx.gen, ctx.p = engine, x, engine.generation, p
        return loss, rec, kl

    @staticmethod
    def backward(ctx, g_loss, g_rec, g_kl):
        grads = ctx.engine.loss_backward(ctx.p, ctx.x, g_loss, g_rec, g_kl, generation=ctx.gen)
        return (None, None, None, None, None, *grads)


class VAE(nn.Module):
    """VAE with N(0, I) prior (reference vae.py:9-23)."""

    def __init__(self, encoder, decoder, args):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
        self.args = args
        self.nz = args.nz
        loc = torch.zeros(self.nz, device=args.device)
        scale = torch.ones(self.nz, device=args.device)
        self.prior = torch.distributions.normal.Normal(loc, scale)

    # convenience names used by north_star; the contract remains the nn.Module parameter API
    def encoder_params(self):
        return list(self.encoder.parameters())

    def decoder_params(self):
        return list(self.decoder.parameters())

    def _is_text(self):
        return isinstance(self.encoder, LSTMEncoder) and isinstance(self.decoder, LSTMDecoder)

    def _text_engine(self):
        e, d = self.encoder, self.decoder
        if (e.ni, e.nh, e.nz) != (d.ni, d.nh, d.nz):
            raise LagvaeError("fused text path needs matching encoder/decoder ni, nh, nz")
        return get_engine(e.vocab_size, e.ni, e.nh, e.nz, e.embed.weight.device)

    # ------------------------------------------------------------------------------------------
    def encode(self, x, nsamples=1):
        """(z [B,ns,nz], KL [B])  — vae.py:25-31."""
        return self.encoder.encode(x, nsamples)

    def encode_stats(self, x):
        """(mean, logvar), each [B,nz] — vae.py:33-40."""
        return self.encoder(x)

    def decode(self, z, strategy, K=5):
        if strategy == "beam":
            return self.decoder.beam_search_decode(z, K)
        if strategy == "greedy":
            return self.decoder.greedy_decode(z)
        if strategy == "sample":
            return self.decoder.sample_decode(z)
        raise ValueError("the decoding strategy is not supported")       # vae.py:61

    def reconstruct(self, x, decoding_strategy="greedy", K=5):
        z = self.sample_from_inference(x).squeeze(1)
        return self.decode(z, decoding_strategy, K)

    def loss(self, x, kl_weight, nsamples=1):
        """(rec + kl_weight*KL, rec, KL), each [B] — vae.py:79-98."""
        if not self._is_text():
            z, KL = self.encode(x, nsamples)
            rec = self.decoder.reconstruct_error(x, z).mean(dim=1)
            return rec + kl_weight * KL, rec, KL
        eng = self._text_engine()
        B, T = x.shape
        eps = torch.empty(B, nsamples, self.nz, dtype=torch.float32, device=x.device).normal_()  # encoder.py:77
        params = _enc_params(self.encoder) + _dec_params(self.decoder)
        group = dp_group()
        if group is not None:
            # SPMD data parallelism inside the boundary: every rank was handed the same full batch and
            # drew the same eps; each computes its row shard and the autograd node exchanges (lagvae/dp.py)
            drop_fn = lambda lo, hi: dropout_spec(self.decoder, hi - lo, T, nsamples, x.device, row_offset=lo, rows_total=B)
            return ShardedTextLoss.apply(eng, group, x, float(kl_weight), eps, drop_fn, *params)
        drop = dropout_spec(self.decoder, B, T, nsamples, x.device)
        return _TextLossFn.apply(eng, x, float(kl_weight), eps, drop, *params)

    def nll_iw(self, x, nsamples, ns=100):
        """Importance-weighted NLL estimate, [B] — vae.py:100-129."""
        tmp = []
        for _ in range(int(nsamples / ns)):
            z, param = self.encoder.sample(x, ns)
            tmp.append(self.eval_complete_ll(x, z) - self.eval_inference_dist(x, z, param))
        return -(log_sum_exp(torch.cat(tmp, dim=-1), dim=-1) - math.log(nsamples))

    def KL(self, x):
        return self.encode(x, 1)[1]

    def eval_prior_dist(self, zrange):
        return self.prior.log_prob(zrange).sum(dim=-1)

    def eval_complete_ll(self, x, z):
        return self.eval_prior_dist(z) + self.eval_cond_ll(x, z)

    def eval_cond_ll(self, x, z):
        return self.decoder.log_probability(x, z)

    def eval_log_model_posterior(self, x, grid_z):
        B = x.size(0)
        grid_z = grid_z.unsqueeze(0).expand(B, *grid_z.size()).contiguous()
        log_comp = self.eval_complete_ll(x, grid_z)
        return log_comp - log_sum_exp(log_comp, dim=1, keepdim=True)

    def sample_from_prior(self, nsamples):
        return self.prior.sample((nsamples,))

    def sample_from_inference(self, x, nsamples=1):
        return self.encoder.sample(x, nsamples)[0]

    def sample_from_posterior(self, x, nsamples):
        """Metropolis-Hastings samples from the model posterior, [B, nsamples, nz] — vae.py:218-254.  The reference's version is
        unreachable (it calls `encoder.sample_from_inference`, which no encoder defines, and no driver sets `args.mh_*`); this
        is the same chain started from `self.sample_from_inference` so that the method exists with the documented contract.
        Needs args.mh_burn_in / mh_thin / mh_std, like the reference."""
        cur = self.sample_from_inference(x, 1)                       # [B, 1, nz]
        cur_ll = self.eval_complete_ll(x, cur)                       # [B, 1]
        total_iter = self.args.mh_burn_in + nsamples * self.args.mh_thin
        samples = []
        for it in range(total_iter):
            nxt = torch.normal(mean=cur, std=cur.new_full(cur.size(), self.args.mh_std))
            nxt_ll = self.eval_complete_ll(x, nxt)
            accept = torch.min((nxt_ll - cur_ll).exp(), torch.ones_like(cur_ll))
            mask = (torch.empty_like(accept).uniform_() < accept).float()
            cur = mask.unsqueeze(2) * nxt + (1 - mask.unsqueeze(2)) * cur
            cur_ll = mask * nxt_ll + (1 - mask) * cur_ll
            if it >= self.args.mh_burn_in and (it - self.args.mh_burn_in) % self.args.mh_thin == 0:
                samples.append(cur.unsqueeze(1))
        return torch.cat(samples, dim=1)

    def calc_model_posterior_mean(self, x, grid_z):
        posterior = self.eval_log_model_posterior(x, grid_z).exp()
        return torch.mul(posterior.unsqueeze(2), grid_z.unsqueeze(0)).sum(1)

    def calc_infer_mean(self, x):
        return self.encoder.forward(x)[0]

    def eval_inference_dist(self, x, z, param=None):
        return self.encoder.eval_inference_dist(x, z, param)

    def calc_mi_q(self, x):
        """MI under q(z|x) -> Python float — vae.py:295-303."""
        return self.encoder.calc_mi(x)
