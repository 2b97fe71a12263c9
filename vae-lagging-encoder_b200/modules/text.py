"""Drop-in text modules: LSTMEncoder / LSTMDecoder with the reference constructors, attribute names,
state_dict keys and method contracts (reference modules/encoders/enc_lstm.py:10-64,
modules/encoders/encoder.py:7-145, modules/decoders/dec_lstm.py:17-161, decoder.py:5-73), computing
on the B200 kernels of liblagvae.so through lagvae.TextEngine.  nn.Embedding / nn.LSTM / nn.Linear
objects are kept only as *parameter containers* (so that `.parameters()`, `.state_dict()`,
`.to(device)` and the stock torch optimisers behave exactly as with the reference); their forward
methods are never called."""
import math
import os

import torch
import torch.nn as nn

from lagvae import DropoutSpec, LagvaeError, TextEngine

_ENGINES = {}
_CALLS = [0]


def get_engine(V, ni, nh, nz, device):
    device = torch.device(device)
    if device.type != "cuda":
        raise LagvaeError("modules.* run on a CUDA (B200) device only — no CPU fallback; got %s" % device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (V, ni, nh, nz, idx)
    if key not in _ENGINES:
        _ENGINES[key] = TextEngine(V, ni, nh, nz, torch.device("cuda", idx))
    return _ENGINES[key]


def _enc_params(enc):
    return [enc.embed.weight, enc.lstm.weight_ih_l0, enc.lstm.weight_hh_l0, enc.lstm.bias_ih_l0,
            enc.lstm.bias_hh_l0, enc.linear.weight]


def _dec_params(dec):
    return [dec.embed.weight, dec.trans_linear.weight, dec.lstm.weight_ih_l0, dec.lstm.weight_hh_l0,
            dec.lstm.bias_ih_l0, dec.lstm.bias_hh_l0, dec.pred_linear.weight]


def _detached(ps):
    return [p.detach() for p in ps]


def dropout_spec(dec, B, T, ns, device):
    """Dropout control for one decoder forward.  train(): in-kernel Philox (default) or, with
    LAGVAE_DROPOUT=torch, explicit Bernoulli keep-masks drawn from torch's generator in the reference's
    order (dropout_in then dropout_out, dec_lstm.py:81,106).  eval(): identity."""
    p_in, p_out = float(dec.dropout_in.p), float(dec.dropout_out.p)
    if not dec.training or (p_in == 0.0 and p_out == 0.0):
        return DropoutSpec()
    if os.environ.get("LAGVAE_DROPOUT", "philox") == "torch":
        m_in = (torch.rand(B, T - 1, dec.ni, device=device) >= p_in).to(torch.uint8) if p_in > 0 else None
        m_out = (torch.rand(B * ns, T - 1, dec.nh, device=device) >= p_out).to(torch.uint8) if p_out > 0 else None
        return DropoutSpec(1, p_in, p_out, m_in, m_out, 0)
    _CALLS[0] += 1
    seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + _CALLS[0] * 0xD1B54A32D192ED03) & (2 ** 64 - 1)
    return DropoutSpec(2, p_in, p_out, None, None, seed)


class GaussianEncoderBase(nn.Module):
    """Diagonal-Gaussian posterior maths shared by the encoders (reference encoder.py:7-145)."""

    def forward(self, x):
        raise NotImplementedError

    def reparameterize(self, mu, logvar, nsamples=1):
        """z = mu + eps*exp(.5 logvar), eps drawn with one normal_() call on a fresh [B,ns,nz] tensor
        (encoder.py:59-79; same generator consumption as the reference on this device)."""
        B, nz = mu.shape
        eps = torch.empty(B, nsamples, nz, dtype=mu.dtype, device=mu.device).normal_()
        return mu.unsqueeze(1) + eps * (0.5 * logvar).exp().unsqueeze(1)

    def sample(self, input, nsamples):
        mu, logvar = self.forward(input)
        return self.reparameterize(mu, logvar, nsamples), (mu, logvar)

    def encode(self, input, nsamples):
        mu, logvar = self.forward(input)
        z = self.reparameterize(mu, logvar, nsamples)
        KL = 0.5 * (mu.pow(2) + logvar.exp() - logvar - 1).sum(dim=1)   # encoder.py:55
        return z, KL

    def eval_inference_dist(self, x, z, param=None):
        """log q(z|x) for z [B,ns,nz] -> [B,ns] (encoder.py:81-109)."""
        nz = z.size(2)
        mu, logvar = param if param else self.forward(x)
        mu, logvar = mu.unsqueeze(1), logvar.unsqueeze(1)
        dev = z - mu
        return -0.5 * ((dev ** 2) / logvar.exp()).sum(dim=-1) - 0.5 * (nz * math.log(2 * math.pi) + logvar.sum(-1))

    def calc_mi(self, x):
        """MI estimate (encoder.py:111-145) -> Python float; all-pairs log-density + logsumexp run in the
        fused lagvae_mi_estimate kernel."""
        mu, logvar = self.forward(x)
        eps = torch.empty(mu.shape[0], 1, mu.shape[1], dtype=mu.dtype, device=mu.device).normal_()  # :128
        return self._engine_for(x).mi(mu, logvar, eps).item()


class LSTMEncoder(GaussianEncoderBase):
    """Gaussian LSTM encoder with constant-length batching (reference enc_lstm.py:10-64)."""

    def __init__(self, args, vocab_size, model_init, emb_init):
        super().__init__()
        self.ni, self.nh, self.nz = args.ni, args.enc_nh, args.nz
        self.vocab_size = vocab_size
        self.embed = nn.Embedding(vocab_size, args.ni)
        self.lstm = nn.LSTM(input_size=args.ni, hidden_size=args.enc_nh, num_layers=1, batch_first=True, dropout=0)
        self.linear = nn.Linear(args.enc_nh, 2 * args.nz, bias=False)
        self.reset_parameters(model_init, emb_init)

    def reset_parameters(self, model_init, emb_init):
        for param in self.parameters():      # every parameter incl. biases (enc_lstm.py:42-44)
            model_init(param)
        emb_init(self.embed.weight)

    def _engine_for(self, x):
        return get_engine(self.vocab_size, self.ni, self.nh, self.nz, self.embed.weight.device)

    def forward(self, input):
        """(mu, logvar), each [B,nz] (enc_lstm.py:47-64).  Inference-only entry: gradients flow through
        VAE.loss, which is the fused differentiable path."""
        eng = self._engine_for(input)
        return eng.encode_stats(_detached(_enc_params(self)) + [None] * 7, input)


class DecoderBase(nn.Module):
    """Abstract decoder API (reference decoder.py:5-73)."""

    def decode(self, x, z):
        raise NotImplementedError

    def reconstruct_error(self, x, z):
        raise NotImplementedError

    def beam_search_decode(self, z, K):
        raise NotImplementedError

    def sample_decode(self, z):
        raise NotImplementedError

    def greedy_decode(self, z):
        raise NotImplementedError

    def log_probability(self, x, z):
        raise NotImplementedError


class LSTMDecoder(DecoderBase):
    """LSTM decoder with constant-length batching (reference dec_lstm.py:17-161)."""

    def __init__(self, args, vocab, model_init, emb_init):
        super().__init__()
        self.ni, self.nh, self.nz = args.ni, args.dec_nh, args.nz
        self.vocab = vocab
        self.device = args.device
        V = len(vocab)
        self.embed = nn.Embedding(V, args.ni, padding_idx=-1)   # resolves to V-1: that row gets no gradient
        self.dropout_in = nn.Dropout(args.dec_dropout_in)
        self.dropout_out = nn.Dropout(args.dec_dropout_out)
        self.trans_linear = nn.Linear(args.nz, args.dec_nh, bias=False)
        self.lstm = nn.LSTM(input_size=args.ni + args.nz, hidden_size=args.dec_nh, num_layers=1, batch_first=True)
        self.pred_linear = nn.Linear(args.dec_nh, V, bias=False)
        # kept for state_dict compatibility (key `decoder.loss.weight`, dec_lstm.py:45-47)
        self.loss = nn.CrossEntropyLoss(weight=torch.ones(V), reduction="none")
        self.reset_parameters(model_init, emb_init)

    def reset_parameters(self, model_init, emb_init):
        for param in self.parameters():
            model_init(param)
        emb_init(self.embed.weight)

    def _engine(self):
        return get_engine(len(self.vocab), self.ni, self.nh, self.nz, self.embed.weight.device)

    def reconstruct_error(self, x, z):
        """Per-(sentence, sample) summed token cross entropy, [B, ns] (dec_lstm.py:113-148).
        Forward-only entry; training gradients flow through VAE.loss."""
        B, T = x.shape
        drop = dropout_spec(self, B, T, z.shape[1], x.device)
        return self._engine().reconstruct_error([None] * 6 + _detached(_dec_params(self)), x, z.detach(), drop)

    def log_probability(self, x, z):
        return -self.reconstruct_error(x, z)                      # dec_lstm.py:151-161

    def decode(self, input, z):
        raise NotImplementedError("materialised logits are never exposed by the fused decoder; "
                                  "use reconstruct_error / log_probability (SURVEY §8 f4: generation is out of scope)")

    def beam_search_decode(self, z, K=5):
        raise NotImplementedError("generation (dec_lstm.py:163-367) is out of scope of the hot path (SURVEY §8 f4)")

    greedy_decode = sample_decode = beam_search_decode
