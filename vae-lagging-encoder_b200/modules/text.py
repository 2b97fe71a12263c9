"""Drop-in text modules: LSTMEncoder / LSTMDecoder with the reference constructors, attribute names,
state_dict keys and method contracts (reference modules/encoders/enc_lstm.py:10-64,
modules/encoders/encoder.py:7-145, modules/decoders/dec_lstm.py:17-161, decoder.py:5-73), computing
on the B200 kernels of liblagvae.so through lagvae.TextEngine.  nn.Embedding / nn.LSTM / nn.Linear
objects are kept only as *parameter containers* (so that `.parameters()`, `.state_dict()`,
`.to(device)` and the stock torch optimisers behave exactly as with the reference); their forward
methods are never called."""
import ctypes as C
import math
import os

import torch
import torch.nn as nn

from lagvae import DropoutSpec, LagvaeError, TextEngine
from lagvae import _backend as be
from lagvae import graph as _graph

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


def dropout_spec(dec, B, T, ns, device, row_offset=0, rows_total=None):
    """Dropout control for one decoder forward on B rows.  train(): in-kernel Philox (default) or, with
    LAGVAE_DROPOUT=torch, explicit Bernoulli keep-masks drawn from torch's generator in the reference's
    order (dropout_in then dropout_out, dec_lstm.py:81,106).  eval(): identity.
    Under data parallelism the B rows are the shard [row_offset, row_offset + B) of a batch of rows_total sentences:
    torch masks are drawn for the whole batch (identical on every rank) and sliced; the Philox stream is keyed per shard
    ("per-rank Philox streams in fast mode; globally generated + sliced in parity mode")."""
    p_in, p_out = float(dec.dropout_in.p), float(dec.dropout_out.p)
    if not dec.training or (p_in == 0.0 and p_out == 0.0):
        return DropoutSpec()
    total = B if rows_total is None else int(rows_total)
    if os.environ.get("LAGVAE_DROPOUT", "philox") == "torch":
        lo, hi = int(row_offset), int(row_offset) + B
        m_in = (torch.rand(total, T - 1, dec.ni, device=device) >= p_in).to(torch.uint8)[lo:hi].contiguous() if p_in > 0 else None
        m_out = (torch.rand(total * ns, T - 1, dec.nh, device=device) >= p_out).to(torch.uint8)[lo * ns:hi * ns].contiguous() if p_out > 0 else None
        return DropoutSpec(1, p_in, p_out, m_in, m_out, 0)
    if device.type == "cuda" and torch.cuda.is_current_stream_capturing():
        # the host seed is frozen into the graph: key the masks by seed + a device word that the graph itself advances by
        # the step an eager call advances _CALLS by — replay k then draws the masks of the k-th eager call (lagvae/graph.py)
        word = _graph.philox_word(device)
        _graph.bump_philox_word(word)
        seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + _CALLS[0] * _graph.PHILOX_STEP + int(row_offset) * 0x2545F4914F6CDD1D) & (2 ** 64 - 1)
        return DropoutSpec(2, p_in, p_out, None, None, seed, word)
    _CALLS[0] += 1
    seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + _CALLS[0] * _graph.PHILOX_STEP + int(row_offset) * 0x2545F4914F6CDD1D) & (2 ** 64 - 1)
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
        """Logits of every position, [B*ns, T', V] (dec_lstm.py:66-111): input int64 [B, T'], z [B, ns, nz].
        Forward-only entry (the training path never materialises this tensor on the host side; gradients flow through
        VAE.loss)."""
        B, Td = input.shape
        drop = dropout_spec(self, B, Td + 1, z.shape[1], input.device)
        return self._engine().decode_logits([None] * 6 + _detached(_dec_params(self)), input, z.detach(), drop)

    # ---- generation (dec_lstm.py:163-367): host-driven token loops, exactly like the reference, over ONE
    #      decoder time step computed by liblagvae.so kernels (input projection GEMM, LSTM cell step, vocabulary projection)
    def _init_state(self, z):
        """(h0, c0) = (tanh(W_t z), W_t z) — dec_lstm.py:186-187."""
        n = z.shape[0]
        w = self.trans_linear.weight.detach()
        c0 = torch.empty(n, self.nh, dtype=torch.float32, device=z.device)
        L = be.lib()
        be.check(L.lagvae_gemm_f32(be.ptr(z), self.nz, 1, be.ptr(w), self.nz, 1, be.ptr(c0), self.nh, n, self.nh, self.nz, 1.0, 0.0,
                                   None, None, 0, _st()), "lagvae_gemm_f32")
        return torch.tanh(c0), c0

    def _step(self, ids, z_rows, h, c):
        """One decoder time step on n rows: (logits [n, V], h', c').  No dropout (the reference applies none here)."""
        _need_cuda_text(z_rows)
        L = be.lib()
        n, V = ids.shape[0], len(self.vocab)
        kin = self.ni + self.nz
        x_in = torch.cat([self.embed.weight.detach().index_select(0, ids), z_rows], dim=1).contiguous()
        w_ih, w_hh = self.lstm.weight_ih_l0.detach(), self.lstm.weight_hh_l0.detach()
        bias = (self.lstm.bias_ih_l0.detach() + self.lstm.bias_hh_l0.detach()).contiguous()
        gates = torch.empty(n, 4 * self.nh, dtype=torch.float32, device=ids.device)
        be.check(L.lagvae_gemm_f32(be.ptr(x_in), kin, 1, be.ptr(w_ih), kin, 1, be.ptr(gates), 4 * self.nh, n, 4 * self.nh, kin, 1.0, 0.0,
                                   be.ptr(bias), None, 0, _st()), "lagvae_gemm_f32")
        h2, c2 = torch.empty_like(h), torch.empty_like(c)
        drop = be.Dropout()
        be.check(L.lagvae_lstm_forward(0, self.nh, 1, n, be.ptr(w_hh), be.ptr(h.contiguous()), be.ptr(c.contiguous()), be.ptr(gates),
                                       be.ptr(c2), be.ptr(h2), None, C.byref(drop), None, 0, _st()), "lagvae_lstm_forward")
        logits = torch.empty(n, V, dtype=torch.float32, device=ids.device)
        w_p = self.pred_linear.weight.detach()
        be.check(L.lagvae_gemm_f32(be.ptr(h2), self.nh, 1, be.ptr(w_p), self.nh, 1, be.ptr(logits), V, n, V, self.nh, 1.0, 0.0,
                                   None, None, 0, _st()), "lagvae_gemm_f32")
        return logits, h2, c2

    def _token_loop(self, z, pick):
        """Shared body of greedy_decode / sample_decode (dec_lstm.py:266-367): a sentence keeps receiving words until (and
        including) its first </s>; at most 99 steps."""
        z = z.detach().float().contiguous()
        n = z.shape[0]
        h, c = self._init_state(z)
        ids = torch.full((n,), self.vocab["<s>"], dtype=torch.long, device=z.device)
        eos = self.vocab["</s>"]
        alive = torch.ones(n, dtype=torch.bool, device=z.device)
        decoded = [[] for _ in range(n)]
        length_c = 1
        while bool(alive.any()) and length_c < 100:
            logits, h, c = self._step(ids, z, h, c)
            ids = pick(logits)
            length_c += 1
            live, words = alive.tolist(), ids.tolist()
            for i in range(n):
                if live[i]:
                    decoded[i].append(self.vocab.id2word(words[i]))
            alive = alive & (ids != eos)
        return decoded

    def greedy_decode(self, z):
        return self._token_loop(z, lambda logits: torch.argmax(logits, dim=1))

    def sample_decode(self, z):
        return self._token_loop(z, lambda logits: torch.multinomial(torch.softmax(logits, dim=1), num_samples=1).squeeze(1))

    def beam_search_decode(self, z, K=5):
        """Beam search, sentence by sentence (dec_lstm.py:163-264): each step scores (live hypotheses x V) continuations,
        keeps the best K - #completed, retires those ending in </s>; returns the best hypothesis incl. the leading <s>."""
        z = z.detach().float().contiguous()
        V, bos, eos = len(self.vocab), self.vocab["<s>"], self.vocab["</s>"]
        h0, c0 = self._init_state(z)
        result = []
        for b in range(z.shape[0]):
            nodes = [(-1, bos)]                                   # (parent node, token)
            live, live_lp = [0], [0.0]
            h, c = h0[b:b + 1], c0[b:b + 1]
            done = []
            t = 0
            while len(done) < K and t < 100:
                t += 1
                ids = torch.tensor([nodes[i][1] for i in live], dtype=torch.long, device=z.device)
                logits, h2, c2 = self._step(ids, z[b:b + 1].expand(len(live), -1).contiguous(), h, c)
                score = torch.log_softmax(logits, dim=-1) + torch.tensor(live_lp, dtype=torch.float32, device=z.device).view(-1, 1)
                top_lp, top_ix = torch.topk(score.reshape(-1), K - len(done))
                new_live, new_lp, rows = [], [], []
                for lp, ix in zip(top_lp.tolist(), top_ix.tolist()):
                    li, w = ix // V, ix % V
                    nodes.append((live[li], w))
                    if w == eos:
                        done.append((lp, len(nodes) - 1))
                    else:
                        new_live.append(len(nodes) - 1)
                        new_lp.append(lp)
                        rows.append(li)
                live, live_lp = new_live, new_lp
                if not live:
                    break
                sel = torch.tensor(rows, dtype=torch.long, device=z.device)
                h, c = h2.index_select(0, sel), c2.index_select(0, sel)
            done += list(zip(live_lp, live))
            best = max(done, key=lambda d: d[0])[1]
            words = []
            while best >= 0:
                words.append(self.vocab.id2word(nodes[best][1]))
                best = nodes[best][0]
            result.append(words[::-1])
        return result


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda_text(t):
    if t.device.type != "cuda":
        raise LagvaeError("modules.* run on a CUDA (B200) device only — no CPU fallback; got %s" % t.device)
