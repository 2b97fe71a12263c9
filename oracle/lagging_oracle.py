"""ORACLE — test infrastructure only (NOT product code).

CPU restatement (plain torch tensor ops, explicit LSTM cell loop, fp32 or fp64) of the
aggressive-inner-loop hot path of jxhe/vae-lagging-encoder.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this file; the product path (`vae-lagging-encoder_b200/`) never does.

Parity status: PINNED.  The reference ships no golden vectors; this
restatement is validated here against the *unmodified reference modules imported from
/root/reference* by `oracle/validate_against_reference.py`, which also writes the fixtures in
`tests/golden/` (script committed).  `tests/test_oracle_golden.py` re-checks the oracle
against those fixtures on every run (no access to /root/reference needed).

Every function cites the reference file:line it restates (paths relative to the reference
root).  Parameter names are the reference `state_dict` keys.
"""
from __future__ import annotations

import math
from typing import List,  Dict, Optional, Tuple

import numpy as np
import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]

ENC_KEYS = [
    "encoder.embed.weight", "encoder.lstm.weight_ih_l0", "encoder.lstm.weight_hh_l0",
    "encoder.lstm.bias_ih_l0", "encoder.lstm.bias_hh_l0", "encoder.linear.weight",
]
DEC_KEYS = [
    "decoder.embed.weight", "decoder.trans_linear.weight", "decoder.lstm.weight_ih_l0",
    "decoder.lstm.weight_hh_l0", "decoder.lstm.bias_ih_l0", "decoder.lstm.bias_hh_l0",
    "decoder.pred_linear.weight",
]
ALL_KEYS = ENC_KEYS + DEC_KEYS  # order of vae.parameters() in the reference (encoder first)


# ----------------------------------------------------------------------------------------------
# parameter construction (text.py:232-241,265-266; enc_lstm.py:31-44; dec_lstm.py:51-64)
# ----------------------------------------------------------------------------------------------
def init_text_params(V: int, ni: int, nh: int, nz: int, seed: int = 0,
                     dtype=torch.float32) -> Params:
    """Reference initialisation: every parameter U(-0.01,0.01), then embeddings U(-0.1,0.1)
    (enc_lstm.py:42-44, dec_lstm.py:62-64, text.py:265-266).  Draw order follows the module
    construction order of text.py:271-277 only loosely — tests never rely on bit-equality with
    a reference-constructed model; they always copy a state_dict."""
    g = torch.Generator().manual_seed(seed)
    shapes = {
        "encoder.embed.weight": (V, ni),
        "encoder.lstm.weight_ih_l0": (4 * nh, ni),
        "encoder.lstm.weight_hh_l0": (4 * nh, nh),
        "encoder.lstm.bias_ih_l0": (4 * nh,),
        "encoder.lstm.bias_hh_l0": (4 * nh,),
        "encoder.linear.weight": (2 * nz, nh),
        "decoder.embed.weight": (V, ni),
        "decoder.trans_linear.weight": (nh, nz),
        "decoder.lstm.weight_ih_l0": (4 * nh, ni + nz),
        "decoder.lstm.weight_hh_l0": (4 * nh, nh),
        "decoder.lstm.bias_ih_l0": (4 * nh,),
        "decoder.lstm.bias_hh_l0": (4 * nh,),
        "decoder.pred_linear.weight": (V, nh),
    }
    p = {}
    for k in ALL_KEYS:
        a = 0.1 if k.endswith("embed.weight") else 0.01
        p[k] = ((torch.rand(shapes[k], generator=g, dtype=torch.float64) * 2 - 1) * a).to(dtype)
    return p


def scale_trained_like(p: Params, gain: float = 4.0) -> Params:
    """In-place re-scale of the reference's U(+-0.01)/U(+-0.1) init to 'trained-like' magnitudes
    (embeddings U(+-0.5), matrices U(+-gain/sqrt(fan_in)), biases U(+-0.1)) so that every tensor
    on the path has O(1) sensitivity.  At the stock init logits ~ 0 and KL ~ 1e-5 (fp32 KL itself
    is ill-conditioned there — SURVEY §7 hard part 3), so a wrong dropout mask or a broken
    recurrence would be invisible in the loss.  Test-fixture helper, not reference behaviour."""
    with torch.no_grad():
        for name, q in p.items():
            if name.endswith("embed.weight"):
                q.mul_(5.0)
            elif q.dim() == 2:
                q.mul_((gain / q.shape[1] ** 0.5) / 0.01)
            else:
                q.mul_(10.0)
    return p


def make_token_batch(B: int, T: int, V: int, seed: int = 1234) -> Tensor:
    """Synthetic ids of SURVEY §8(d2): uniform on [4,V), col 0 = <s>(1), last col = </s>(2)
    (special ids: data/text_data.py:19-22)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(4, V, (B, T), generator=g, dtype=torch.int64)
    x[:, 0] = 1
    x[:, -1] = 2
    return x


# ----------------------------------------------------------------------------------------------
# LSTM cell recurrence (torch.nn.LSTM semantics as used at enc_lstm.py:20-24,60 and
# dec_lstm.py:37-40,104; gate order i,f,g,o — SURVEY Appendix A.1)
# ----------------------------------------------------------------------------------------------
def lstm_sequence(x_in: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor,
                  h0: Optional[Tensor] = None, c0: Optional[Tensor] = None
                  ) -> Tuple[Tensor, Tensor, Tensor]:
    """x_in [B,T,K] -> (all h [B,T,nh], h_T [B,nh], c_T [B,nh])."""
    B, T, _ = x_in.shape
    nh = w_hh.shape[1]
    h = x_in.new_zeros(B, nh) if h0 is None else h0
    c = x_in.new_zeros(B, nh) if c0 is None else c0
    pre = x_in @ w_ih.t() + (b_ih + b_hh)          # input projection for all steps
    outs = []
    for t in range(T):
        a = pre[:, t] + h @ w_hh.t()
        i, f, g, o = a.split(nh, dim=1)
        i, f, o = torch.sigmoid(i), torch.sigmoid(f), torch.sigmoid(o)
        g = torch.tanh(g)
        c = f * c + i * g
        h = o * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, dim=1), h, c


# ----------------------------------------------------------------------------------------------
# encoder (modules/encoders/enc_lstm.py:47-64) and Gaussian posterior maths
# (modules/encoders/encoder.py:40-79)
# ----------------------------------------------------------------------------------------------
def encoder_forward(p: Params, x: Tensor) -> Tuple[Tensor, Tensor]:
    """enc_lstm.py:58-64: embed -> LSTM (h0=c0=0) -> Linear(nh,2nz,bias=False) -> chunk."""
    emb = p["encoder.embed.weight"][x]                                    # :58
    _, h_last, _ = lstm_sequence(emb, p["encoder.lstm.weight_ih_l0"], p["encoder.lstm.weight_hh_l0"],
                                 p["encoder.lstm.bias_ih_l0"], p["encoder.lstm.bias_hh_l0"])  # :60
    ml = h_last @ p["encoder.linear.weight"].t()                          # :62
    nz = ml.shape[1] // 2
    return ml[:, :nz], ml[:, nz:]


def reparameterize(mu: Tensor, logvar: Tensor, eps: Tensor) -> Tensor:
    """encoder.py:59-79 with the N(0,1) draw `eps` [B,ns,nz] made explicit (drawn at :77)."""
    std = (0.5 * logvar).exp()
    return mu.unsqueeze(1) + eps * std.unsqueeze(1)


def kl_gaussian(mu: Tensor, logvar: Tensor) -> Tensor:
    """encoder.py:55."""
    return 0.5 * (mu.pow(2) + logvar.exp() - logvar - 1).sum(dim=1)


# ----------------------------------------------------------------------------------------------
# decoder (modules/decoders/dec_lstm.py:66-148)
# ----------------------------------------------------------------------------------------------
def decoder_reconstruct_error(p: Params, x: Tensor, z: Tensor,
                              mask_in: Optional[Tensor] = None,
                              mask_out: Optional[Tensor] = None) -> Tensor:
    """dec_lstm.py:113-148 (+decode 66-111).  z [B,ns,nz]; returns [B,ns].
    mask_in [B,T-1,ni] / mask_out [B*ns,T-1,nh] are the *scaled* inverted-dropout masks
    (values 0 or 1/(1-p)); None = eval mode (identity)."""
    src, tgt = x[:, :-1], x[:, 1:]                                        # :124,127
    B, Tm = src.shape
    ns, nz = z.shape[1], z.shape[2]
    emb = p["decoder.embed.weight"][src]                                  # :80
    if mask_in is not None:
        emb = emb * mask_in                                               # :81
    if ns == 1:
        z_ = z.expand(B, Tm, nz)                                          # :84
    else:
        emb = emb.unsqueeze(1).expand(B, ns, Tm, emb.shape[-1]).reshape(B * ns, Tm, -1)  # :87-91
        z_ = z.unsqueeze(2).expand(B, ns, Tm, nz).reshape(B * ns, Tm, nz)               # :93-94
    inp = torch.cat((emb, z_), -1)                                        # :97
    zf = z.reshape(B * ns, nz)
    c0 = zf @ p["decoder.trans_linear.weight"].t()                        # :100
    h0 = torch.tanh(c0)                                                   # :101
    out, _, _ = lstm_sequence(inp, p["decoder.lstm.weight_ih_l0"], p["decoder.lstm.weight_hh_l0"],
                              p["decoder.lstm.bias_ih_l0"], p["decoder.lstm.bias_hh_l0"], h0, c0)  # :104
    if mask_out is not None:
        out = out * mask_out                                              # :106
    logits = out @ p["decoder.pred_linear.weight"].t()                    # :109
    tg = tgt if ns == 1 else tgt.unsqueeze(1).expand(B, ns, Tm)           # :135-140
    lse = torch.logsumexp(logits, dim=-1)
    picked = logits.gather(-1, tg.reshape(B * ns, Tm, 1)).squeeze(-1)
    loss = lse - picked                                                   # CrossEntropy(reduce=False) :47,143
    return loss.view(B, ns, Tm).sum(-1)                                   # :148


# ----------------------------------------------------------------------------------------------
# VAE.loss (modules/vae.py:79-98)
# ----------------------------------------------------------------------------------------------
def vae_loss(p: Params, x: Tensor, kl_weight: float, eps: Tensor,
             mask_in: Optional[Tensor] = None, mask_out: Optional[Tensor] = None
             ) -> Tuple[Tensor, Tensor, Tensor]:
    mu, logvar = encoder_forward(p, x)
    z = reparameterize(mu, logvar, eps)                                   # vae.py:92 -> encoder.py:40-57
    KL = kl_gaussian(mu, logvar)
    rec = decoder_reconstruct_error(p, x, z, mask_in, mask_out).mean(dim=1)  # vae.py:95
    return rec + kl_weight * KL, rec, KL                                  # vae.py:98


# ----------------------------------------------------------------------------------------------
# MI estimate (modules/encoders/encoder.py:111-145; modules/utils.py:3-16)
# ----------------------------------------------------------------------------------------------
def calc_mi_from_stats(mu: Tensor, logvar: Tensor, eps: Tensor) -> float:
    """eps [B,1,nz] is the draw of encoder.py:128 (via reparameterize)."""
    B, nz = mu.shape
    neg_entropy = (-0.5 * nz * math.log(2 * math.pi) - 0.5 * (1 + logvar).sum(-1)).mean()  # :125
    z = reparameterize(mu, logvar, eps)                                   # :128   [B,1,nz]
    mu_, lv_ = mu.unsqueeze(0), logvar.unsqueeze(0)                       # :131
    var = lv_.exp()
    dev = z - mu_                                                         # :135   [B,B,nz]
    log_density = -0.5 * ((dev ** 2) / var).sum(dim=-1) - \
        0.5 * (nz * math.log(2 * math.pi) + lv_.sum(-1))                  # :138-139
    log_qz = torch.logsumexp(log_density, dim=1) - math.log(B)            # :143 (utils.py:3-16)
    return (neg_entropy - log_qz.mean(-1)).item()                         # :145


def calc_mi(p: Params, x: Tensor, eps: Tensor) -> float:
    mu, logvar = encoder_forward(p, x)
    return calc_mi_from_stats(mu, logvar, eps)


# ----------------------------------------------------------------------------------------------
# one aggressive inner step (text.py:371-391): zero_grad, loss, sum, mean.backward, clip, SGD(enc)
# ----------------------------------------------------------------------------------------------
def clip_coef(total_norm: float, max_norm: float = 5.0) -> float:
    """torch.nn.utils.clip_grad_norm_ (text.py:385, clip_grad text.py:17): coef clamped to 1."""
    return min(1.0, max_norm / (total_norm + 1e-6))


def inner_step(p: Params, x: Tensor, kl_weight: float, eps: Tensor,
               mask_in: Optional[Tensor] = None, mask_out: Optional[Tensor] = None,
               lr: float = 1.0, max_norm: float = 5.0, update: bool = True):
    """Returns dict(loss[B], rec[B], kl[B], loss_sum, grads{all 13}, grad_norm, coef) and,
    when update=True, applies p_enc -= lr*coef*g in place (SGD momentum 0, text.py:325,387).
    Decoder grads are computed (they enter the norm, text.py:385) and NOT applied (text.py:387)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    loss, rec, kl = vae_loss(leaves, x, kl_weight, eps, mask_in, mask_out)
    loss.mean(dim=-1).backward()                                          # text.py:382-384
    grads = {k: (leaves[k].grad if leaves[k].grad is not None else torch.zeros_like(leaves[k]))
             for k in ALL_KEYS}
    # decoder embedding: padding_idx=-1 -> row V-1 receives no gradient (dec_lstm.py:28)
    grads["decoder.embed.weight"][-1].zero_()
    total = math.sqrt(sum(float(g.double().pow(2).sum()) for g in grads.values()))
    coef = clip_coef(total, max_norm)
    if update:
        with torch.no_grad():
            for k in ENC_KEYS:
                p[k] -= lr * coef * grads[k]
    return {"loss": loss.detach(), "rec": rec.detach(), "kl": kl.detach(),
            "loss_sum": float(loss.detach().sum()), "grads": grads, "grad_norm": total, "coef": coef}


# ----------------------------------------------------------------------------------------------
# inner-loop control (text.py:366-400): which sub-iterations run, when the loop breaks
# ----------------------------------------------------------------------------------------------
class BurnWindow:
    """Host-side restatement of the convergence rule of text.py:366-400 (window 15, per word) /
    image.py:295-327 (window 10, per example)."""

    def __init__(self, window: int = 15, max_sub_iter: int = 100):
        self.window, self.max_sub_iter = window, max_sub_iter
        self.sub_iter = 1
        self.pre = 1e4
        self.cur = 0.0
        self.den = 0

    def keep_going(self) -> bool:
        return self.sub_iter < self.max_sub_iter                          # text.py:371

    def update(self, loss_sum: float, denom: int) -> bool:
        """Feed Σloss of the step and its word (or example) count; returns True if loop breaks."""
        self.cur += loss_sum                                              # :381
        self.den += denom                                                 # :377
        if self.sub_iter % self.window == 0:                              # :393
            cur = self.cur / self.den
            if self.pre - cur < 0:                                        # :395
                return True
            self.pre, self.cur, self.den = cur, 0.0, 0                    # :397-398
        self.sub_iter += 1                                                # :400
        return False


def next_batch_index(rng: np.random.RandomState, nbatch: int) -> int:
    """text.py:389 `np.random.random_integers(0, nbatch-1)` (inclusive upper bound)."""
    return int(rng.randint(0, nbatch))  # randint high is exclusive -> same support & stream


# ----------------------------------------------------------------------------------------------
# "fast port": the same path expressed with torch.nn modules (the arithmetic library the reference
# itself calls — SURVEY §8(c2)); used ONLY for timing the CPU baseline in bench.py.
# ----------------------------------------------------------------------------------------------
def nll_iw(p: Params, x: Tensor, eps_chunks: List[Tensor]) -> Tensor:
    """VAE.nll_iw (modules/vae.py:100-129): importance-weighted estimate of -log p(x), [B].  eps_chunks =
    the successive N(0,1) draws [B, ns, nz] of encoder.sample (encoder.py:59-79), one per chunk of ns samples; weights
    log p(z) + log p(x|z) - log q(z|x) (vae.py:147-168, encoder.py:81-109), log-sum-exp over all samples - log(nsamples)."""
    mu, logvar = encoder_forward(p, x)
    nz = mu.shape[1]
    parts = []
    for eps in eps_chunks:
        z = reparameterize(mu, logvar, eps)                                   # [B, ns, nz]
        log_prior = (-0.5 * z.pow(2) - 0.5 * math.log(2 * math.pi)).sum(dim=-1)
        log_lik = -decoder_reconstruct_error(p, x, z)                         # [B, ns]
        dev = z - mu.unsqueeze(1)
        log_q = -0.5 * (dev.pow(2) / logvar.exp().unsqueeze(1)).sum(dim=-1) \
            - 0.5 * (nz * math.log(2 * math.pi) + logvar.sum(dim=-1, keepdim=True))
        parts.append(log_prior + log_lik - log_q)
    w = torch.cat(parts, dim=-1)
    return -(torch.logsumexp(w, dim=-1) - math.log(w.shape[-1]))


# ------------------------------------------------------------------------------------------------------
# generation — host-driven token loops of dec_lstm.py:163-367, restated on the explicit cell
# ------------------------------------------------------------------------------------------------------
def decoder_step(p: Params, ids: Tensor, z_rows: Tensor, h: Tensor, c: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """One decoder time step on n rows (dec_lstm.py:208-216 / 291-298): logits [n, V], h', c'.  No dropout on these
    paths (the reference applies neither dropout_in nor dropout_out while generating)."""
    x_in = torch.cat([p["decoder.embed.weight"][ids], z_rows], dim=1)
    a = x_in @ p["decoder.lstm.weight_ih_l0"].t() + p["decoder.lstm.bias_ih_l0"] + h @ p["decoder.lstm.weight_hh_l0"].t() \
        + p["decoder.lstm.bias_hh_l0"]
    i, f, g, o = a.chunk(4, dim=1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    return h2 @ p["decoder.pred_linear.weight"].t(), h2, c2


def decoder_init_state(p: Params, z: Tensor) -> Tuple[Tensor, Tensor]:
    c0 = z @ p["decoder.trans_linear.weight"].t()                 # dec_lstm.py:186 / 278
    return torch.tanh(c0), c0


def greedy_decode(p: Params, z: Tensor, bos: int = 1, eos: int = 2, max_len: int = 100) -> List[List[int]]:
    """LSTMDecoder.greedy_decode (dec_lstm.py:266-314) as token ids: argmax per step, a sentence keeps receiving tokens
    until (and including) its first </s>; at most max_len - 1 steps."""
    n = z.shape[0]
    h, c = decoder_init_state(p, z)
    ids = torch.full((n,), bos, dtype=torch.long)
    alive = [True] * n
    out: List[List[int]] = [[] for _ in range(n)]
    steps = 1
    while any(alive) and steps < max_len:
        logits, h, c = decoder_step(p, ids, z, h, c)
        ids = logits.argmax(dim=1)
        steps += 1
        for b in range(n):
            if alive[b]:
                out[b].append(int(ids[b]))
                alive[b] = int(ids[b]) != eos
    return out


def beam_search_decode(p: Params, z: Tensor, K: int = 5, bos: int = 1, eos: int = 2, max_steps: int = 100) -> List[List[int]]:
    """LSTMDecoder.beam_search_decode (dec_lstm.py:163-264), sentence by sentence: every step scores
    (live hypotheses x V) continuations, keeps the best K - #completed, moves those ending in </s> to the completed set;
    returns the best-scoring hypothesis including the leading <s>."""
    V = p["decoder.pred_linear.weight"].shape[0]
    res: List[List[int]] = []
    h0, c0 = decoder_init_state(p, z)
    for b in range(z.shape[0]):
        nodes: List[Tuple[int, int]] = [(-1, bos)]            # (parent node, token)
        live, live_lp = [0], [0.0]
        h, c = h0[b:b + 1], c0[b:b + 1]
        done: List[Tuple[float, int]] = []
        t = 0
        while len(done) < K and t < max_steps:
            t += 1
            ids = torch.tensor([nodes[i][1] for i in live], dtype=torch.long)
            logits, h2, c2 = decoder_step(p, ids, z[b:b + 1].expand(len(live), -1), h, c)
            score = torch.log_softmax(logits, dim=-1) + torch.tensor(live_lp, dtype=torch.float32).view(-1, 1)
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
            h, c = h2[rows], c2[rows]
        done += list(zip(live_lp, live))
        best = max(done, key=lambda d: d[0])[1]
        toks = []
        while best >= 0:
            toks.append(nodes[best][1])
            best = nodes[best][0]
        res.append(toks[::-1])
    return res


class FastPort(torch.nn.Module):
    """Module-based port of modules/vae.py + enc_lstm.py + dec_lstm.py with the reference's own
    layer types (nn.Embedding / nn.LSTM / nn.Linear / CrossEntropyLoss) so that a CPU (or
    cuDNN) timing of it is representative of the reference's stock path."""

    def __init__(self, V, ni, nh, nz, p_in=0.5, p_out=0.5):
        super().__init__()
        nn = torch.nn
        self.nz = nz
        self.e_embed = nn.Embedding(V, ni)
        self.e_lstm = nn.LSTM(ni, nh, 1, batch_first=True)
        self.e_lin = nn.Linear(nh, 2 * nz, bias=False)
        self.d_embed = nn.Embedding(V, ni, padding_idx=-1)
        self.d_in, self.d_out = nn.Dropout(p_in), nn.Dropout(p_out)
        self.d_trans = nn.Linear(nz, nh, bias=False)
        self.d_lstm = nn.LSTM(ni + nz, nh, 1, batch_first=True)
        self.d_pred = nn.Linear(nh, V, bias=False)
        for q in self.parameters():
            torch.nn.init.uniform_(q, -0.01, 0.01)
        torch.nn.init.uniform_(self.e_embed.weight, -0.1, 0.1)
        torch.nn.init.uniform_(self.d_embed.weight, -0.1, 0.1)

    def enc_params(self):
        return list(self.e_embed.parameters()) + list(self.e_lstm.parameters()) + list(self.e_lin.parameters())

    def loss(self, x, kl_weight):
        _, (h, _) = self.e_lstm(self.e_embed(x))
        mu, lv = self.e_lin(h).chunk(2, -1)
        mu, lv = mu.squeeze(0), lv.squeeze(0)
        z = mu + torch.zeros_like(mu).normal_() * (0.5 * lv).exp()
        KL = 0.5 * (mu.pow(2) + lv.exp() - lv - 1).sum(1)
        src, tgt = x[:, :-1], x[:, 1:]
        B, Tm = src.shape
        we = self.d_in(self.d_embed(src))
        inp = torch.cat((we, z.unsqueeze(1).expand(B, Tm, self.nz)), -1)
        c0 = self.d_trans(z).unsqueeze(0)
        out, _ = self.d_lstm(inp, (torch.tanh(c0), c0))
        logits = self.d_pred(self.d_out(out))
        ce = torch.nn.functional.cross_entropy(logits.view(B * Tm, -1), tgt.reshape(-1), reduction="none")
        rec = ce.view(B, Tm).sum(-1)
        return rec + kl_weight * KL, rec, KL

    def inner_step(self, x, kl_weight, lr=1.0, max_norm=5.0):
        """One iteration of text.py:371-391 (stock torch ops)."""
        for q in self.parameters():
            q.grad = None
        loss, _, _ = self.loss(x, kl_weight)
        s = loss.sum().item()
        loss.mean(dim=-1).backward()
        torch.nn.utils.clip_grad_norm_(self.parameters(), max_norm)
        with torch.no_grad():
            for q in self.enc_params():
                q -= lr * q.grad
        return s
