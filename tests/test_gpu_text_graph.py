"""The text inner loop as a CUDA graph ("15-step windows in a CUDA graph"; text.py:371-400 repeats the same
statement sequence on same-shaped batches and only reads the ACCUMULATED Σloss every 15 steps, text.py:389-398).

* a window of fused inner steps (lagvae_text_inner_step) captured once by lagvae.GraphedStep and replayed equals the same
  steps run eagerly — train()-mode in-kernel Philox dropout included: the captured step keys its masks by
  `seed + *seed_dev` (include/lagvae.h) and advances the device word inside the graph, so replay k draws the masks of the
  k-th eager call; a decoder update between two replays is picked up (the first captured step re-splits the decoder
  weights at every replay, the rest of the window reuses them);
* the module-level sequence (`vae.loss` -> backward) captured in train() mode draws fresh masks per replay and each replay
  equals the eager call with the corresponding seed."""
import os
import sys
import types

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (os.path.join(ROOT, "vae-lagging-encoder_b200"), os.path.join(ROOT, "oracle")):
    if p_ not in sys.path:
        sys.path.insert(0, p_)

SHAPES = [(520, 64, 256, 8, 8, 12), (20001, 512, 1024, 32, 32, 24)]   # (V, ni, nh, nz, B, T); the second runs k_lstm_v2


def _params(shape, seed=5):
    import lagging_oracle as O
    V, ni, nh, nz, B, T = shape
    p = O.scale_trained_like(O.init_text_params(V, ni, nh, nz, seed=seed), 4.0)
    return [p[k].cuda().contiguous() for k in O.ALL_KEYS]


@pytest.mark.parametrize("shape", SHAPES, ids=["small", "yahoo-dims-T24"])
def test_graphed_window_of_fused_inner_steps_equals_eager(shape):
    import lagvae
    import lagging_oracle as O
    from lagvae import graph as G
    V, ni, nh, nz, B, T = shape
    W, NWIN = 3, 2                                        # steps per graph, replays
    dev = torch.device("cuda")
    eng = lagvae.TextEngine(V, ni, nh, nz, "cuda")
    init = _params(shape)
    gen = torch.Generator().manual_seed(77)
    xs = [O.make_token_batch(B, T, V, seed=100 + k).cuda() for k in range(W * NWIN)]
    es = [torch.randn(B, 1, nz, generator=gen).cuda() for _ in range(W * NWIN)]
    base = 0x1234567890ABCDEF
    klw = 0.37

    # ---- eager: K steps, host seed advanced by PHILOX_STEP per step; the decoder is perturbed after the first window
    pe = [t.clone() for t in init]
    gw = eng.grad_workspace()
    ol = torch.empty(B, device=dev)
    want_sc, want_loss = [], []
    for k in range(W * NWIN):
        if k == W:
            for t in pe[6:]:
                t.mul_(1.03)
        sc = torch.empty(4, device=dev)
        seed = (base + (k + 1) * G.PHILOX_STEP) & (2 ** 64 - 1)
        eng.inner_step(pe, xs[k], es[k], klw, lagvae.DropoutSpec(2, 0.5, 0.5, None, None, seed), gw, ol, sc)
        want_sc.append(sc.clone())
        want_loss.append(ol.clone())
    torch.cuda.synchronize()

    # ---- graph: one capture of W steps
    pg = [t.clone() for t in init]
    word = G.philox_word(dev)
    gw2 = eng.grad_workspace()
    out_l = torch.empty(W, B, device=dev)
    out_s = torch.empty(W, 4, device=dev)

    def body(**kw):
        for i in range(W):
            G.bump_philox_word(word)
            eng.inner_step(pg, kw["x%d" % i], kw["e%d" % i], klw, lagvae.DropoutSpec(2, 0.5, 0.5, None, None, base, word),
                           gw2, out_l[i], out_s[i])
        return out_s, out_l

    ex = {}
    for i in range(W):
        ex["x%d" % i], ex["e%d" % i] = xs[i], es[i]
    gs = lagvae.GraphedStep(body, ex, warmup=2)
    # the warm-up ran real steps: restore the starting point (in place: the graph holds the pointers)
    for t, s in zip(pg, init):
        t.copy_(s)
    word.zero_()
    for wdw in range(NWIN):
        if wdw == 1:
            for t in pg[6:]:
                t.mul_(1.03)                               # decoder update between two replays (text.py:424)
        kw = {}
        for i in range(W):
            kw["x%d" % i], kw["e%d" % i] = xs[wdw * W + i], es[wdw * W + i]
        s, l = gs(**kw)
        torch.cuda.synchronize()
        for i in range(W):
            k = wdw * W + i
            # same kernels in the same order; what differs is the summation order of the atomics (embedding scatter-add,
            # split-K), amplified step by step through the lr-1.0 updates: 1e-4, against O(1) for a wrong mask or weight
            assert torch.allclose(l[i], want_loss[k], rtol=1e-4, atol=1e-4), (k, (l[i] - want_loss[k]).abs().max())
            assert torch.allclose(s[i], want_sc[k], rtol=1e-4, atol=1e-5), (k, s[i], want_sc[k])
    for a, b in zip(pg, pe):
        assert float((a - b).abs().max()) <= 1e-3 * max(float(b.abs().max()), 1e-6)   # six lr-1.0 steps of atomics-order drift
    # fresh masks per step: two steps on different masks cannot produce the same reconstruction sum
    assert float(want_sc[0][1]) != float(want_sc[1][1])
    if nh >= 256 and B <= 32:
        assert lagvae.lstm_variant()["forward"].startswith("v2")   # the cooperative cluster kernels were captured


class _Vocab(dict):
    def __init__(self, V):
        super().__init__()
        self.V = V
        self["<s>"], self["</s>"] = 1, 2

    def __len__(self):
        return self.V

    def id2word(self, i):
        return str(i)


def test_graphed_module_loss_in_train_mode_draws_fresh_masks():
    import modules
    import lagvae
    import lagging_oracle as O
    from modules import text as MT
    V, ni, nh, nz, B, T = SHAPES[0]
    dev = torch.device("cuda")
    a = types.SimpleNamespace(ni=ni, enc_nh=nh, dec_nh=nh, nz=nz, dec_dropout_in=0.5, dec_dropout_out=0.5, device=dev)
    init = lambda t: torch.nn.init.uniform_(t, -0.01, 0.01)
    vae = modules.VAE(modules.LSTMEncoder(a, V, init, init), modules.LSTMDecoder(a, _Vocab(V), init, init), a).to(dev)
    p = O.scale_trained_like(O.init_text_params(V, ni, nh, nz, seed=9), 4.0)
    sd = vae.state_dict()
    sd.update({k: p[k].to(dev) for k in O.ALL_KEYS})
    vae.load_state_dict(sd)
    vae.train()
    x = O.make_token_batch(B, T, V, seed=3).to(dev)
    params = list(vae.parameters())
    eps = torch.randn(B, 1, nz, device=dev)

    def eager():
        for q in params:
            q.grad = None
        loss, rec, kl = vae.loss(x, 0.5, nsamples=1)
        loss.mean(dim=-1).backward()
        return loss.detach().clone(), [q.grad.detach().clone() for q in params]

    static_g = [torch.zeros_like(q) for q in params]

    def body(x):
        for q in params:
            q.grad = None
        loss, rec, kl = vae.loss(x, 0.5, nsamples=1)
        loss.mean(dim=-1).backward()
        for s, q in zip(static_g, params):
            s.copy_(q.grad)
        return loss.detach()

    # eps inside a captured vae.loss comes from torch's graph-safe generator: compare on the reconstruction masks through a
    # fixed eps instead — patch the draw for this test only
    word = lagvae.graph.philox_word(dev)
    orig = torch.Tensor.normal_
    try:
        torch.Tensor.normal_ = lambda self, *a_, **k_: self.copy_(eps.view_as(self)) if self.shape[-1] == nz else orig(self, *a_, **k_)
        gs = lagvae.GraphedStep(body, {"x": x}, warmup=2)
        c0 = MT._CALLS[0]
        word.zero_()
        outs = []
        for j in range(3):
            l = gs(x=x).clone()
            torch.cuda.synchronize()
            outs.append((l, [g.clone() for g in static_g]))
        assert not torch.equal(outs[0][0], outs[1][0]) and not torch.equal(outs[1][0], outs[2][0])   # fresh masks per replay
        for j in range(3):
            MT._CALLS[0] = c0 + j                        # the eager call increments to c0 + j + 1 = replay j+1's key
            l, gr = eager()
            assert torch.allclose(l, outs[j][0], rtol=1e-6, atol=1e-6), (j, (l - outs[j][0]).abs().max())
            for a_, b_ in zip(gr, outs[j][1]):
                assert float((a_ - b_).abs().max()) <= 1e-5 * max(float(b_.abs().max()), 1e-8)   # atomics order only
    finally:
        torch.Tensor.normal_ = orig
