"""The oracle (oracle/lagging_oracle.py) re-checked against fixtures produced by the UNMODIFIED
reference (oracle/validate_against_reference.py) — CPU only, no /root/reference needed."""
import numpy as np
import pytest
import torch

import lagging_oracle as O
from util import FULL_CASES, assert_close, case_inputs, case_params


@pytest.mark.parametrize("name", FULL_CASES)
def test_inner_step_matches_reference(golden, name):
    g = golden(name)
    c = case_inputs(g)
    p = case_params(g)
    r = O.inner_step(p, c["x"], c["klw"], c["eps"], c["mask_in"], c["mask_out"], update=True)
    assert_close(r["loss"], g["loss"], 1e-5, "loss")
    assert_close(r["rec"], g["rec"], 1e-5, "rec")
    assert_close(r["kl"], g["kl"], 1e-5, "kl", floor=1e-3)
    for k in O.ALL_KEYS:
        assert_close(r["grads"][k], g["g." + k], 2e-4, "grad " + k, floor=1e-8)
    assert abs(r["grad_norm"] - float(g["grad_norm"])) <= 1e-4 * float(g["grad_norm"])
    for k in O.ENC_KEYS:
        assert_close(p[k], g["post." + k], 1e-6, "post-step " + k)
    # padding_idx row of the decoder embedding receives no gradient (dec_lstm.py:28)
    assert float(torch.from_numpy(g["g.decoder.embed.weight"])[-1].abs().max()) == 0.0


@pytest.mark.parametrize("name", FULL_CASES + ["aligned_ns3_eval"])
def test_mi_and_stats_match_reference(golden, name):
    g = golden(name)
    c = case_inputs(g)
    p = case_params(g)
    mu, lv = O.encoder_forward(p, c["x"])
    assert_close(mu, g["mu"], 1e-5, "mu", floor=1e-3)
    assert_close(lv, g["logvar"], 1e-5, "logvar", floor=1e-3)
    mi = O.calc_mi_from_stats(mu, lv, torch.from_numpy(g["eps_mi"]))
    assert abs(mi - float(g["mi"])) <= 1e-4 * max(1.0, abs(float(g["mi"])))


def test_multisample_forward_matches_reference(golden):
    g = golden("aligned_ns3_eval")
    c = case_inputs(g)
    loss, rec, kl = O.vae_loss(case_params(g), c["x"], c["klw"], c["eps"])
    assert_close(loss, g["loss"], 1e-5, "loss")
    assert_close(rec, g["rec"], 1e-5, "rec")
    assert_close(kl, g["kl"], 1e-5, "kl")


def test_stock_init_kat_values(golden):
    """Known-answer values quoted in SURVEY §8(c3) / BASELINE.md for the toy shape at the stock init."""
    g = golden("toy_stockinit_eval")
    assert abs(float(g["loss"].sum()) - 2432.931396) < 2e-3
    assert abs(float(g["mi"]) - (-0.0321596)) < 1e-5
    assert abs(float(g["grad_norm"]) - 0.0469287) < 1e-6


def test_fast_port_matches_explicit_oracle():
    """bench.py times FastPort (nn.LSTM-based); pin it to the explicit restatement."""
    torch.manual_seed(0)
    V, ni, nh, nz, B, T = 97, 12, 20, 3, 4, 6
    m = O.FastPort(V, ni, nh, nz).eval()
    p = {
        "encoder.embed.weight": m.e_embed.weight, "encoder.lstm.weight_ih_l0": m.e_lstm.weight_ih_l0,
        "encoder.lstm.weight_hh_l0": m.e_lstm.weight_hh_l0, "encoder.lstm.bias_ih_l0": m.e_lstm.bias_ih_l0,
        "encoder.lstm.bias_hh_l0": m.e_lstm.bias_hh_l0, "encoder.linear.weight": m.e_lin.weight,
        "decoder.embed.weight": m.d_embed.weight, "decoder.trans_linear.weight": m.d_trans.weight,
        "decoder.lstm.weight_ih_l0": m.d_lstm.weight_ih_l0, "decoder.lstm.weight_hh_l0": m.d_lstm.weight_hh_l0,
        "decoder.lstm.bias_ih_l0": m.d_lstm.bias_ih_l0, "decoder.lstm.bias_hh_l0": m.d_lstm.bias_hh_l0,
        "decoder.pred_linear.weight": m.d_pred.weight,
    }
    p = O.scale_trained_like({k: v.detach().clone() for k, v in p.items()})
    with torch.no_grad():
        for k, q in zip(O.ALL_KEYS, [m.e_embed.weight, m.e_lstm.weight_ih_l0, m.e_lstm.weight_hh_l0, m.e_lstm.bias_ih_l0,
                                     m.e_lstm.bias_hh_l0, m.e_lin.weight, m.d_embed.weight, m.d_trans.weight,
                                     m.d_lstm.weight_ih_l0, m.d_lstm.weight_hh_l0, m.d_lstm.bias_ih_l0,
                                     m.d_lstm.bias_hh_l0, m.d_pred.weight]):
            q.copy_(p[k])
    x = O.make_token_batch(B, T, V)
    torch.manual_seed(5)
    loss, rec, kl = m.loss(x, 0.7)
    torch.manual_seed(5)
    eps = torch.zeros(B, nz).normal_().unsqueeze(1)
    l2, r2, k2 = O.vae_loss(p, x, 0.7, eps)
    assert_close(loss, l2, 1e-5, "fastport loss")
    assert_close(kl, k2, 1e-5, "fastport kl")


def test_burn_window_rule():
    """text.py:366-400: break when the 15-step windowed mean loss per word rises."""
    w = O.BurnWindow(window=15)
    steps = 0
    while w.keep_going():
        steps += 1
        # first window mean 10.0, second window mean 11.0 -> break at sub_iter 30
        val = 10.0 if steps <= 15 else 11.0
        if w.update(val * 100, 100):
            break
    assert steps == 30
    w = O.BurnWindow(window=15)
    n = 0
    while w.keep_going():
        n += 1
        if w.update(1000.0 - n, 100):
            break
    assert n == 99  # hard cap: sub_iter < 100 -> at most 99 updates (text.py:371)


def test_generation_oracle_matches_reference_tokens(golden):
    """Greedy and beam decoding (dec_lstm.py:163-314) of the oracle against the token ids the unmodified
    reference produced (fixture written by oracle/validate_against_reference.py::run_generation_case)."""
    import lagging_oracle as O
    g = golden("generation_small")
    p = {k: torch.from_numpy(g["p." + k]) for k in O.ALL_KEYS}
    z = torch.from_numpy(g["z"])
    want_g = [[int(t) for t in row if t >= 0] for row in g["greedy"]]
    want_b = [[int(t) for t in row if t >= 0] for row in g["beam"]]
    assert O.greedy_decode(p, z) == want_g
    assert O.beam_search_decode(p, z, 5) == want_b
    assert any(len(s) < 99 for s in want_g) and want_g != [s[1:] for s in want_b]      # the case is not degenerate


def test_oracle_full_shape_forward_vs_reference_fingerprints(golden):
    """The oracle at a full BASELINE shape (configs[2], Yelp: V=19997, T=100, B=32, ni=512, nh=1024, nz=32) against the
    outputs of the unmodified reference (forward only: loss, rec, KL, mu)."""
    import lagging_oracle as O
    g = golden("yelp_eval")
    V, ni, nh, nz, B, T, ns, _ = [int(v) for v in g["meta"]]
    p = O.scale_trained_like(O.init_text_params(V, ni, nh, nz, seed=0), 4.0)
    x = O.make_token_batch(B, T, V)
    with torch.no_grad():
        loss, rec, kl = O.vae_loss(p, x, float(g["kl_weight"]), torch.from_numpy(g["eps"]))
        mu, _ = O.encoder_forward(p, x)
    assert_close(loss, g["loss"], 2e-5, "loss")
    assert_close(rec, g["rec"], 2e-5, "rec")
    assert_close(kl, g["kl"], 2e-5, "kl", floor=1e-2)
    assert_close(mu, g["mu"], 2e-5, "mu", floor=1e-2)


def test_nll_iw_oracle_matches_reference(golden):
    """Importance-weighted NLL (vae.py:100-129) of the oracle against the unmodified reference's value."""
    import lagging_oracle as O
    g2 = golden("aligned_nll_iw")
    g = golden(str(g2["base"]))
    p = case_params(g)
    x = torch.from_numpy(g["x"])
    with torch.no_grad():
        nll = O.nll_iw(p, x, [torch.from_numpy(e) for e in g2["eps"]])
    assert_close(nll, g2["nll"], 1e-6, "nll_iw")


def test_image_oracle_eval_mode_vs_reference(golden):
    """eval() forward of the image model (BatchNorm running statistics) of the oracle against the unmodified reference."""
    import image_oracle as IO
    g = golden("omniglot_eval_b5")
    B, nz, ns = [int(v) for v in g["meta"]]
    p = IO.init_image_params(nz, seed=0)
    for k in g:
        if k.startswith("stat."):
            p[k[5:]] = torch.from_numpy(g[k])
    with torch.no_grad():
        loss, rec, kl = IO.vae_loss(p, torch.from_numpy(g["x"]), float(g["kl_weight"]), torch.from_numpy(g["eps"]), training=False)
    assert_close(loss, g["loss"], 2e-5, "loss")
    assert_close(rec, g["rec"], 2e-5, "rec")
    assert_close(kl, g["kl"], 2e-5, "kl", floor=1e-2)
