"""Pins the oracle: runs the UNMODIFIED reference modules (imported from /root/reference, CPU)
and the restatement in oracle/lagging_oracle.py on the same seeded inputs, asserts agreement,
and writes the golden fixtures under tests/golden/.  Run here (authoring container) only:

    python oracle/validate_against_reference.py [--yahoo]

/root/reference does not exist on the GPU box; tests read only the committed .npz files.
"""
import argparse
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import lagging_oracle as O  # noqa: E402

REF = os.environ.get("VAE_REF_PATH", "/root/reference")


def load_reference_modules():
    """Import the reference `modules` package under a private name so it can coexist with the
    new drop-in `modules` package."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "ref_modules", os.path.join(REF, "modules", "__init__.py"),
        submodule_search_locations=[os.path.join(REF, "modules")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_modules"] = mod
    spec.loader.exec_module(mod)
    return mod


class _Vocab(dict):
    """Minimal stand-in for data.text_data.VocabEntry (len, ['<s>'], id2word)."""

    def __init__(self, V):
        super().__init__()
        self.V = V
        self["<pad>"], self["<s>"], self["</s>"], self["<unk>"] = 0, 1, 2, 3

    def __len__(self):
        return self.V

    def id2word(self, i):
        return str(i)


def build_reference(ref, V, ni, nh, nz, p_in, p_out, seed=0):
    torch.manual_seed(seed)
    args = types.SimpleNamespace(ni=ni, enc_nh=nh, dec_nh=nh, nz=nz, dec_dropout_in=p_in,
                                 dec_dropout_out=p_out, device=torch.device("cpu"))
    mi = lambda t: torch.nn.init.uniform_(t, -0.01, 0.01)
    ei = lambda t: torch.nn.init.uniform_(t, -0.1, 0.1)
    enc = ref.LSTMEncoder(args, V, mi, ei)
    dec = ref.LSTMDecoder(args, _Vocab(V), mi, ei)
    return ref.VAE(enc, dec, args)


def scale_head(vae, s):
    """s == 1.0 keeps the stock init; otherwise O.scale_trained_like (see its docstring)."""
    if s == 1.0:
        return
    O.scale_trained_like({k: q for k, q in vae.named_parameters()}, 4.0)


def params_of(vae):
    sd = vae.state_dict()
    return {k: sd[k].clone() for k in O.ALL_KEYS}


def run_case(ref, name, V, ni, nh, nz, B, T, ns, train, klw, head_scale, out_dir, full=True):
    vae = build_reference(ref, V, ni, nh, nz, 0.5, 0.5)
    scale_head(vae, head_scale)
    p0 = params_of(vae)
    x = O.make_token_batch(B, T, V)
    vae.train() if train else vae.eval()

    # --- reference forward/backward with its own RNG consumption order (eps, drop_in, drop_out)
    torch.manual_seed(1)
    loss, rec, kl = vae.loss(x, klw, nsamples=ns)
    vae.zero_grad()
    loss.mean(dim=-1).backward()
    ref_grads = {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p))
                 for k, p in zip(O.ALL_KEYS, vae.parameters())}
    ref_norm = float(torch.nn.utils.clip_grad_norm_(vae.parameters(), 5.0))

    # --- regenerate the same random draws explicitly (CPU generator, same call shapes/order)
    torch.manual_seed(1)
    eps = torch.zeros(B, ns, nz).normal_()                                # encoder.py:77
    mask_in = mask_out = None
    if train:
        mask_in = torch.nn.functional.dropout(torch.ones(B, T - 1, ni), 0.5, True)       # dec_lstm.py:81
        # nn.LSTM(batch_first) returns a [T,B,nh]-strided view on CPU; bernoulli_ follows memory order
        mask_out = torch.nn.functional.dropout(torch.ones(T - 1, B * ns, nh).transpose(0, 1), 0.5, True)  # dec_lstm.py:106

    # --- oracle
    p = {k: v.clone() for k, v in p0.items()}
    if ns == 1:
        r = O.inner_step(p, x, klw, eps, mask_in, mask_out, update=True)
        o_loss, o_rec, o_kl = r["loss"], r["rec"], r["kl"]
    else:
        o_loss, o_rec, o_kl = O.vae_loss(p, x, klw, eps, mask_in, mask_out)
        r = None

    def close(a, b, tol, what):
        err = float((a - b).abs().max() / (b.abs().max() + 1e-30))
        assert err < tol, f"{name}: {what} mismatch rel {err:.3e}"
        return err

    e1 = close(o_loss, loss.detach(), 2e-5, "loss")
    close(o_rec, rec.detach(), 2e-5, "rec")
    e2 = float((o_kl - kl.detach()).abs().max())
    assert e2 < 1e-5 * max(1.0, float(kl.abs().max())), f"{name}: KL abs err {e2}"
    gerr = 0.0
    if r is not None:
        for k in O.ALL_KEYS:
            gerr = max(gerr, close(r["grads"][k], ref_grads[k], 5e-4, "grad " + k))
        assert abs(r["grad_norm"] - ref_norm) < 2e-5 * ref_norm, (r["grad_norm"], ref_norm)
        # reference post-step encoder params: SGD lr 1 on clipped grads (text.py:385-387)
        opt = torch.optim.SGD(vae.encoder.parameters(), lr=1.0, momentum=0)
        opt.step()
        for k, q in zip(O.ENC_KEYS, vae.encoder.parameters()):
            close(p[k], q.detach(), 1e-6, "post-step " + k)

    # --- MI (eval-mode forward; one draw, encoder.py:128)
    vae.eval()
    vae.load_state_dict({**vae.state_dict(), **p0})
    torch.manual_seed(2)
    mi_ref = vae.calc_mi_q(x)
    torch.manual_seed(2)
    eps_mi = torch.zeros(B, 1, nz).normal_()
    mi_o = O.calc_mi(p0, x, eps_mi)
    assert abs(mi_o - mi_ref) < 1e-4 * max(1.0, abs(mi_ref)), (mi_o, mi_ref)
    mu_ref, lv_ref = vae.encode_stats(x)

    print(f"[{name}] loss.sum={float(loss.sum()):.6f} rec.sum={float(rec.sum()):.6f} "
          f"KL.sum={float(kl.sum()):.6e} MI={mi_ref:.7f} gnorm={ref_norm:.7f} "
          f"(oracle rel err loss {e1:.1e}, KL abs {e2:.1e}, grads {gerr:.1e})")

    out = {
        "meta": np.array([V, ni, nh, nz, B, T, ns, int(train)], dtype=np.int64),
        "kl_weight": np.float64(klw),
        "x": x.numpy(), "eps": eps.numpy(), "eps_mi": eps_mi.numpy(),
        "loss": loss.detach().numpy(), "rec": rec.detach().numpy(), "kl": kl.detach().numpy(),
        "mu": mu_ref.detach().numpy(), "logvar": lv_ref.detach().numpy(),
        "mi": np.float64(mi_ref), "grad_norm": np.float64(ref_norm),
    }
    if train:
        out["mask_in"] = (mask_in != 0).numpy()
        out["mask_out"] = (mask_out != 0).numpy()
    if full:
        for k in O.ALL_KEYS:
            out["p." + k] = p0[k].numpy()
            if r is not None:
                out["g." + k] = ref_grads[k].numpy()
        if r is not None:
            for k, q in zip(O.ENC_KEYS, vae.encoder.parameters()):
                pass
            for k in O.ENC_KEYS:
                out["post." + k] = p[k].numpy()
    else:
        # big shape: parameters are regenerated by seed in the tests (build_reference is the
        # reference ctor -> not available there), so store the parameters' fingerprints plus
        # gradient fingerprints instead of 215 MB of tensors.
        for k in O.ALL_KEYS:
            g = ref_grads[k]
            out["gnorm." + k] = np.float64(g.double().norm())
            out["gslice." + k] = g.reshape(-1)[:: max(1, g.numel() // 64)][:64].numpy()
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **out)
    return p0


def run_nll_iw_case(ref, name, base, V, ni, nh, nz, B, T, nsamples, ns, out_dir):
    """Evaluation path: the reference's VAE.nll_iw in eval() mode against the oracle restatement, on the
    parameters / tokens of the fixture `base` (run_case with the same arguments rebuilds them); stores draws + result."""
    vae = build_reference(ref, V, ni, nh, nz, 0.5, 0.5)
    scale_head(vae, 1.5)
    p0 = params_of(vae)
    x = O.make_token_batch(B, T, V)
    vae.eval()
    torch.manual_seed(7)
    with torch.no_grad():
        want = vae.nll_iw(x, nsamples, ns=ns)
    torch.manual_seed(7)
    chunks = [torch.zeros(B, ns, nz).normal_() for _ in range(nsamples // ns)]       # encoder.py:77, one draw per chunk
    with torch.no_grad():
        got = O.nll_iw(p0, x, chunks)
    err = float((got - want).abs().max() / want.abs().max())
    assert err < 2e-6, (name, err)
    g = dict(np.load(os.path.join(out_dir, base + ".npz")))
    assert np.array_equal(g["x"], x.numpy()) and all(np.array_equal(g["p." + k], p0[k].numpy()) for k in O.ALL_KEYS), \
        "nll_iw case must share parameters and tokens with " + base
    print("[%s] nll_iw mean %.6f (oracle rel err %.1e)" % (name, float(want.mean()), err))
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), base=np.array(base), nll=want.numpy(),
                        eps=np.stack([c.numpy() for c in chunks]), ns=np.int64(ns))


def run_generation_case(ref, name, V, ni, nh, nz, n, out_dir):
    """Generation paths: the reference's greedy and beam decoders on a trained-like model against the oracle
    restatement; writes the parameters, latents and token ids as a fixture."""
    vae = build_reference(ref, V, ni, nh, nz, 0.5, 0.5, seed=5)
    O.scale_trained_like({k: q for k, q in vae.named_parameters()}, 6.0)
    with torch.no_grad():                                  # decisive, non-degenerate next-token distributions
        vae.decoder.pred_linear.weight.mul_(6.0)
        vae.decoder.lstm.weight_ih_l0.mul_(2.0)
    vae.eval()
    p0 = params_of(vae)
    z = torch.randn(n, nz, generator=torch.Generator().manual_seed(11)) * 1.5
    with torch.no_grad():
        g_ref = [[int(w) for w in s] for s in vae.decode(z, "greedy")]
        b_ref = [[int(w) for w in s] for s in vae.decode(z, "beam", K=5)]
    g_o = O.greedy_decode(p0, z)
    b_o = O.beam_search_decode(p0, z, 5)
    assert g_o == g_ref, (name, "greedy", g_o[:2], g_ref[:2])
    assert b_o == b_ref, (name, "beam", b_o[:2], b_ref[:2])
    lens = [len(s) for s in g_ref]
    print("[%s] greedy lengths %s, beam lengths %s: oracle == reference" % (name, lens, [len(s) for s in b_ref]))
    assert len(set(tuple(s) for s in g_ref)) > 1 and min(lens) < 99, "degenerate generation case"
    out = {"meta": np.array([V, ni, nh, nz, n], dtype=np.int64), "z": z.numpy(),
           "greedy": np.array([s + [-1] * (100 - len(s)) for s in g_ref], dtype=np.int64),
           "beam": np.array([s + [-1] * (102 - len(s)) for s in b_ref], dtype=np.int64)}
    for k in O.ALL_KEYS:
        out["p." + k] = p0[k].numpy()
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--yahoo", action="store_true", help="also run the Yahoo-shape KAT (slow)")
    ap.add_argument("--yahoo-train", dest="yahoo_train", action="store_true", help="only the Yahoo-shape train()-mode fixture")
    ap.add_argument("--only-big", action="store_true", help="skip the small cases")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    ref = load_reference_modules()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    if not a.only_big:
        # config #1 of BASELINE.json (toy.py: nz=1, toy.py:102), eval + train-mode masks
        run_case(ref, "toy_eval", 1004, 50, 50, 1, 32, 12, 1, False, 1.0, 1.5, out_dir)
        run_case(ref, "toy_train", 1004, 50, 50, 1, 32, 12, 1, True, 0.5, 1.5, out_dir)
        # ragged batch (< batch_size rows, text_data.py:241-247) + odd sizes
        run_case(ref, "ragged_train", 301, 24, 40, 3, 5, 7, 1, True, 0.1, 1.5, out_dir)
        # tensor-core-aligned small shape
        run_case(ref, "aligned_train", 520, 64, 128, 8, 16, 10, 1, True, 0.1, 1.5, out_dir)
        # multi-sample forward (nsamples>1: dec_lstm.py:86-94,139-140)
        run_case(ref, "aligned_ns3_eval", 520, 64, 128, 8, 8, 9, 3, False, 1.0, 1.5, out_dir)
        # KL ~ 0 regime at the stock init (no head scaling) — conditioning check
        run_case(ref, "toy_stockinit_eval", 1004, 50, 50, 1, 32, 12, 1, False, 1.0, 1.0, out_dir)
        run_generation_case(ref, "generation_small", 24, 16, 32, 4, 8, out_dir)
        run_nll_iw_case(ref, "aligned_nll_iw", "aligned_ns3_eval", 520, 64, 128, 8, 8, 9, 6, 3, out_dir)
    if a.yahoo:
        # BASELINE.json configs[1] and configs[2] at their full shapes; the parameters are regenerated in the tests from
        # O.init_text_params(seed) (the reference is built FROM the oracle parameters here), fixtures keep fingerprints
        run_big_case(ref, "yahoo_eval", 20001, 512, 1024, 32, 32, 200, 0.1, out_dir)
        run_big_case(ref, "yelp_eval", 19997, 512, 1024, 32, 32, 100, 1.0, out_dir)
    if a.yahoo or a.yahoo_train:
        # the configuration bench.py times: Yahoo shape, train() mode (dropout 0.5/0.5), kl_weight 0.1
        run_big_case(ref, "yahoo_train", 20001, 512, 1024, 32, 32, 200, 0.1, out_dir, train=True)


def run_big_case(ref, name, V, ni, nh, nz, B, T, klw, out_dir, train=False):
    """Full-shape fingerprint fixture.  train=True: train() mode with the reference's own dropout draws (replayed and
    stored bit-packed: np.packbits over the flattened keep-masks, ~1.2 MB at the Yahoo shape) — the configuration
    bench.py times; also stores the post-step encoder parameters (clip 5.0 + SGD lr 1, text.py:385-387) as fingerprints."""
    vae = build_reference(ref, V, ni, nh, nz, 0.5, 0.5)
    p0 = O.init_text_params(V, ni, nh, nz, seed=0)
    O.scale_trained_like(p0, 4.0)
    sd = vae.state_dict()
    sd.update(p0)
    vae.load_state_dict(sd)
    x = O.make_token_batch(B, T, V)
    vae.train() if train else vae.eval()
    torch.manual_seed(1)
    loss, rec, kl = vae.loss(x, klw)
    vae.zero_grad()
    loss.mean(dim=-1).backward()
    grads = {k: (q.grad.clone() if q.grad is not None else None) for k, q in zip(O.ALL_KEYS, vae.parameters())}  # clone: clip scales .grad in place
    gn = float(torch.nn.utils.clip_grad_norm_(vae.parameters(), 5.0))
    torch.manual_seed(1)
    eps = torch.zeros(B, 1, nz).normal_()
    out = {}
    if train:
        mask_in = torch.nn.functional.dropout(torch.ones(B, T - 1, ni), 0.5, True) != 0                      # dec_lstm.py:81
        mask_out = torch.nn.functional.dropout(torch.ones(T - 1, B, nh).transpose(0, 1), 0.5, True) != 0    # dec_lstm.py:106
        # the replayed draws must be the ones the reference consumed: check through the oracle's loss
        o_loss, _, _ = O.vae_loss({k: v.clone() for k, v in p0.items()}, x, klw, eps, mask_in.float() * 2, mask_out.float() * 2)
        err = float((o_loss - loss.detach()).abs().max() / loss.detach().abs().max())
        assert err < 2e-5, (name, "replayed dropout masks do not reproduce the reference loss", err)
        out["mask_in_bits"] = np.packbits(mask_in.contiguous().numpy().reshape(-1))
        out["mask_out_bits"] = np.packbits(mask_out.contiguous().numpy().reshape(-1))
        torch.optim.SGD(vae.encoder.parameters(), lr=1.0, momentum=0).step()                 # text.py:387
        for k, q in zip(O.ENC_KEYS, vae.encoder.parameters()):
            q = q.detach()
            out["postnorm." + k] = np.float64(q.double().norm())
            out["postslice." + k] = q.reshape(-1)[:: max(1, q.numel() // 64)][:64].clone().numpy()   # clone: the values are restored below
            out["dnorm." + k] = np.float64((q - p0[k]).double().norm())     # size of the update itself
        vae.load_state_dict({**vae.state_dict(), **p0})
    vae.eval()
    torch.manual_seed(2)
    mi = vae.calc_mi_q(x)
    torch.manual_seed(2)
    eps_mi = torch.zeros(B, 1, nz).normal_()
    out.update({"meta": np.array([V, ni, nh, nz, B, T, 1, int(train)], dtype=np.int64), "kl_weight": np.float64(klw),
                "eps": eps.numpy(), "eps_mi": eps_mi.numpy(), "loss": loss.detach().numpy(),
                "rec": rec.detach().numpy(), "kl": kl.detach().numpy(), "mi": np.float64(mi),
                "grad_norm": np.float64(gn)})
    mu, lv = vae.encode_stats(x)
    out["mu"], out["logvar"] = mu.detach().numpy(), lv.detach().numpy()
    for k in O.ALL_KEYS:
        g = grads[k] if grads[k] is not None else torch.zeros_like(p0[k])
        out["gnorm." + k] = np.float64(g.double().norm())
        out["gslice." + k] = g.reshape(-1)[:: max(1, g.numel() // 64)][:64].numpy()
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **out)
    print(f"[{name}] loss.sum={float(loss.sum()):.6f} KL.sum={float(kl.sum()):.6e} MI={mi:.7f} gnorm={gn:.7f}")


if __name__ == "__main__":
    main()
