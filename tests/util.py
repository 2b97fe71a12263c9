"""Shared helpers for the parity tests (oracle side = checker only)."""
import numpy as np
import torch

import lagging_oracle as O

FULL_CASES = ["toy_eval", "toy_train", "ragged_train", "aligned_train", "toy_stockinit_eval"]


def case_params(g):
    return {k: torch.from_numpy(g["p." + k]).clone() for k in O.ALL_KEYS}


def case_inputs(g):
    V, ni, nh, nz, B, T, ns, train = [int(v) for v in g["meta"]]
    x = torch.from_numpy(g["x"])
    eps = torch.from_numpy(g["eps"])
    mi = mo = None
    if train:
        mi = torch.from_numpy(g["mask_in"]).float() * 2.0
        mo = torch.from_numpy(g["mask_out"]).float() * 2.0
    return dict(V=V, ni=ni, nh=nh, nz=nz, B=B, T=T, ns=ns, train=bool(train), x=x, eps=eps,
                mask_in=mi, mask_out=mo, klw=float(g["kl_weight"]))


def rel_err(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def assert_close(a, b, tol, what, floor=0.0):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    err = float((a - b).abs().max())
    scale = max(float(b.abs().max()), floor)
    assert err <= tol * scale, "%s: abs err %.3e > %.1e * %.3e" % (what, err, tol, scale)
