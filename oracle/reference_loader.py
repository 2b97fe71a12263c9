"""Locate and import the UNMODIFIED reference `modules` package under a private name (test / bench infrastructure only:
it never sits on the product path).  Search order: $VAE_REF_PATH, <repo>/baseline/_ref (scripts/stage_reference.sh;
git-ignored, travels to the GPU box), /root/reference (authoring container)."""
import importlib.util
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_dir():
    for c in (os.environ.get("VAE_REF_PATH"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if c and os.path.isfile(os.path.join(c, "modules", "__init__.py")):
            return c
    return None


def load_reference_modules(ref=None):
    ref = ref or reference_dir()
    if ref is None:
        return None
    if "ref_modules" in sys.modules:
        return sys.modules["ref_modules"]
    spec = importlib.util.spec_from_file_location("ref_modules", os.path.join(ref, "modules", "__init__.py"),
                                                  submodule_search_locations=[os.path.join(ref, "modules")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_modules"] = mod
    spec.loader.exec_module(mod)
    return mod


class Vocab(dict):
    """Minimal stand-in for data.text_data.VocabEntry (len, ['<s>'], id2word)."""

    def __init__(self, V):
        super().__init__()
        self.V = V
        self["<pad>"], self["<s>"], self["</s>"], self["<unk>"] = 0, 1, 2, 3

    def __len__(self):
        return self.V

    def id2word(self, i):
        return str(i)


def build_reference_vae(ref, V, ni, nh, nz, device, p_in=0.5, p_out=0.5, seed=0):
    """The reference's own constructors with the initialisers of text.py:265-266."""
    import warnings
    torch.manual_seed(seed)
    args = types.SimpleNamespace(ni=ni, enc_nh=nh, dec_nh=nh, nz=nz, dec_dropout_in=p_in, dec_dropout_out=p_out,
                                 device=torch.device(device))
    mi = lambda t: torch.nn.init.uniform_(t, -0.01, 0.01)
    ei = lambda t: torch.nn.init.uniform_(t, -0.1, 0.1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        vae = ref.VAE(ref.LSTMEncoder(args, V, mi, ei), ref.LSTMDecoder(args, Vocab(V), mi, ei), args).to(args.device)
    return vae


class ReferenceStepper:
    """One iteration of text.py:373-387 on the reference's modules (stock torch optimisers)."""

    def __init__(self, vae, lr=1.0):
        self.vae = vae.train()
        self.enc_opt = torch.optim.SGD(vae.encoder.parameters(), lr=lr, momentum=0)
        self.dec_opt = torch.optim.SGD(vae.decoder.parameters(), lr=lr, momentum=0)

    def inner_step(self, x, kl_weight):
        self.enc_opt.zero_grad()
        self.dec_opt.zero_grad()
        loss, _, _ = self.vae.loss(x, kl_weight, nsamples=1)          # text.py:379
        s = loss.sum().item()                                         # text.py:381
        loss.mean(dim=-1).backward()                                  # text.py:382-384
        torch.nn.utils.clip_grad_norm_(self.vae.parameters(), 5.0)    # text.py:385
        self.enc_opt.step()                                           # text.py:387
        return s
