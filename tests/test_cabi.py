"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads and exports every symbol
include/lagvae.h declares; the product path refuses to run without a CUDA device (no CPU fallback)."""
import os
import re
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "lagvae.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lagvae_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import lagvae._backend as be
    L = be.lib()
    syms = _declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(L, s), "missing export " + s
        assert s in be.PROTOTYPES, "binding lacks prototype for " + s
    assert sorted(be.PROTOTYPES) == syms
    assert L.lagvae_abi_version() == be.ABI_VERSION


def test_workspace_query_and_param_count_are_host_only():
    import ctypes as C
    import lagvae._backend as be
    d = be.TextDims(32, 200, 1, 20001, 512, 1024, 32)
    n = be.lib().lagvae_text_param_count(C.byref(d))
    assert n == 16605696 + 37185024          # SURVEY §2.1 parameter counts
    ws = be.lib().lagvae_text_workspace_bytes(C.byref(d), 0)
    assert 1 << 28 < ws < 1 << 34
    bad = be.TextDims(0, 200, 1, 20001, 512, 1024, 32)
    assert be.lib().lagvae_text_workspace_bytes(C.byref(bad), 0) == 0


def test_no_cpu_fallback():
    import lagvae
    import modules
    with pytest.raises(lagvae.LagvaeError):
        lagvae.TextEngine(100, 8, 16, 2, "cpu")
    a = types.SimpleNamespace(ni=8, enc_nh=16, dec_nh=16, nz=2, dec_dropout_in=0.5, dec_dropout_out=0.5, device="cpu")

    class Vocab(dict):
        def __len__(self):
            return 50
    init = lambda t: torch.nn.init.uniform_(t, -0.01, 0.01)
    vae = modules.VAE(modules.LSTMEncoder(a, 50, init, init), modules.LSTMDecoder(a, Vocab(), init, init), a)
    with pytest.raises(lagvae.LagvaeError):
        vae.loss(torch.zeros(2, 5, dtype=torch.long), 1.0)
    with pytest.raises(lagvae.LagvaeError):
        vae.calc_mi_q(torch.zeros(2, 5, dtype=torch.long))


def test_state_dict_keys_match_reference():
    import modules
    a = types.SimpleNamespace(ni=8, enc_nh=16, dec_nh=16, nz=2, dec_dropout_in=0.5, dec_dropout_out=0.5, device="cpu")

    class Vocab(dict):
        def __len__(self):
            return 50
    init = lambda t: torch.nn.init.uniform_(t, -0.01, 0.01)
    vae = modules.VAE(modules.LSTMEncoder(a, 50, init, init), modules.LSTMDecoder(a, Vocab(), init, init), a)
    want = ["encoder.embed.weight", "encoder.lstm.weight_ih_l0", "encoder.lstm.weight_hh_l0",
            "encoder.lstm.bias_ih_l0", "encoder.lstm.bias_hh_l0", "encoder.linear.weight",
            "decoder.embed.weight", "decoder.trans_linear.weight", "decoder.lstm.weight_ih_l0",
            "decoder.lstm.weight_hh_l0", "decoder.lstm.bias_ih_l0", "decoder.lstm.bias_hh_l0",
            "decoder.pred_linear.weight", "decoder.loss.weight"]
    assert list(vae.state_dict().keys()) == want
    assert [n for n, _ in vae.named_parameters()] == want[:-1]
    assert len(vae.encoder_params()) == 6 and len(vae.decoder_params()) == 7
    assert vae.decoder.embed.padding_idx == 49


def test_philox_host_device_contract():
    """Philox4x32-10 known-answer (Random123 kat: counter=0,key=0)."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c, k = [0, 0, 0, 0], [0, 0]
    for _ in range(10):
        hi0, lo0 = (M0 * c[0]) >> 32, (M0 * c[0]) & 0xFFFFFFFF
        hi1, lo1 = (M1 * c[2]) >> 32, (M1 * c[2]) & 0xFFFFFFFF
        c = [hi1 ^ c[1] ^ k[0], lo1, hi0 ^ c[3] ^ k[1], lo0]
        k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
    assert c == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]


def test_image_geometry_queries_are_host_only():
    """convtc / pixelblock size queries of include/lagvae.h answer without a device; unsupported geometries are refused."""
    import ctypes as C
    import lagvae._backend as be
    L = be.lib()
    assert L.lagvae_convtc_supported(64, 28, 28, 32, 32, 7, 7) == 1 and L.lagvae_convtc_supported(64, 28, 28, 64, 32, 1, 1) == 1
    assert L.lagvae_convtc_supported(64, 28, 28, 5, 64, 7, 7) == 0 and L.lagvae_convtc_supported(64, 28, 28, 32, 32, 4, 4) == 0
    # forward tiles: 25 live taps reserved as 49 x [64 rows x 64] bf16, dgrad the same
    assert L.lagvae_convtc_wbuf_bytes(32, 32, 7, 7) == 2 * 49 * 64 * 64 * 2 + 256
    d = be.PixelBlockDims(64, 28, 28, 64, 32, 7, 1e-5, 0.1, 0)
    R = 64 * 28 * 28
    stash = L.lagvae_pixelblock_stash_bytes(C.byref(d))
    # y1, y2 (fp32 x 32), y3 (fp32 x 64), a1cat, a2cat (bf16 x 64), xcat (bf16 x 128) + weight tiles + statistics
    assert stash >= R * (2 * 32 * 4 + 64 * 4 + 2 * 64 * 2 + 128 * 2) and stash < 2 * R * (2 * 32 * 4 + 64 * 4 + 2 * 64 * 2 + 128 * 2)
    assert L.lagvae_pixelblock_scratch_bytes(C.byref(d)) > R * 64 * 4
    bad = be.PixelBlockDims(64, 28, 28, 48, 32, 7, 1e-5, 0.1, 0)
    assert L.lagvae_pixelblock_stash_bytes(C.byref(bad)) == 0


def test_image_modules_refuse_cpu_tensors():
    import lagvae
    import modules
    a = types.SimpleNamespace(nz=8, latent_feature_map=4, device="cpu")
    vae = modules.VAE(modules.ResNetEncoderV2(a), modules.PixelCNNDecoderV2(a), a)
    with pytest.raises(lagvae.LagvaeError):
        vae.loss(torch.zeros(2, 1, 28, 28), 1.0)
    with pytest.raises(lagvae.LagvaeError):
        vae.decoder.decode(torch.zeros(2, 8), True)


def test_bench_work_model_matches_survey_figures():
    """The algorithmic-work figures bench.py / scripts/bench_sweep.py report against the numbers stated in SURVEY §8 d4
    (Yahoo config: F_step = 1.2695 TFLOP, linear in B; compulsory HBM bytes ~1.4-1.5 GB at B=32)."""
    import importlib.util
    root = ROOT
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    c = bench.CFG
    F = bench.flops_step(c["B"], c["T"], c["V"], c["ni"], c["nh"], c["nz"])
    assert abs(F - 1.2695e12) < 1e9
    assert bench.flops_step(512, c["T"], c["V"], c["ni"], c["nh"], c["nz"]) == 16 * F
    enc_fwd = 2 * 32 * (200 * 512 * 4096 + 200 * 1024 * 4096 + 1024 * 64)
    assert abs(enc_fwd - 80.53e9) < 1e8                                  # encoder forward 80.53 GFLOP
    assert bench.METRIC.startswith("aggressive inner-loop encoder steps/sec")


def test_dropout_struct_matches_the_header_and_graph_seed_word_arithmetic():
    """lagvae_dropout gained `seed_dev` (ABI 2): the ctypes mirror must have the C layout (4 + 2*4 + pad, two pointers, u64,
    pointer = 48 bytes on LP64) and the header must declare the field; the graph-side seed word advances by PHILOX_STEP mod 2^64
    exactly as the eager seed does (lagvae/graph.py, modules/text.py::dropout_spec)."""
    import ctypes as C
    import torch
    from lagvae import _backend as be
    from lagvae import graph as G
    names = [f[0] for f in be.Dropout._fields_]
    assert names == ["mode", "p_in", "p_out", "mask_in", "mask_out", "seed", "seed_dev"]
    assert C.sizeof(be.Dropout) == 48 and be.Dropout.seed_dev.offset == 40 and be.Dropout.seed.offset == 32
    hdr = open(os.path.join(ROOT, "include", "lagvae.h")).read()
    assert "const uint64_t* seed_dev;" in hdr and "#define LAGVAE_ABI_VERSION 2" in hdr
    word = torch.zeros(1, dtype=torch.int64)
    for k in range(1, 6):
        G.bump_philox_word(word)
        assert int(word) & (2 ** 64 - 1) == (k * G.PHILOX_STEP) & (2 ** 64 - 1)
    G.bump_philox_word(word, times=3)
    assert int(word) & (2 ** 64 - 1) == (8 * G.PHILOX_STEP) & (2 ** 64 - 1)
    assert G.capture_serial() is None                     # no GraphedStep capture in progress


def test_dp_overlap_default_is_one_bucket_after_the_backward(monkeypatch):
    from lagvae import dp
    monkeypatch.delenv("LAGVAE_DP_OVERLAP", raising=False)
    assert dp._overlap_default() is False
    monkeypatch.setenv("LAGVAE_DP_OVERLAP", "1")
    assert dp._overlap_default() is True


def test_graft_entry_abi_check_agrees_with_header_binding_and_library():
    """__graft_entry__.build() ends with this check (the driver runs build() on the CPU box every round)."""
    import sys
    sys.path.insert(0, ROOT)
    import __graft_entry__ as G
    G.check_abi()


def test_public_method_surface_covers_the_reference_classes():
    """Every public method of the reference's VAE / LSTMEncoder / LSTMDecoder / ResNetEncoderV2 / PixelCNNDecoderV2 exists
    on the drop-in class of the same name (b1: 'same names'); needs the staged reference (scripts/stage_reference.sh)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import reference_loader as RL
    ref = RL.load_reference_modules()
    if ref is None:
        pytest.skip("reference not staged")
    import modules as ours
    missing = []
    for cls in ("VAE", "LSTMEncoder", "LSTMDecoder", "ResNetEncoderV2", "PixelCNNDecoderV2"):
        rc, oc = getattr(ref, cls, None), getattr(ours, cls, None)
        if rc is None:
            continue
        assert oc is not None, cls
        own = {n for n, v in vars(rc).items() if callable(v) and not n.startswith("_")}
        for base in rc.__mro__[1:]:
            if base.__module__.startswith("ref_modules"):
                own |= {n for n, v in vars(base).items() if callable(v) and not n.startswith("_")}
        missing += ["%s.%s" % (cls, n) for n in sorted(own) if not hasattr(oc, n)]
    assert not missing, missing
