"""The drop-in claim at script level (SURVEY §4 test plan item 4, §8 b1): the UNMODIFIED reference drivers `text.py`
(text.py:524-528), `image.py` (image.py:452-454) and `toy.py` run end to end against the B200 `modules` back-end through
scripts/run_reference_driver.py on small synthetic data, and the numbers they print (avg_loss / kl / mi / recon per
logged iteration, VAL / TEST lines, iw nll) agree with the same script on the reference's own `modules`:

* same GPU, same seeds, dropout-free configuration (`tinysyn`, `tinyomni`): both back-ends consume torch's and numpy's
  generators identically, so the logs are compared line by line (tolerances below: fp32 noise amplified by SGD lr 1.0);
* dropout 0.5 (`tinysyndrop`, toy.py's shipped config): the masks differ (torch bernoulli vs in-kernel Philox), so the
  comparison is statistical, against the logs the reference back-end produced on the CPU (tests/golden/drivers/*.log,
  written by this file's `--make-fixtures` entry in the authoring container).

The reference scripts are read from $VAE_REF_PATH, baseline/_ref (scripts/stage_reference.sh; git-ignored, travels to the
GPU box) or /root/reference; without them the tests skip — they never touch the hot path's product code with the oracle.
"""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRV = os.path.join(ROOT, "tests", "drivers")
GOLD = os.path.join(ROOT, "tests", "golden", "drivers")
KEEP = re.compile(r"^(epoch:|VAL|TEST|pre mi|STOP|kl weight|iw nll:|[0-9]+ active)")
NUM = re.compile(r"(avg_loss|kl|mi|recon|nll|ppl|iw nll|iw ppl|pre mi|cur mi|au|kl weight)[: ]+(-?[0-9.]+(?:e-?[0-9]+)?)")


def ref_dir():
    for c in (os.environ.get("VAE_REF_PATH"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if c and os.path.exists(os.path.join(c, "text.py")) and os.path.isdir(os.path.join(c, "modules")):
            return c
    return None


def make_workdir(path):
    sys.path.insert(0, DRV)
    import make_data
    os.makedirs(path, exist_ok=True)
    make_data.make_text(path)
    make_data.make_image(path)
    syn = os.path.join(path, "datasets", "synthetic_data")          # toy.py's shipped config (config_synthetic.py:13-15)
    os.makedirs(syn, exist_ok=True)
    src = os.path.join(path, "datasets", "tinysyn_data")
    for a, b in (("train.txt", "synthetic_train.txt"), ("valid.txt", "synthetic_test.txt")):
        with open(os.path.join(src, a)) as f, open(os.path.join(syn, b), "w") as g:
            g.write(f.read())
    return path


def run_driver(script, backend, args, workdir, timeout, cpu=False):
    env = dict(os.environ, LAGVAE_RUN_DIR=workdir, PYTHONWARNINGS="ignore")
    if cpu:
        env["CUDA_VISIBLE_DEVICES"] = ""
    cmd = [sys.executable, os.path.join(ROOT, "scripts", "run_reference_driver.py"), "--backend", backend,
           "--extra-path", DRV, os.path.join(ref_dir(), script)] + args
    try:
        r = subprocess.run(cmd, env=env, cwd=workdir, capture_output=True, text=True, timeout=timeout)
        out, rc = r.stdout + "\n" + r.stderr, r.returncode
    except subprocess.TimeoutExpired as e:       # time-boxed runs: the printed prefix is what gets compared
        out = (e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")) + "\n[timeout]"
        rc = -9
    return rc, out


def parse(out):
    """[(kind, {field: value})] for every result line the drivers print."""
    rows = []
    for ln in out.splitlines():
        if not KEEP.match(ln) or ln.strip().endswith(("VAL", "TEST")) and "---" not in ln:
            continue
        kind = ln.split()[0].rstrip(":,")
        if kind == "epoch":
            m = re.match(r"epoch: (\d+), iter: (\d+)", ln)
            kind = "it%s" % m.group(2) if m else "epoch"
        if ln.startswith("iw nll"):
            kind = "iw"
        if re.match(r"^[0-9]+ active", ln):
            rows.append(("active", {"n": float(ln.split()[0])}))
            continue
        vals = {k.replace(" ", "_"): float(v.rstrip(".")) for k, v in NUM.findall(ln)}
        rows.append((kind, vals))
    return rows


def save(name, out):
    d = os.path.join(ROOT, "gpurun_out", "drivers")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, name), "w") as f:
        f.write(out)


def compare_logs(a_rows, b_rows, what, n_tight, rel_tight, kl_abs=3e-2, mi_abs=0.1, sanity=True):
    """Two logs of the same script.  The first `n_tight` logged training iterations (each closes an aggressive inner loop of
    up to 99 encoder updates) must agree to `rel_tight`.  After that the runs are different samples of a CHAOTIC, spiky
    trajectory (SGD lr 1.0 + clip on a 160-sentence corpus, and the data-dependent break rule of text.py:393-398: one
    flipped comparison changes the number of inner steps and with it every later batch pick).  Measured on the B200
    (profiles/r2_driver_logs/): both back-ends print the same 4 decimals for the first 6 iterations and stay within 1e-3
    through iteration 8, while the REFERENCE back-end on the GPU vs the REFERENCE back-end on the CPU already differ by
    60 % at the first VAL line (69.0 vs 43.0).  So beyond the prefix only structure and sanity are asserted: the same
    kinds of lines, finite numbers, and a best validation NLL below the first logged training loss (training worked)."""
    ia = [(k, v) for k, v in a_rows if k.startswith("it")]
    ib = [(k, v) for k, v in b_rows if k.startswith("it")]
    assert [k for k, _ in ia] == [k for k, _ in ib], "%s: different iteration lines" % what
    bad = []
    for (k, x), (_, y) in list(zip(ia, ib))[:n_tight]:
        for f in ("avg_loss", "recon"):
            if abs(x[f] - y[f]) > rel_tight * max(abs(y[f]), 1.0):
                bad.append("%s %s: %.4f vs %.4f" % (k, f, x[f], y[f]))
        if abs(x["kl"] - y["kl"]) > kl_abs + rel_tight * abs(y["kl"]):
            bad.append("%s kl: %.4f vs %.4f" % (k, x["kl"], y["kl"]))
        if "mi" in x and "mi" in y and mi_abs is not None and abs(x["mi"] - y["mi"]) > mi_abs:
            bad.append("%s mi: %.4f vs %.4f" % (k, x["mi"], y["mi"]))
    for kind in ("VAL", "TEST", "iw"):
        va, vb = [v for k, v in a_rows if k == kind], [v for k, v in b_rows if k == kind]
        if len(va) != len(vb):
            bad.append("%s: %d vs %d lines" % (kind, len(va), len(vb)))
        for v in va:
            if not all(x == x and abs(x) < 1e9 for x in v.values()):
                bad.append("%s: non-finite value %r" % (kind, v))
    va = [v for k, v in a_rows if k == "VAL"]
    if sanity and va and ia and not min(v["nll"] for v in va) < ia[0][1]["avg_loss"]:      # the spiky trajectory may END on a spike
        bad.append("best VAL nll %.4f is not below the first logged training loss %.4f" % (min(v["nll"] for v in va), ia[0][1]["avg_loss"]))
    assert not bad, "%s:\n%s" % (what, "\n".join(bad[:40]))


needs_ref = pytest.mark.skipif(ref_dir() is None, reason="reference scripts not staged (scripts/stage_reference.sh)")
TEXT_ARGS = ["--aggressive", "1", "--warm_up", "10", "--kl_start", "0.1", "--iw_nsamples", "100"]


@pytest.mark.gpu
@needs_ref
def test_unmodified_text_py_same_gpu_line_by_line(tmp_path):
    """text.py --dataset tinysyn --aggressive 1 --warm_up 10 --kl_start 0.1 (README.md:53 with the test configuration):
    2 epochs, aggressive inner loops, val-MI switch-off rule, VAL/TEST evaluation, checkpoint save + reload, iw nll."""
    wd = make_workdir(str(tmp_path / "run"))
    rc_a, out_a = run_driver("text.py", "lagvae", ["--dataset", "tinysyn"] + TEXT_ARGS, wd, 900)
    save("text_tinysyn_lagvae_gpu.log", out_a)
    assert rc_a == 0, out_a[-3000:]
    rc_b, out_b = run_driver("text.py", "reference", ["--dataset", "tinysyn"] + TEXT_ARGS, wd, 900)
    save("text_tinysyn_reference_gpu.log", out_b)
    assert rc_b == 0, out_b[-3000:]
    a, b = parse(out_a), parse(out_b)
    assert len(a) >= 30 and any(k == "iw" for k, _ in a)
    # measured on the B200: the first 6 logged iterations (~600 encoder updates) print the same 4 decimals on both back-ends
    compare_logs(a, b, "text.py tinysyn (same GPU)", n_tight=8, rel_tight=1e-3)
    # and against the log the reference produced on the CPU (different eps stream from the first draw on)
    with open(os.path.join(GOLD, "text_tinysyn_reference_cpu.log")) as f:
        c = parse(f.read())
    compare_logs(a, c, "text.py tinysyn (CPU reference log)", n_tight=2, rel_tight=2e-2, kl_abs=0.5, mi_abs=None)


@pytest.mark.gpu
@needs_ref
def test_unmodified_text_py_with_dropout_trajectory(tmp_path):
    """Dropout 0.5 / 0.5 (the shipped text configurations): in-kernel Philox masks vs torch's — statistical agreement with
    the log of the reference back-end (CPU, committed)."""
    wd = make_workdir(str(tmp_path / "run"))
    rc, out = run_driver("text.py", "lagvae", ["--dataset", "tinysyndrop"] + TEXT_ARGS, wd, 900)
    save("text_tinysyndrop_lagvae_gpu.log", out)
    assert rc == 0, out[-3000:]
    a = parse(out)
    with open(os.path.join(GOLD, "text_tinysyndrop_reference_cpu.log")) as f:
        c = parse(f.read())
    compare_logs(a, c, "text.py tinysyndrop (CPU reference log)", n_tight=2, rel_tight=3e-2, kl_abs=0.5, mi_abs=None)


@pytest.mark.gpu
@needs_ref
def test_unmodified_image_py_same_gpu(tmp_path):
    """image.py --dataset tinyomni --aggressive 1 (ResNet encoder + PixelCNN decoder, Adam, window 10): one epoch of
    5 outer iterations with their aggressive inner loops, VAL/TEST, iw nll — same GPU, same seeds, both back-ends."""
    pytest.importorskip("torchvision")
    wd = make_workdir(str(tmp_path / "run"))
    # nll_iw draws ns = 100 samples per chunk (vae.py:100): iw_nsamples below 100 gives an empty list in the reference too
    args = ["--dataset", "tinyomni", "--aggressive", "1", "--warm_up", "10", "--kl_start", "0.1", "--iw_nsamples", "100"]
    rc_a, out_a = run_driver("image.py", "lagvae", args, wd, 1500)
    save("image_tinyomni_lagvae_gpu.log", out_a)
    assert rc_a == 0, out_a[-3000:]
    rc_b, out_b = run_driver("image.py", "reference", args, wd, 1500)
    save("image_tinyomni_reference_gpu.log", out_b)
    assert rc_b == 0, out_b[-3000:]
    a, b = parse(out_a), parse(out_b)
    assert len(a) >= 6
    # Adam (lr 1e-3) on BatchNorm'ed PixelCNN activations: the first logged iteration closes one aggressive inner loop
    compare_logs(a, b, "image.py tinyomni (same GPU)", n_tight=1, rel_tight=5e-3, kl_abs=0.2, mi_abs=0.3)


@pytest.mark.gpu
@needs_ref
def test_unmodified_toy_py_runs_and_tracks_the_reference(tmp_path):
    """toy.py --aggressive 1 --plot_mode multiple (BASELINE.json configs[0]; nz = 1, ni = nh = 50: the fp32 SIMT tier),
    time-boxed: the printed prefix must follow the reference's CPU log at trajectory level (dropout 0.5: statistical)."""
    wd = make_workdir(str(tmp_path / "run"))
    rc, out = run_driver("toy.py", "lagvae", ["--aggressive", "1", "--plot_mode", "multiple", "--num_plot", "32",
                                              "--plot_niter", "10", "--iw_nsamples", "100"], wd, 420)
    save("toy_lagvae_gpu.log", out)
    assert rc in (0, -9), out[-3000:]
    a = parse(out)
    with open(os.path.join(GOLD, "toy_reference_cpu.log")) as f:
        c = parse(f.read())
    va, vc = [r[1] for r in a if r[0] == "VAL"], [r[1] for r in c if r[0] == "VAL"]
    assert len(va) >= 2, "toy.py did not reach the second validation pass inside the time box"
    # at the stock initialisation dropout barely matters: the first logged iterations agree to ~1e-4 (measured), then the
    # runs are independent samples of the trajectory
    ia = [v for k, v in a if k.startswith("it")][:5]
    ic = [v for k, v in c if k.startswith("it")][:5]
    for x, y in zip(ia, ic):
        assert abs(x["avg_loss"] - y["avg_loss"]) <= 2e-3 * y["avg_loss"], (x, y)
    assert all(v["nll"] == v["nll"] and v["nll"] < 1e6 for v in va)
    assert va[-1]["nll"] < ia[0]["avg_loss"]                 # training made progress (first logged loss 42.7)
    assert os.path.isdir(os.path.join(wd, "plot_data", "multiple")) and os.listdir(os.path.join(wd, "plot_data", "multiple"))


def test_driver_log_parser_on_the_committed_fixtures():
    """CPU: the parser used above reads the committed reference logs (41 / 451 result lines)."""
    for name, n_val in (("text_tinysyn_reference_cpu.log", 2), ("text_tinysyndrop_reference_cpu.log", 2)):
        with open(os.path.join(GOLD, name)) as f:
            rows = parse(f.read())
        assert sum(1 for k, _ in rows if k == "VAL") == n_val
        assert sum(1 for k, _ in rows if k.startswith("it")) == 28
        assert "iw_nll" in dict(rows)["iw"]
        assert all("avg_loss" in v and "recon" in v for k, v in rows if k.startswith("it"))


if __name__ == "__main__" and "--make-fixtures" in sys.argv:
    # authoring container only (reference back-end on the CPU): rewrites tests/golden/drivers/*.log
    import tempfile
    wd = make_workdir(tempfile.mkdtemp(prefix="lagvae_drv_"))
    os.makedirs(GOLD, exist_ok=True)
    jobs = [("text_tinysyn", "text.py", ["--dataset", "tinysyn"] + TEXT_ARGS),
            ("text_tinysyndrop", "text.py", ["--dataset", "tinysyndrop"] + TEXT_ARGS),
            ("toy", "toy.py", ["--aggressive", "1", "--plot_mode", "multiple", "--num_plot", "32", "--plot_niter", "10",
                               "--iw_nsamples", "100"])]
    for name, script, args in jobs:
        rc, out = run_driver(script, "reference", args, wd, 3600, cpu=True)
        assert rc == 0, out[-2000:]
        with open(os.path.join(GOLD, name + "_reference_cpu.log"), "w") as f:
            f.write("\n".join(ln for ln in out.splitlines() if KEEP.match(ln)) + "\n")
        print(name, "->", len(parse(out)), "result lines")
