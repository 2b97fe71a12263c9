#!/usr/bin/env python
"""Run an UNMODIFIED reference driver (text.py / image.py / toy.py) against the B200 `modules` backend.

    python scripts/run_reference_driver.py [--backend lagvae|reference] [--extra-path DIR]... \
        /path/to/reference/text.py --dataset yahoo --aggressive 1 ...

The reference directory supplies `data`, `config`, `logger` (and, with --backend reference, its own `modules`: the
same script on the stock torch path, for side-by-side logs); this repo supplies `modules`.
--extra-path directories go in front of sys.path after the backend (the reference's `config` is a namespace package,
so a directory holding `config/config_<name>.py` adds a dataset configuration without touching the reference).
The working directory (datasets/, models/, logs/) is $LAGVAE_RUN_DIR or the current directory."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    argv = sys.argv[1:]
    backend, extra = "lagvae", []
    while argv and argv[0].startswith("--"):
        if argv[0] == "--backend" and len(argv) > 1:
            backend, argv = argv[1], argv[2:]
        elif argv[0] == "--extra-path" and len(argv) > 1:
            extra.append(os.path.abspath(argv[1]))
            argv = argv[2:]
        else:
            break
    if not argv or backend not in ("lagvae", "reference"):
        sys.exit(__doc__)
    script = os.path.abspath(argv[0])
    ref_dir = os.path.dirname(script)
    sys.argv = [script] + argv[1:]
    front = [os.path.join(ROOT, "vae-lagging-encoder_b200")] if backend == "lagvae" else []
    sys.path[:0] = front + extra + [ref_dir]
    os.chdir(os.environ.get("LAGVAE_RUN_DIR", os.getcwd()))
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
