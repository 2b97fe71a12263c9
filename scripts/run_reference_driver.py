#!/usr/bin/env python
"""Run an UNMODIFIED reference driver (text.py / toy.py) against the B200 `modules` backend.

    python scripts/run_reference_driver.py /path/to/reference/text.py --dataset yahoo --aggressive 1 ...

The reference directory supplies `data`, `config`, `logger`; this repo supplies `modules`."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    script = os.path.abspath(sys.argv[1])
    ref_dir = os.path.dirname(script)
    sys.argv = [script] + sys.argv[2:]
    sys.path[:0] = [os.path.join(ROOT, "vae-lagging-encoder_b200"), ref_dir]
    os.chdir(os.environ.get("LAGVAE_RUN_DIR", os.getcwd()))
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
