#!/usr/bin/env python
"""Run an UNMODIFIED reference driver (text.py / image.py / toy.py) against the B200 `modules` backend.

    python scripts/run_reference_driver.py [--backend lagvae|reference] [--extra-path DIR]... \
        /path/to/reference/text.py --dataset yahoo --aggressive 1 ...

The reference directory supplies `data`, `config`, `logger` (and, with --backend reference, its own `modules`: the
same script on the stock torch path, for side-by-side logs); this repo supplies `modules`.
--extra-path directories go in front of sys.path after the backend (the reference's `config` is a namespace package,
so a directory holding `config/config_<name>.py` adds a dataset configuration without touching the reference).
The working directory (datasets/, models/, logs/) is $LAGVAE_RUN_DIR or the current directory.

Several GPUs: launch this shim with torchrun (`python -m torch.distributed.run --nproc-per-node N
scripts/run_reference_driver.py .../text.py ...`).  It initialises the NCCL process group, binds rank r to GPU r and gives
every rank > 0 its own working directory `<run dir>/rank<r>` (with `datasets` linked in) so that checkpoints and logs do
not collide; the unmodified script then runs SPMD with identical seeds and `VAE.loss` shards every batch by rank."""
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
    run_dir = os.environ.get("LAGVAE_RUN_DIR", os.getcwd())
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", str(rank)))
        torch.cuda.set_device(local)                       # the drivers use torch.device("cuda") = the current device
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        if rank > 0:
            sub = os.path.join(run_dir, "rank%d" % rank)
            os.makedirs(sub, exist_ok=True)
            link = os.path.join(sub, "datasets")
            if not os.path.exists(link):
                os.symlink(os.path.join(run_dir, "datasets"), link)
            run_dir = sub
    os.chdir(run_dir)
    try:
        runpy.run_path(script, run_name="__main__")
    finally:
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.barrier()
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
