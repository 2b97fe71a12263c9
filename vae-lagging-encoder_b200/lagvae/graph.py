"""One training step as ONE CUDA graph.

The aggressive inner loop repeats the same statement sequence (zero_grad, `vae.loss`, backward, clip_grad_norm_, encoder
optimizer step — text.py:373-387 / image.py:300-314) on same-shaped batches; for the image model that is ~800 kernels of a
few microseconds each plus ~120 autograd nodes, and the eager loop is host-bound (profiles/README.md: 13.2 ms eager vs
8.7 ms replayed).  `GraphedStep` captures the sequence once — the liblagvae.so launches go to the capturing stream like any
torch op — and replays it per step: inputs are copied into static tensors, outputs are read from static tensors.

Constraints (those of torch.cuda.graph): fixed shapes; optimizers constructed with `capturable=True`; nothing inside the
body may synchronise (read Σloss from the returned static tensor AFTER the replay, as text.py:381 does with `.item()`).

Two more, specific to this library:
* a captured graph holds RAW device pointers into the engine workspace / the image scratch buffers.  Those buffers grow on
  demand (a larger eager call later, e.g. `nll_iw` on B*ns rows); while any GraphedStep is alive a buffer that is outgrown is
  RETIRED (kept allocated, `retire()` below) instead of freed, so a replay never writes freed memory;
* host scalars are frozen at capture.  The Philox dropout seed of the text decoder is one: a captured train()-mode step
  therefore keys its masks by `seed + *seed_dev` (include/lagvae.h, lagvae_dropout.seed_dev) — `philox_word(device)` is
  that device word and `bump_philox_word` advances it INSIDE the graph by the same constant an eager call advances the host
  seed by, so replay k draws exactly the masks of the k-th eager call (tests/test_gpu_text_graph.py);
* the decoder-weight operand cache (lagvae_text_decoder_weights_epoch) is decided on the host, i.e. at capture: inside a
  GraphedStep capture the engine declares one epoch per capture (`capture_serial()`), so the FIRST captured step re-splits
  the decoder weights at every replay and the following steps of the same graph reuse them (a 15-step window of
  text.py:371-400 splits once); under a bare torch.cuda.graph the cache is off (epoch 0)."""
import weakref
from typing import Callable, Dict, Sequence, Union

import torch

_LIVE = weakref.WeakSet()      # GraphedStep objects that are alive
_RETIRED = []                  # outgrown buffers kept allocated while a graph may still reference them
_SERIAL = [0, False]           # [captures started so far, a GraphedStep capture is in progress]
_WORDS = {}                    # device index -> int64 [1] Philox offset word
PHILOX_STEP = 0xD1B54A32D192ED03   # what one eager decoder forward adds to the host seed (modules/text.py::dropout_spec)


def capture_serial():
    """Non-zero and constant while ONE GraphedStep capture is in progress, None otherwise."""
    return _SERIAL[0] if _SERIAL[1] else None


def philox_word(device):
    """The device word a captured step adds to its (frozen) Philox seed; one per device, shared by all graphs."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _WORDS:
        _WORDS[idx] = torch.zeros(1, dtype=torch.int64, device=torch.device("cuda", idx))
    return _WORDS[idx]


def bump_philox_word(word, times=1):
    """Enqueue word += times * PHILOX_STEP (mod 2^64) on the current stream — call it inside the captured body, before the
    forward that uses the word."""
    inc = (PHILOX_STEP * int(times)) & (2 ** 64 - 1)
    word.add_(inc - 2 ** 64 if inc >= 2 ** 63 else inc)


def retire(t):
    """Called by the buffer owners (TextEngine.plan, modules.image._scratch) INSTEAD of dropping an outgrown buffer."""
    if len(_LIVE) > 0 and t is not None:
        _RETIRED.append(t)
    elif len(_LIVE) == 0:
        _RETIRED.clear()



class GraphedStep:
    def __init__(self, body: Callable[..., Union[torch.Tensor, Sequence[torch.Tensor]]], example_inputs: Dict[str, torch.Tensor],
                 warmup: int = 3):
        """body(**static_inputs) runs one full step and returns the tensor(s) to read back (e.g. loss.sum()).  It is run
        `warmup` times eagerly on a side stream (allocator / lazy-init warm-up, as the torch.cuda.graph recipe asks — these
        ARE real optimisation steps), then captured once."""
        self.static_in = {k: v.clone() for k, v in example_inputs.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                body(**self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        _SERIAL[0] += 1
        _SERIAL[1] = True
        try:
            with torch.cuda.graph(self.graph):
                out = body(**self.static_in)
        finally:
            _SERIAL[1] = False
        self._single = isinstance(out, torch.Tensor)
        self.static_out = [out] if self._single else list(out)
        _LIVE.add(self)

    def __call__(self, **inputs: torch.Tensor):
        for k, v in inputs.items():
            self.static_in[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.static_out[0] if self._single else self.static_out
