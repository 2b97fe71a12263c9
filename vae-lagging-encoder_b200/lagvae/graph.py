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
* host scalars are frozen at capture.  That includes the Philox dropout seed of the text decoder: capturing a text model in
  train() mode with in-kernel dropout would replay ONE mask for ever, so `modules.text.dropout_spec` refuses to build a Philox
  spec while a stream is capturing (use eval(), p = 0, or LAGVAE_DROPOUT=torch masks drawn outside the graph)."""
import weakref
from typing import Callable, Dict, Sequence, Union

import torch

_LIVE = weakref.WeakSet()      # GraphedStep objects that are alive
_RETIRED = []                  # outgrown buffers kept allocated while a graph may still reference them


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
        with torch.cuda.graph(self.graph):
            out = body(**self.static_in)
        self._single = isinstance(out, torch.Tensor)
        self.static_out = [out] if self._single else list(out)
        _LIVE.add(self)

    def __call__(self, **inputs: torch.Tensor):
        for k, v in inputs.items():
            self.static_in[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.static_out[0] if self._single else self.static_out
