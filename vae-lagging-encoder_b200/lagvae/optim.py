"""Host side of the two boundary entry points beyond the text plan (include/lagvae.h): the fused clip + Adam step over a
device-resident parameter table (image.py:312-314) and the library-owned NCCL communicator for the flat gradient bucket
(SURVEY §8 e1).  PyTorch is plumbing (device memory, streams, the out-of-band exchange of the NCCL id)."""
import ctypes as C
import os

import torch

from . import _backend as be


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class ClipAdam:
    """clip_grad_norm_(all `params`, max_norm) + Adam on the first `n_update` of them, semantics of torch.optim.Adam with its
    defaults (betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad).  `grads` are caller-owned gradient tensors that pair
    up with `params` by index (e.g. views of a flat bucket); they are rescaled in place like clip_grad_norm_ does.  The step
    is three kernel launches and can be captured in a CUDA graph (lagvae.GraphedStep): Adam's step count lives on the device."""

    def __init__(self, params, grads, n_update, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, max_norm=5.0, initial_step=0):
        n = len(params)
        assert len(grads) == n and 0 <= n_update <= n
        for p, g in zip(params, grads):
            if p.device.type != "cuda" or p.dtype != torch.float32 or not p.is_contiguous() or g.shape != p.shape or not g.is_contiguous():
                raise be.LagvaeError("ClipAdam needs contiguous fp32 CUDA parameters and matching gradients (no CPU path)")
        self.params, self.grads, self.n_update = list(params), list(grads), int(n_update)
        self.exp_avg = [torch.zeros_like(p) for p in params[:n_update]]
        self.exp_avg_sq = [torch.zeros_like(p) for p in params[:n_update]]
        self.lr, self.betas, self.eps, self.max_norm = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(max_norm)
        dev = params[0].device
        L = be.lib()
        arr = lambda ts: (C.c_void_p * n)(*([t.data_ptr() for t in ts] + [None] * (n - len(ts))))
        cnt = (C.c_int64 * n)(*[p.numel() for p in params])
        nbytes = int(L.lagvae_adam_table_bytes(cnt, n))
        self._mem = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.norm = torch.zeros(1, dtype=torch.float32, device=dev)
        h = C.c_void_p()
        with torch.cuda.device(dev):
            be.check(L.lagvae_adam_table_create(arr(self.params), arr(self.grads), arr(self.exp_avg), arr(self.exp_avg_sq), cnt, n,
                                                self.n_update, int(initial_step), be.ptr(self._mem), nbytes, _stream(), C.byref(h)),
                     "lagvae_adam_table_create")
        self._h, self._dev = h, dev

    def step(self, scale_all=True):
        """Returns the device tensor holding the pre-clip total gradient norm."""
        with torch.cuda.device(self._dev):
            be.check(be.lib().lagvae_clip_adam_step(self._h, self.max_norm, self.lr, self.betas[0], self.betas[1], self.eps,
                                                    1 if scale_all else 0, be.ptr(self.norm), _stream()), "lagvae_clip_adam_step")
        return self.norm

    def __del__(self):
        try:
            be.lib().lagvae_adam_table_destroy(self._h)
        except Exception:
            pass


def _find_libnccl():
    cands = [os.environ.get("LAGVAE_NCCL_LIB")]
    try:
        import nvidia.nccl
        cands.append(os.path.join(os.path.dirname(nvidia.nccl.__file__), "lib", "libnccl.so.2"))
    except Exception:
        pass
    cands += ["libnccl.so.2"]
    for c in cands:
        if c and (os.path.sep not in c or os.path.exists(c)):
            return c
    raise be.LagvaeError("libnccl.so.2 not found (set LAGVAE_NCCL_LIB)")


class BucketComm:
    """Library-owned NCCL communicator over the ranks of a torch.distributed group (used only to hand the 128-byte NCCL id
    from rank 0 to the others).  all_reduce(flat) = lagvae_allreduce_bucket on the current stream."""

    def __init__(self, group=None):
        import torch.distributed as dist
        L = be.lib()
        be.check(L.lagvae_comm_load(_find_libnccl().encode()), "lagvae_comm_load")
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        uid = (C.c_uint8 * 128)()
        if self.rank == 0:
            be.check(L.lagvae_comm_unique_id(uid), "lagvae_comm_unique_id")
        box = [bytes(uid)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        buf = (C.c_uint8 * 128).from_buffer_copy(box[0])
        h = C.c_void_p()
        be.check(L.lagvae_comm_init(buf, self.rank, self.world, C.byref(h)), "lagvae_comm_init")
        self._h = h

    def all_reduce(self, flat, stream=None):
        if flat.dtype != torch.float32 or not flat.is_contiguous() or flat.device.type != "cuda":
            raise be.LagvaeError("BucketComm.all_reduce needs a contiguous fp32 CUDA tensor")
        st = C.c_void_p((stream or torch.cuda.current_stream()).cuda_stream)
        be.check(be.lib().lagvae_allreduce_bucket(self._h, be.ptr(flat), flat.numel(), st), "lagvae_allreduce_bucket")

    def close(self):
        if self._h is not None:
            be.lib().lagvae_comm_destroy(self._h)
            self._h = None
