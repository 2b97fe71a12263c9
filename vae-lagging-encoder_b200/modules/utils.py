"""Numerics helpers with the reference's names and semantics (modules/utils.py:3-36)."""
import torch


def log_sum_exp(value, dim=None, keepdim=False):
    """Stable log(sum(exp(value))) — same contract as reference modules/utils.py:3-16."""
    if dim is None:
        return torch.logsumexp(value.reshape(-1), dim=0)
    return torch.logsumexp(value, dim=dim, keepdim=keepdim)


def generate_grid(zmin, zmax, dz, device, ndim=2):
    """1-D / 2-D evaluation grid (reference modules/utils.py:19-36; toy.py plotting only)."""
    axis = torch.arange(zmin, zmax, dz)
    if ndim == 1:
        return axis.unsqueeze(1).to(device)
    if ndim == 2:
        k = axis.numel()
        first = axis.repeat_interleave(k)
        second = axis.repeat(k)
        return torch.stack((first, second), dim=-1).to(device), k
    raise ValueError("ndim must be 1 or 2")
