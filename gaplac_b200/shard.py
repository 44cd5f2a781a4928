"""Multi-GPU sharding of independent GP evaluations.

Every batch element (a hyperparameter proposal, a chain, a per-feature model) is an independent Cholesky: the
batch is split into contiguous blocks, one per rank, with no data-path collective; the only exchange is one
all-gather of the per-rank log-densities.  The reference has no counterpart (single process, one chain:
CLI/src/mcmc.jl:41).  `torch.distributed` is plumbing only: NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def shard_range(B: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; the first B % world ranks get one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    q, r = divmod(B, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def shard_sizes(B: int, world: int) -> list[int]:
    return [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]


def gather_lml(local, B: int, group=None):
    """All-gather the per-rank results (1-D torch tensors of this rank's block) into the full length-B vector,
    in batch order, on every rank.  Uneven blocks are padded to the largest block for the collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = shard_sizes(B, world)
    m = max(sizes)
    buf = torch.full((m,), float("nan"), dtype=local.dtype, device=local.device)
    buf[: local.numel()] = local
    out = torch.empty(world * m, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf, group=group) if local.is_cuda else dist.all_gather(
        list(out.view(world, m).unbind(0)), buf, group=group)
    return torch.cat([out[r * m: r * m + sizes[r]] for r in range(world)])


def sharded_logpdf(ctx, prog, X, Y, Theta, sigma2, jitter: float = 0.0, group=None):
    """Evaluate this rank's block of a (B, p) hyperparameter batch on its own GPU and gather the B log-densities.
    X: (n, d) shared or (B, n, d); Y: (n,) shared or (B, n); sigma2: scalar or (B,).  Returns (lml[B], info[B]) as NumPy arrays on every rank."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    Theta = np.atleast_2d(np.asarray(Theta, dtype=np.float64))
    B = Theta.shape[0]
    lo, hi = shard_range(B, rank, world)
    Yl = np.asarray(Y)
    Yl = Yl[lo:hi] if Yl.ndim == 2 else Yl
    s2 = np.atleast_1d(np.asarray(sigma2, dtype=np.float64))
    s2 = s2[lo:hi] if s2.size > 1 else s2
    Xl = np.asarray(X)
    Xl = Xl[lo:hi] if Xl.ndim == 3 else Xl          # per-item inputs (B, n, d) are sharded like Theta
    if hi > lo:
        lml, info = ctx.lml_batched(prog, Xl, Yl, Theta[lo:hi], s2, jitter)
    else:
        lml, info = np.empty(0), np.empty(0, dtype=np.int32)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    g = gather_lml(torch.from_numpy(lml).to(dev), B, group)
    gi = gather_lml(torch.from_numpy(info.astype(np.float64)).to(dev), B, group)
    return g.cpu().numpy(), gi.cpu().numpy().astype(np.int32)
