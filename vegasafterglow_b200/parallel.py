"""Walker partition across the GPUs of one box.

Every walker / parameter set is an independent model evaluation, so a batch is split into
contiguous blocks, one per rank (one process per GPU), with NO data-path collective; the only
inter-GPU traffic is the gather of the float64[n_walkers] chi-squared vector
(``torch.distributed`` all_gather over NCCL on GPUs -- NVLink/NVSwitch --, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def partition(n: int, world: int, rank: int):
    """Contiguous block [lo, hi) of ``n`` walkers owned by ``rank``: ceil(n/world) per rank."""
    per = -(-n // world) if world > 0 else n
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def gather_blocks(local: np.ndarray, n: int, device=None) -> np.ndarray:
    """All-gather the per-rank float64 blocks of ``partition`` back into a length-``n`` vector."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(local, dtype=np.float64)
    world = dist.get_world_size()
    per = -(-n // world)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    buf = torch.full((per,), float("inf"), dtype=torch.float64, device=device)
    buf[: len(local)] = torch.as_tensor(np.asarray(local, dtype=np.float64), device=device)
    out = torch.empty(per * world, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, buf)
    out = out.cpu().numpy()
    pieces = [out[r * per: r * per + (partition(n, world, r)[1] - partition(n, world, r)[0])] for r in range(world)]
    return np.concatenate(pieces)


def partitioned_chi2(engine, params, t, nu, lnF, sig, w, evaluate=None) -> np.ndarray:
    """chi2 of the whole ensemble, each rank evaluating its block on its own GPU.

    ``evaluate(params_block) -> chi2_block`` defaults to ``engine.chi2_series``; the CPU tests pass
    a host evaluator to exercise the partition/gather logic under gloo."""
    import torch.distributed as dist

    n = len(params)
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(), dist.get_rank()
    else:
        world, rank = 1, 0
    lo, hi = partition(n, world, rank)
    block = params[lo:hi]
    if evaluate is None:
        evaluate = lambda P: engine.chi2_series(P, t, nu, lnF, sig, w)  # noqa: E731
    local = evaluate(block) if hi > lo else np.zeros(0)
    return gather_blocks(local, n)
