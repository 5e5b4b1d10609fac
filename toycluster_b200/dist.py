"""Host-side plumbing for one-process-per-GPU runs.

The path shards by TARGET particle: every rank holds all positions, sorts and indexes them
redundantly (bit-identical on every rank), sweeps the Peano-order slice
``[rank*chunk, min(n, (rank+1)*chunk))`` and then the moved ``(x, y, z, Hsml)`` slices are
re-assembled with ONE all-gather per step; the error statistics of wvt_relax.c:73-87 need one
all-reduce of two scalars.  Works on NCCL (device tensors wrapping the library's buffers) and
on gloo (CPU tensors; used by the world-size-2 tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def rank_slice(n: int, rank: int, nranks: int):
    """Mirror of tg_create's partition: equal chunks of whole 32-target tiles."""
    chunk = ((n + nranks - 1) // nranks + 31) // 32 * 32
    lo = min(n, rank * chunk)
    hi = min(n, lo + chunk)
    return lo, hi, chunk


def allgather_slices(full: torch.Tensor, rank: int, chunk: int, width: int) -> None:
    """In-place all-gather: ``full`` holds nranks*chunk records of ``width`` elements and this
    rank's records are already in place."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    mine = full[rank * chunk * width:(rank + 1) * chunk * width]
    if full.is_cuda:
        dist.all_gather_into_tensor(full, mine)        # NCCL: in place
    else:
        dist.all_gather_into_tensor(full, mine.clone())


def reduce_errors(err_sum: float, err_max: float, count: int, device="cpu"):
    """Global (errMax, errMean) from per-rank sums (wvt_relax.c:73-87)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return err_max, (err_sum / count if count else 0.0)
    s = torch.tensor([err_sum, float(count)], dtype=torch.float64, device=device)
    m = torch.tensor([err_max], dtype=torch.float64, device=device)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return m.item(), s[0].item() / s[1].item()


def regularise(g, exchange, max_iters=1 << 30, mtotal=1e5, log=None, device="cpu"):
    """wvt_relax.c:25-225 for one-process-per-GPU runs: the reference's control flow on
    all-reduced error statistics, around HotPath.wvt_begin / wvt_finish.  ``exchange()``
    re-assembles the moved slices (allgather_slices on the library's state buffer)."""
    step = 0.0085
    if mtotal < 1e5:
        step /= 2
    err_last = err_diff_last = float("inf")
    it, rows = -1, []
    while True:
        it += 1
        if it - 1 >= 64 or it >= max_iters:
            break
        s, m, n = g.wvt_begin(step)
        err_max, err_mean = reduce_errors(s, m, n, device)
        err_diff = (err_last - err_mean) / err_mean
        rows.append(dict(it=it, max=err_max, mean=err_mean, diff=err_diff, step=step))
        stop = bool(log(it, err_max, err_mean, err_diff, step)) if log else False
        if stop or (err_diff < 0.01 and it > 25) or (err_diff < 0 and err_diff_last < 0 and it > 10):
            g.wvt_finish(0.0)
            exchange()
            break
        if err_diff < 0.01 and it > 1:
            step *= 0.8
        err_last, err_diff_last = err_mean, err_diff
        g.wvt_finish(step)
        exchange()
    return rows
