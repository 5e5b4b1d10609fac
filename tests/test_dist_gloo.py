"""World-size-2 gloo test of the multi-rank host logic: slice partition, in-place slice
all-gather and the error reduction.  CPU only."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from toycluster_b200 import dist as tdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi, chunk = tdist.rank_slice(n, rank, world)
    # every rank knows the whole array (replicated positions) but only "computes" its slice
    truth = np.arange(n * 4, dtype=np.float32).reshape(n, 4) * 0.5 + 1
    full = torch.zeros(world * chunk * 4, dtype=torch.float32)
    full[lo * 4:hi * 4] = torch.from_numpy(truth[lo:hi].ravel())
    tdist.allgather_slices(full, rank, chunk, 4)
    ok = np.array_equal(full.numpy()[:n * 4].reshape(n, 4), truth)
    err = np.linspace(0, 1, n)
    emax, emean = tdist.reduce_errors(float(err[lo:hi].sum()), float(err[lo:hi].max()), hi - lo)
    ok &= abs(emean - err.mean()) < 1e-12 and emax == err.max()
    q.put((rank, bool(ok), lo, hi, chunk))
    dist.destroy_process_group()


def test_slices_cover_and_align():
    for n in (1, 31, 32, 33, 4099, 10_000_000):
        for world in (1, 2, 4, 8):
            spans = [tdist.rank_slice(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and max(s[1] for s in spans) == n
            for r in range(world - 1):
                assert spans[r][1] == spans[r + 1][0] or spans[r + 1][0] == n
            assert all(s[2] % 32 == 0 and (s[0] % 32 == 0 or s[0] == n) for s in spans)
            assert spans[0][2] * world >= n


def test_two_rank_exchange_and_reduction():
    world, n = 2, 4099                     # ragged: the last slice is short
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert res[0][2] == 0 and res[1][3] == n and res[0][3] == res[1][2]
