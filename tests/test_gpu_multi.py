"""Two real GPUs over NCCL (skipped on a one-GPU box): the partitioned run, driven exactly
like the driver drives bench.py under torchrun, ends in the same state as one GPU."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench(extra, launcher=()):
    cmd = list(launcher) + ["bench.py", "--workload", "merger_1e6", "--n-gas", "200000", "--steps", "3",
                            "--warmup", "2", "--no-cpu-baseline"] + extra
    out = subprocess.run([sys.executable] + cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_two_ranks_equal_one_rank():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    one = _bench(["--gpus", "1"])
    two = _bench(["--gpus", "2"], ["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                                   "--master-addr", "127.0.0.1", "--master-port", "29533"])
    assert two["n_gpus"] == 2 and one["n_gpus"] == 1
    assert one["e2e"]["state_checksum"] == two["e2e"]["state_checksum"]
    assert one["pair_evals_per_particle"] == two["pair_evals_per_particle"]
