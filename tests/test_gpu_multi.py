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


def _two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")


def test_group_context_equals_one_gpu():
    """tg_config.ngpus = 2: ONE process, the library partitions the targets over two devices,
    exchanges the slices and reduces the error statistics itself (NCCL inside libtoygpu.so).
    Sequential mode: every printed number and the final state equal the one-GPU run bit for bit,
    through the AoS operator boundary as well."""
    _two_gpus()
    import numpy as np
    import toycluster_b200 as tc
    from toycluster_b200 import workloads
    w = workloads.make("merger_1e6", n_gas=60000)
    res = {}
    for name, kw in (("one", {}), ("group", {"ngpus": 2})):
        g = tc.HotPath.from_workload(w, flags=tc.WVT_SEQUENTIAL, **kw)
        g.upload(w.pos)
        done, rows = g.regularise_sph_particles(max_iters=6)
        g.find_sph_quantities()
        res[name] = (rows, g.download(), g.stats())
        g.close()
    assert res["one"][0] == res["group"][0]                     # the '#NN: Err ...' numbers
    for k in ("id", "pos", "hsml", "rho", "varhsml", "rho_model"):
        assert np.array_equal(res["one"][1][k], res["group"][1][k]), k
    assert res["one"][2]["pair_evals"] == res["group"][2]["pair_evals"]


def test_group_context_records_round_trip():
    _two_gpus()
    import numpy as np
    import toycluster_b200 as tc
    from toycluster_b200 import workloads
    w = workloads.make("merger_1e6", n_gas=50003)
    n = w.n_gas
    Pdt = np.dtype([("Pos", "3f4"), ("Vel", "3f4"), ("ID", "i4"), ("Type", "i4"),
                    ("Key", "2u8"), ("Tree_Parent", "i4"), ("pad", "3i4")])
    Sdt = np.dtype([("U", "f4"), ("Rho", "f4"), ("Hsml", "f4"), ("VarHsmlFac", "f4"),
                    ("Bfld", "3f4"), ("Apot", "3f4"), ("ID", "f4"), ("Rho_Model", "f4"), ("Rs", "3f4")])
    out = {}
    for name, kw in (("one", {}), ("group", {"ngpus": 2})):
        P, S = np.zeros(n, Pdt), np.zeros(n, Sdt)
        P["Pos"], P["ID"], S["U"] = w.pos, np.arange(n) * 3 + 1, np.arange(n) * 0.25
        g = tc.HotPath.from_workload(w, **kw)
        g.upload_records(P, S)
        g.regularise_sph_particles(max_iters=3)
        g.download_records(P, S)
        out[name] = (P.copy(), S.copy())
        g.close()
    assert out["one"][0].tobytes() == out["group"][0].tobytes()
    assert out["one"][1].tobytes() == out["group"][1].tobytes()


def test_driver_with_two_gpus_writes_the_same_file(tmp_path):
    """The reference's whole, unmodified C driver on top of gpu_shim.c with TOYGPU_NGPUS=2."""
    _two_gpus()
    from test_driver_e2e import GPU, PAR, read_gadget2, run
    if not os.path.exists(GPU):
        pytest.skip("oracle/_ref drivers not built")
    for tag in ("a", "b"):
        (tmp_path / f"{tag}.par").write_text(PAR.format(out=f"IC_{tag}", ntotal=40000, mass_ratio=0.3125, bnorm="20e-6"))
    out_a = run(GPU, "a.par", tmp_path, {"TOYGPU_FLAGS": "1"})
    out_b = run(GPU, "b.par", tmp_path, {"TOYGPU_FLAGS": "1", "TOYGPU_NGPUS": "2"})
    it_a = [l for l in out_a.splitlines() if l.lstrip().startswith("#")]
    it_b = [l for l in out_b.splitlines() if l.lstrip().startswith("#")]
    assert len(it_a) >= 3 and it_a == it_b
    a, b = read_gadget2(tmp_path / "IC_a"), read_gadget2(tmp_path / "IC_b")
    assert list(a) == list(b)
    for label in a:
        assert a[label] == b[label], label
