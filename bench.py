#!/usr/bin/env python
"""bench.py -- WVT relax steps/s and neighbour interactions/s of the SPH/WVT hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of wvt_relax.c:66-214 (Peano sort, index build, WC6 density/hsml solve,
error pass, model hsml, displacement, move) over the whole synthetic particle set.  The
default workload is BASELINE.json's metric configuration: the two-cluster merger with
10 M gas particles (configs[2]); it fits one B200 (about 1.5 GB).  With N > 1 ranks the same
10 M targets are partitioned by Peano-order slice (strong scaling): every rank sorts and
indexes all positions redundantly, sweeps its own slice and the moved (x, y, z, Hsml)
slices are re-assembled with one NCCL all-gather per step.

Printed line (rank 0): value = whole-job steps/s with inputs resident in HBM, device-timed
(CUDA events on the stream the kernels run on, max over ranks); e2e = the same steps driven
through the C ABI from pinned host buffers (upload + step + read-back inside the timed
region); roofline = the sweep kernel's algorithmic bytes / its event-timed duration against
the measured HBM copy bandwidth; cpu_baseline = the reference's own code (oracle/_ref) on the
host cores over a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULT_WORKLOAD = "merger_1e7"
STEP0 = 0.0085            # wvt_relax.c:51 (Mtotal >= 1e5 in every config)
CPU_SAMPLE_N = 200_000    # bounded CPU sample: same merger model at reduced N_gas
L2_BYTES = 126 << 20


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--n-gas", type=int, default=None, help="override particle count")
    ap.add_argument("--mode", default="fast", choices=["fast", "exact", "sequential"],
                    help="fast: TG_FAST (FP32 kernel arithmetic, 1e-5 distribution parity); exact: the "
                         "reference's mixed precision (rho/hsml bit-identical); sequential: "
                         "TG_WVT_SEQUENTIAL (everything bit-identical)")
    ap.add_argument("--sequential", action="store_true", help="same as --mode sequential")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--full-relaxation", action="store_true",
                    help="also time Regularise_sph_particles from the cold start to its own termination")
    return ap.parse_args()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic():
    """Per-launch DRAM bytes of the sweep kernel from the last committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "sweep_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------- reference arm

def run_reference(workload_name, n_target, steps, warmup, full_size):
    """The reference's own CPU code (oracle/_ref: the unmodified sources), all host threads.

    full_size (the --impl reference arm): the TRUE workload -- same N_gas as our arm.  A 10 M
    iteration of the reference costs ~40 s on 16 cores (the cold first one ~100 s), so the arm
    is bounded by the number of iterations, never by shrinking the problem: one cold + one warm
    iteration as warm-up, then min(steps, 2) timed steady-state iterations; the printed line
    says how many were timed.
    not full_size (the cpu_baseline leg inside our own arm, ~10-30 s of CPU work): the same
    merger model at N_gas = CPU_SAMPLE_N, steps/s scaled by N -- marked `scaled`."""
    from toycluster_b200 import workloads
    from oracle import ref
    if not ref.available():
        return None
    cores = os.cpu_count() or 1
    if full_size:
        n = n_target
        warmup = min(warmup, 2) if n > 2_000_000 else warmup
        steps = min(steps, 2) if n > 2_000_000 else (min(steps, 5) if n > 300_000 else steps)
    else:
        n = CPU_SAMPLE_N if n_target > CPU_SAMPLE_N else n_target
    n = max(n, 256 * cores)                       # wvt_relax.c:127 chunk must stay >= 1
    w = workloads.make(workload_name, n_gas=n)
    r = ref.Ref(w.n_gas, w.boxsize, w.mpart_gas, w.mtotal, w.halo_table(), cores)
    r.load(w.pos)
    stamps = []

    def cb(it):
        stamps.append(time.perf_counter())
        return 0

    total = warmup + steps
    r.regularise(total, cb)
    # stamps[k] = start of iteration k; iteration k spans stamps[k]..stamps[k+1]
    dt = np.diff(np.array(stamps))[warmup:warmup + steps]
    s_per_step = float(dt.mean())
    scaled = n != n_target
    return dict(n_sample=n, cores=r.nthreads, s_per_step_sample=s_per_step, steps=len(dt), warmup=warmup,
                scaled=scaled, steps_per_s=(n / n_target) / s_per_step,
                sample=(f"{workload_name} at N_gas={n}" + ("" if not scaled else f" (target {n_target})") +
                        f", {len(dt)} steady-state WVT iterations after {warmup} warm-up (the first one cold); "
                        "unmodified reference sources (oracle/_ref)" +
                        (f"; steps/s scaled by N_gas/{n_target} (cost per particle taken as constant)"
                         if scaled else "; full size, nothing scaled")))


# --------------------------------------------------------------------------- our arm

def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from toycluster_b200 import workloads
    n_gas = args.n_gas or workloads.CONFIGS[args.workload]["n_gas"]
    config = {"workload": args.workload, "n_gas": n_gas,
              "parallelism": f"targets partitioned over {world} rank(s), positions replicated",
              "l2": "inputs larger than L2" if n_gas * 16 > L2_BYTES else "L2 flushed between steps"}

    if args.impl == "reference":
        if rank != 0:
            return
        res = run_reference(args.workload, n_gas, args.steps, args.warmup, full_size=True)
        if res is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libtoyref.so not built"}))
            return
        v = res["steps_per_s"]
        config["n_gas_timed"] = res["n_sample"]
        config["scaled"] = res["scaled"]
        line = {"metric": "wvt_relax_steps_per_s", "value": v, "unit": "steps/s", "n_gpus": args.gpus,
                "steps": res["steps"], "warmup": res["warmup"], "steps_requested": args.steps,
                "warmup_requested": args.warmup, "ms_per_step": 1e3 / v,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32/f64",
                "data": "synthetic", "config": config, "impl": "reference",
                "cpu_baseline": {"value": v, "unit": "steps/s", "cores": res["cores"],
                                 "kind": "reference", "sample": res["sample"]},
                "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import toycluster_b200 as tc

    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator is created:
        # create it here with fd 1 pointed at stderr, so that stdout carries the JSON line only
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    stream = torch.cuda.Stream()            # the library launches on this stream, so the
    torch.cuda.set_stream(stream)           # events below see exactly its kernels

    w = workloads.make(args.workload, n_gas=n_gas)          # same seed on every rank
    mode = "sequential" if args.sequential else args.mode
    flags = {"fast": tc.FAST, "exact": 0, "sequential": tc.WVT_SEQUENTIAL}[mode]
    config["mode"] = mode
    g = tc.HotPath.from_workload(w, device=local_rank, flags=flags, rank=rank, nranks=world,
                                 stream=stream.cuda_stream)
    n = w.n_gas
    if world > 1:
        # the collectives of a step (all-gather of the moved slices, reduction of the error
        # statistics) run INSIDE libtoygpu.so on its own NCCL communicator; torch.distributed
        # only carries the 128-byte id to the other ranks and the timing reductions below
        ident = [tc.HotPath.comm_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            g.comm_init(ident[0])
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if n * 16 <= L2_BYTES else None

    def one_step(step):
        if flush is not None:
            flush.zero_()
        g.wvt_begin(step)              # sort, index, sweep, error statistics (global when multi-rank)
        g.wvt_finish(step)             # move + exchange of the moved slices: one pass of
        return g.stats()               # wvt_relax.c:66-214 exactly as tg_regularise runs it

    pos_np0 = w.pos.copy()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- resident-in-HBM timing -------------------------------------------------------
    g.upload(pos_np0)
    step = STEP0
    for _ in range(args.warmup):
        one_step(step)
    sync_all()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc = dict(pair_evals=0, gathered=0, kernels=0, sweep_ms=0.0, step_ms=0.0, searches=0, hsml_iters=0,
               handed_back=0, index_ms=0.0, tail_ms=0.0)
    t_wall = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        s = one_step(step)
        for k in acc:
            acc[k] += s[k]
    e1.record(stream)
    sync_all()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if sampler else None
    ms_total = e0.elapsed_time(e1)

    red = torch.tensor([ms_total, t_wall * 1e3, acc["sweep_ms"]], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(acc["pair_evals"]), float(acc["gathered"]), float(acc["searches"]),
                        float(acc["hsml_iters"]), float(acc["handed_back"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    per_rank = torch.tensor([acc["sweep_ms"] / args.steps, acc["step_ms"] / args.steps,
                             acc["index_ms"] / args.steps, acc["tail_ms"] / args.steps],
                            dtype=torch.float64, device="cuda")
    if world > 1:
        allr = [torch.empty_like(per_rank) for _ in range(world)]
        dist.all_gather(allr, per_rank)
        per_rank = torch.stack(allr)
    else:
        per_rank = per_rank[None]
    ms_total, wall_ms, sweep_ms_max = red.tolist()
    pair_evals, gathered, searches, hsml_iters, handed = tot.tolist()

    # ---- end to end through the operator boundary the C driver uses (main.c:52) ----------
    # Regularise_sph_particles() == tg_upload(AoS ParticleData 64 B + GasParticleData 60 B) ->
    # tg_regularise -> tg_download(AoS): the records live in page-locked HOST memory (as the shim
    # pins the driver's P / SphP), cross the bus whole in both directions inside the timed
    # region, and are unpacked / permuted / patched on the device.  One call runs `steps`
    # iterations, so the marshalling is amortised exactly as it is in the real program.
    e2e = None
    if not args.no_e2e:
        out = g.download()
        Pdt = np.dtype([("Pos", "3f4"), ("Vel", "3f4"), ("ID", "i4"), ("Type", "i4"),
                        ("Key", "2u8"), ("Tree_Parent", "i4"), ("pad", "3i4")])           # globals.h:161-168
        Sdt = np.dtype([("U", "f4"), ("Rho", "f4"), ("Hsml", "f4"), ("VarHsmlFac", "f4"),
                        ("Bfld", "3f4"), ("Apot", "3f4"), ("ID", "f4"), ("Rho_Model", "f4"),
                        ("Rs", "3f4")])                                                   # globals.h:170-180
        assert Pdt.itemsize == 64 and Sdt.itemsize == 60
        P = torch.zeros(n * 64, dtype=torch.uint8).pin_memory().numpy().view(Pdt)
        S = torch.zeros(n * 60, dtype=torch.uint8).pin_memory().numpy().view(Sdt)
        P["Pos"], P["ID"] = out["pos"], np.arange(n)
        S["Hsml"] = out["hsml"]                              # warm start: a mid-relaxation call
        stamps = {}
        for attempt in ("warm-up", "timed"):     # the first call allocates the record buffers
            S["Hsml"] = out["hsml"]              # (and, multi-rank, sets up NCCL for their size)
            P["Pos"] = out["pos"]
            sync_all()
            t0 = time.perf_counter()
            g.upload_records(P, S)                               # H2D: 124 B per particle
            t1 = time.perf_counter()
            done, _rows = g.regularise_sph_particles(max_iters=args.steps)
            t2 = time.perf_counter()
            g.download_records(P, S)                             # D2H: 124 B per particle, permuted
            sync_all()
            t3 = time.perf_counter()
            stamps = {"upload_ms": (t1 - t0) * 1e3, "regularise_ms": (t2 - t1) * 1e3,
                      "download_ms": (t3 - t2) * 1e3}
        dt = torch.tensor([t3 - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        final = g.download()
        checksum = float(final["pos"].astype(np.float64).sum()) + float(final["hsml"].astype(np.float64).sum())
        e2e = {"value": done / dt.item(), "unit": "steps/s", "state_checksum": checksum,
               "call": "tg_upload(AoS 64+60 B) -> tg_regularise -> tg_download(AoS), pinned host records",
               "iterations_per_call": done,
               "h2d_bytes_per_step": 124 * n / done, "d2h_bytes_per_step": 124 * n / done,
               "h2d_bytes_per_call": 124 * n, "d2h_bytes_per_call": 124 * n,
               "ms_per_step": dt.item() * 1e3 / done, "ms_per_call": dt.item() * 1e3, "rank0_ms": stamps}

    full = None
    if args.full_relaxation:
        g.upload(pos_np0)
        sync_all()
        t0 = time.perf_counter()
        iters, rows = g.regularise_sph_particles()
        sync_all()
        full = {"iterations": iters, "seconds": time.perf_counter() - t0,
                "final_err_mean": rows[-1]["mean"], "final_step": rows[-1]["step"]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = ms_total / args.steps
    value = 1e3 / ms_per_step
    peak, peak_kind = measured_peak()
    # algorithmic bytes of the sweep launch: 40 B per target (own x,y,z,h read,
    # Hsml/Rho/VarHsmlFac + displacement written) + 16 B per distinct gathered neighbour
    sweep_bytes = (40.0 * n + 16.0 * gathered / args.steps) / world
    sweep_ms = sweep_ms_max / args.steps
    achieved = sweep_bytes / (sweep_ms * 1e-3) / 1e9
    traffic = ncu_traffic()
    # (the sweep's events bracket the tile walk, the tile sweep and the generic sweep of the
    # hand-backs: the dominant kernel with its two satellites, 92 + 3 + 1 % of a step)
    kname = {"fast": "k_sweep_tile_fast<density+wvt> (+ k_tile_walk, hand-backs on k_sweep_fast)",
             "exact": "k_sweep_tile<density+wvt> (+ k_tile_walk, hand-backs on k_sweep)",
             "sequential": "k_sweep_tile<density> + k_sweep<wvt_seq> (+ k_tile_walk, hand-backs)"}[mode]
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak,
                "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic["bytes_per_launch"] if traffic else None,
                "sweep_ms": sweep_ms, "sweep_share_of_step": sweep_ms / ms_per_step,
                "step_bytes_model": 300.0 * n + 16.0 * gathered / args.steps,
                "step_frac": (300.0 * n + 16.0 * gathered / args.steps) / world
                             / (ms_per_step * 1e-3) / 1e9 / peak}

    line = {"metric": "wvt_relax_steps_per_s", "value": value, "unit": "steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": ("f32 (TG_FAST: exact f32 neighbour predicate, packed-f32 kernels, f64 reduction across lanes)"
                      if mode == "fast" else "f32 predicate / f64 sums (reference's mixed precision)"),
            "data": "synthetic", "config": config,
            "interactions_per_s": pair_evals / args.steps * value,
            "pair_evals_per_particle": pair_evals / args.steps / n,
            "gathered_per_particle": gathered / args.steps / n,
            "wall_ms_per_step": wall_ms / args.steps,
            "searches_per_particle": searches / args.steps / n,
            "hsml_iters_per_particle": hsml_iters / args.steps / n,
            "handed_back_per_step": handed / args.steps,
            "gpu_launches": int(acc["kernels"]), "clocks": clocks, "roofline": roofline}
    # per-rank device times: load balance of the target partition (sweep), the part of a step every
    # rank repeats for all n (index: keys, sort, model pass, box hierarchy, displaced nodes) and
    # what follows the sweep (tail: error sums, move, exchange of the moved slices)
    line["per_rank_ms"] = {k: [round(v, 3) for v in per_rank[:, j].tolist()]
                           for j, k in enumerate(("sweep", "step", "index", "tail"))}
    if e2e:
        line["e2e"] = e2e
    if full:
        line["full_relaxation"] = full
    if not args.no_cpu_baseline and world == 1:
        res = run_reference(args.workload, n, 2, 1, full_size=False)
        if res:
            line["cpu_baseline"] = {"value": res["steps_per_s"], "unit": "steps/s",
                                    "cores": res["cores"], "kind": "reference",
                                    "sample": res["sample"], "scaled": res["scaled"],
                                    "n_sample": res["n_sample"],
                                    "s_per_step_sample": res["s_per_step_sample"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
