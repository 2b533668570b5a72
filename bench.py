#!/usr/bin/env python
"""bench.py — ICP registrations/s on 120k-point scan pairs (BASELINE.json metric).

Workload (config.workload): BASELINE.json configs[1] — synthetic 64-beam
120 000-point scan pair, point-to-plane ICP with the SurfaceNormal(knn=10)
reference filter, KDTreeMatcher k=1 eps=0, TrimmedDist 0.85, Counter(40) +
Differential checkers.  One "step" = one pass of the whole registration path
(reference filter -> index build -> ICP loop -> result) over one batch of
`--pairs` distinct pairs per GPU; pairs are independent, so N GPUs each take
their own batch (weak scaling) and only the per-pair 4x4 results are gathered
(NCCL all_gather), inside the timed region.

  value : registrations/s with the clouds already resident in HBM
  e2e   : the same through the public API from pinned HOST buffers (H2D of
          both clouds and D2H of the result inside the timed region)
  roofline : the dominant kernel (match = fused transform + exact kNN), against
          the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline : the CPU oracle (a port of the libpointmatcher/libnabo path) on
          this box's host cores, same workload, bounded sample

`--impl reference` times that CPU path alone (the reference's libpointmatcher
cannot be built here: un-vendored, un-pinned, absent — DESIGN.md).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "icp_registrations_per_s_120k_pt_pairs"
UNIT = "registrations/s"
WORKLOAD = ("C2: synthetic 64-beam 120000-pt scan pair, point-to-plane ICP, SurfaceNormal(knn=10) reference "
            "filter, KDTreeMatcher k=1 eps=0, TrimmedDist 0.85, Counter(40)+Differential(1e-3,1e-3,3)")


def c2_config():
    from tests import util
    return util.C2


def gen_pair(seed):
    from pgslam_b200 import synth
    rd, rf, _ = synth.scan_pair(seed, beams=64, az_steps=1875)
    return rd, rf


def gen_pairs(seeds):
    """Distinct synthetic pairs; generated in parallel on the host cores."""
    seeds = list(seeds)
    if len(seeds) <= 2:
        return [gen_pair(s) for s in seeds]
    import multiprocessing as mp
    with mp.get_context("fork").Pool(min(len(seeds), os.cpu_count() or 1)) as pool:
        return pool.map(gen_pair, seeds)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_run(steps, warmup, sample_pairs=1):
    """The CPU path (oracle port) on all host threads: one registration per step."""
    from oracle import binding as ob
    cfg = ob.config_from_dict(c2_config())
    pairs = gen_pairs(range(1000, 1000 + sample_pairs))
    clouds = [(ob.Cloud(rd), ob.Cloud(rf)) for rd, rf in pairs]
    # all the host threads the box has (torchrun exports OMP_NUM_THREADS=1 to its workers)
    ob.lib().orc_set_num_threads(os.cpu_count() or 1)
    threads = ob.lib().orc_num_threads()
    its = []
    for i in range(warmup):
        ob.icp_run(cfg, *clouds[i % len(clouds)])
    t0 = time.perf_counter()
    for i in range(steps):
        r = ob.icp_run(cfg, *clouds[i % len(clouds)])
        its.append(r["iterations"])
    dt = time.perf_counter() - t0
    return dict(value=steps / dt, seconds=dt, cores=threads, iterations=its,
                sample=f"{steps} registration(s) of the same C2 workload ({len(clouds)} distinct pair(s)), "
                       f"OpenMP over queries on {threads} host thread(s)")


def emit(line: dict, fd: int):
    os.write(fd, (json.dumps(line) + "\n").encode())


def main():
    # exactly ONE line may reach stdout: native libraries (NCCL prints its version)
    # write to fd 1 behind Python's back, so fd 1 is pointed at stderr for the
    # whole run and the JSON line goes to the saved descriptor
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=96, help="distinct pairs per GPU per step")
    ap.add_argument("--cpu-steps", type=int, default=0, help="registrations for the cpu_baseline leg (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch-streams", type=int, default=4,
                    help="worker streams the library splits a batch over (pgs_ctx_set_batch_streams)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # ------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(max(args.steps, 1), args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / max(args.steps, 1),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 points / f64 reductions",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "pairs_per_gpu_per_step": args.pairs, "points_per_scan": 120000},
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "iterations": r["iterations"]}
        emit(line, out_fd)
        return

    # --------------------------------------------------------------------- ours
    import torch
    import torch.distributed as dist
    from pgslam_b200 import build, pm
    from tests import util

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    build.build()
    torch.cuda.set_device(local_rank)
    multi = world > 1
    if multi:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.current_stream()
    ctx = pm.Context(local_rank, stream.cuda_stream)
    ctx.set_batch_streams(args.batch_streams)
    icp = pm.ICP(ctx)
    icp.loadFromYaml(util.to_yaml(c2_config()))

    B = args.pairs
    pairs = gen_pairs(range(rank * B, rank * B + B))
    n_pts = pairs[0][0].shape[1]
    in_bytes = sum(rd.nbytes + rf.nbytes for rd, rf in pairs)
    # pinned host copies for the e2e leg (point-major float32, exactly what the ABI takes)
    host = [(torch.from_numpy(np.ascontiguousarray(rd.T)).pin_memory(), torch.from_numpy(np.ascontiguousarray(rf.T)).pin_memory())
            for rd, rf in pairs]
    dev_rd = [pm.DataPoints(ctx=ctx, device_ptr=None, features=rd) for rd, _ in pairs]
    dev_rf = [pm.DataPoints(ctx=ctx, device_ptr=None, features=rf) for _, rf in pairs]

    gathered = torch.empty((world * B, 16), dtype=torch.float64, device="cuda") if multi else None

    def gather(results):
        """the path's only exchange step: per-pair transforms to every rank"""
        if not multi:
            return
        loc = torch.tensor(np.stack([r["T"].ravel(order="F") for r in results]), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(gathered, loc)

    resident_wall = []  # host wall-clock of every device-resident step (warm-up included)

    def step_resident():
        t0 = time.perf_counter()
        res = icp.compute_batch(dev_rd, dev_rf)
        gather(res)
        resident_wall.append(round(1e3 * (time.perf_counter() - t0), 2))
        return res

    def upload():
        """H2D of one step's inputs: 2 x B clouds from pinned host memory, asynchronous on the
        library's copy stream (public API: DataPoints(pinned_host_ptr=...))."""
        rds = [pm.DataPoints(ctx=ctx, pinned_host_ptr=hrd.data_ptr(), n=hrd.shape[0]) for hrd, _ in host]
        rfs = [pm.DataPoints(ctx=ctx, pinned_host_ptr=hrf.data_ptr(), n=hrf.shape[0]) for _, hrf in host]
        return rds, rfs

    pending = []
    e2e_wall = []  # host wall-clock of every e2e step (warm-up included), for the record
    from concurrent.futures import ThreadPoolExecutor
    uploader = ThreadPoolExecutor(max_workers=1)

    def step_e2e():
        # software pipeline: this step's inputs were queued on the copy stream while the previous
        # step computed, and the next step's are queued by a helper thread while this one computes
        # (pinned uploads touch only the library's copy stream); every step still uploads its own
        # inputs and downloads its own results
        t0 = time.perf_counter()
        rds, rfs = pending.pop().result() if pending else upload()
        t1 = time.perf_counter()
        pending.append(uploader.submit(upload))
        res = icp.compute_batch(rds, rfs)
        gather(res)
        e2e_wall.append(round(1e3 * (time.perf_counter() - t0), 2))
        if os.environ.get("BENCH_TRACE_E2E"):
            sys.stderr.write("e2e step: waited %.1f ms for the upload calls, compute_batch %.1f ms\n"
                             % (1e3 * (t1 - t0), 1e3 * (time.perf_counter() - t1)))
        return res

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = None
        for _ in range(steps):
            out = fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if multi:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    # a generation-2 Python garbage collection in a process that has imported torch takes ~35 ms
    # (measured: one e2e step in ~15 doubled); the steps allocate nothing cyclic, so the
    # collector is parked for the measurement instead of being timed
    import gc
    gc.collect()
    gc.freeze()
    gc.disable()

    # clocks are sampled from the warm-up steps on (same load as the timed steps): the
    # timed region alone is ~100 ms, too short for more than a sample or two
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("BENCH_NO_SAMPLER"):
        sampler.start()
    for _ in range(args.warmup):
        step_resident()
    launches0 = ctx.launch_count
    ms, res = timed(step_resident, args.steps)
    launches = ctx.launch_count - launches0
    # per-kernel CUDA-event timings of one more identical step, taken with the batch on ONE
    # stream: with the sub-batches overlapping on several streams an event pair around a kernel
    # would also time whatever the other streams run in between
    ctx.set_profiling(True)
    step_resident()
    stage = ctx.stage_times()
    ctx.set_profiling(False)
    clocks = sampler.stop() if rank == 0 else None

    for _ in range(args.warmup):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)

    if pending:
        pending.pop().result()  # drain the pipeline before the latency measurement
    uploader.shutdown()

    # single-pair latency (batch of 1), for the record
    one_rd, one_rf = [dev_rd[0]], [dev_rf[0]]
    for _ in range(3):
        icp.compute_batch(one_rd, one_rf)
    ms_one, _ = timed(lambda: icp.compute_batch(one_rd, one_rf), 10)

    if rank != 0:
        if multi:
            dist.destroy_process_group()
        return

    total_pairs = world * B
    value = total_pairs * args.steps / (ms / 1e3)
    e2e_value = total_pairs * args.steps / (ms_e2e / 1e3)
    iters = [r["iterations"] for r in res]
    ok = sum(r["status"] == 0 for r in res)
    # roofline of the dominant kernel (match): algorithmic bytes = per query 16 B
    # read + 8 B match written, plus one pass over the reference (16 B/pt) per
    # launch and pair; only pairs still iterating do work.
    n_ref = res[0]["n_reference"]
    alg_bytes = sum(r["iterations"] * (24.0 * r["n_reading"] + 16.0 * r["n_reference"]) for r in res)
    match_s = stage["match_ms"] / 1e3
    peak, peak_src = peaks()
    achieved = alg_bytes / match_s / 1e9 if match_s > 0 else 0.0
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture,
    # valid only for the batch size it was captured at
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r1_match_kernel_ncu_full.json")) as f:
            cap = json.load(f)
        if cap.get("pairs") == B:
            traffic = cap["dram_traffic_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "match_kernel (fused rigid transform + exact k=1 NN traversal)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "algorithmic_bytes_per_launch_all_pairs_active": sum(24.0 * r["n_reading"] + 16.0 * r["n_reference"] for r in res),
                "peak_source": peak_src, "launches": stage["iterations_launched"],
                "avg_launch_ms": stage["match_ms"] / max(stage["iterations_launched"], 1),
                "algorithmic_bytes_per_step": alg_bytes,
                "timing": "CUDA events around every launch of one extra step after the timed region, whole batch on one stream",
                "stage_ms_last_step": {k: round(float(v), 3) for k, v in stage.items()}}
    # the metric's second figure: exact nearest-neighbour queries answered per second by the
    # matcher kernel alone (every launch answers one query per reading point of an active pair)
    knn_qps = sum(r["iterations"] * r["n_reading"] for r in res) / match_s if match_s > 0 else 0.0
    reg_bytes = 68.0 * n_ref + 32.0 * n_pts + statistics.mean(iters) * ((32 + 32 * 0.85) * n_pts + 16.0 * n_ref)
    roofline["whole_registration"] = {"algorithmic_bytes": reg_bytes, "achieved_gbs": reg_bytes * value / world / 1e9,
                                      "frac": reg_bytes * value / world / 1e9 / peak}

    cpu = None
    if not args.no_cpu_baseline and world == 1:  # reported on rank 0 at N = 1 only
        n = args.cpu_steps or 64  # ~10 s of CPU work on this class of host
        r = cpu_reference_run(n, 1, sample_pairs=min(4, n))
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 points / f64 reductions", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu_per_step": B, "points_per_scan": n_pts,
                       "batch_streams": args.batch_streams,
                       "l2_policy": f"inputs larger than L2: {in_bytes / 1e6:.0f} MB of distinct clouds per GPU per step",
                       "parallelism": f"{world} GPU(s), independent pairs per rank, NCCL all_gather of 4x4 results"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes * world,
                    "d2h_bytes_per_step": int(total_pairs * 480), "ms_per_step": ms_e2e / args.steps,
                    "wall_ms_each_step_incl_warmup": e2e_wall},
            "resident_wall_ms_each_step_incl_warmup": resident_wall[:args.warmup + args.steps],
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "knn_queries_per_s": knn_qps, "iterations_mean": statistics.mean(iters), "pairs_ok": ok,
            "single_pair_latency_ms": ms_one / 10.0}
    emit(line, out_fd)
    if multi:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
