#!/usr/bin/env python
"""bench.py — ICP registrations/s on 120k-point scan pairs (BASELINE.json metric).

Workload (config.workload): every registration is BASELINE.json configs[1] — synthetic 64-beam
120 000-point scan pair, point-to-plane ICP with the SurfaceNormal(knn=10) reference filter,
KDTreeMatcher k=1 eps=0, TrimmedDist 0.85, Counter(40) + Differential checkers — and the batch
is configs[3]: a fixed pool of `--pool` (4096) DISTINCT candidate pairs, cut into contiguous
blocks over the N GPUs (pgslam_b200/dist.py shard_range; strong scaling, no data-path
collective).  One "step" = every rank registers its whole block (reference filter -> index
build -> ICP loop -> result) and the per-pair result records are gathered to every rank
(NCCL all_gather), inside the timed region.

  value : registrations/s with the clouds already resident in HBM (pgs_icp_run_batch)
  e2e   : the same through the host-memory entry pgs_icp_run_batch_multi from PINNED HOST buffers
          (H2D of both clouds of every pair and D2H of the results inside the timed region)
  roofline : the dominant kernel (match = fused transform + exact kNN), against the measured HBM
          copy bandwidth in MEASURED_PEAKS.json
  bench_parity : >= 16 pairs of the pool re-registered by the CPU oracle inside the run
          (iterations equal, pose deltas); results_sha256 = hash of all gathered records, equal at
          every N because a pair's result does not depend on how the pool is cut
  cpu_baseline : the CPU oracle (a port of the libpointmatcher/libnabo path) on this box's host
          cores (all threads, and one thread), same workload, bounded sample
  c3 / c5 / dropin_ms : the other BASELINE configs and the C++ drop-in call sites (N = 1 only)

`--impl reference` times that CPU path alone (the reference's libpointmatcher cannot be built
here: un-vendored, un-pinned, absent — DESIGN.md).
"""
from __future__ import annotations

import argparse
import hashlib
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
WORKLOAD = ("C2 registrations drawn from the C4 pool: synthetic 64-beam 120000-pt scan pairs, point-to-plane ICP, "
            "SurfaceNormal(knn=10) reference filter, KDTreeMatcher k=1 eps=0, TrimmedDist 0.85, "
            "Counter(40)+Differential(1e-3,1e-3,3); fixed pool of distinct candidate pairs sharded over the GPUs")
BEAMS, AZ = 64, 1875


def c2_config():
    from tests import util
    return util.C2


def gen_pair(seed):
    from pgslam_b200 import synth
    rd, rf, _ = synth.scan_pair(seed, beams=BEAMS, az_steps=AZ)
    return rd, rf


def gen_pairs(seeds):
    """Distinct synthetic pairs from the numpy generator; in parallel on the host cores."""
    seeds = list(seeds)
    if len(seeds) <= 2:
        return [gen_pair(s) for s in seeds]
    import multiprocessing as mp
    with mp.get_context("fork").Pool(min(len(seeds), os.cpu_count() or 1)) as pool:
        return pool.map(gen_pair, seeds)


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region, every 100 ms: through NVML in
    this process (a helper thread; the timed calls release the GIL), or with an `nvidia-smi -lms 100` child
    when pynvml is missing.  (The child costs the host-entry leg about 2 %: 4 442 vs 4 533 registrations/s
    with and without it; the in-process queries touch the driver far less.)"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.nvml = None
        self.handle = None
        self.source = None
        self._stop = threading.Event()

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:  # the CUDA ordinal need not be NVML's index: go by UUID
            import torch
            uuid = "GPU-" + str(torch.cuda.get_device_properties(self.index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if isinstance(uuid, str) else uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        return pynvml, h

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self._sample_nvml()  # fails here, not in the thread, when a query is unsupported
            self.rows = []
            self.source = "nvml"
            self.t = threading.Thread(target=self._loop_nvml, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _sample_nvml(self):
        n, h = self.nvml, self.handle
        sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
        pw = n.nvmlDeviceGetPowerUsage(h) / 1000.0
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        bits = [n.nvmlClocksThrottleReasonHwSlowdown, n.nvmlClocksThrottleReasonHwThermalSlowdown,
                n.nvmlClocksThrottleReasonSwThermalSlowdown, n.nvmlClocksThrottleReasonSwPowerCap]
        self.rows.append((float(sm), float(mx), float(pw), [bool(r & b) for b in bits]))

    def _loop_nvml(self):
        while not self._stop.is_set():
            try:
                self._sample_nvml()
            except Exception:
                pass
            self._stop.wait(0.1)

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) < 7:
                continue
            try:
                self.rows.append((float(f[0]), float(f[1]), float(f[2]), [v.lower().startswith("active") for v in f[3:7]]))
            except ValueError:
                continue

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.t.join(timeout=2)
        elif self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML and no nvidia-smi"]}
        rows = list(self.rows)
        sm = [r[0] for r in rows]
        reasons = sorted({nme for r in rows for nme, on in zip(self.NAMES, r[3]) if on})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(r[1] for r in rows) if rows else None,
                "power_w_max": max(r[2] for r in rows) if rows else None, "samples": len(rows), "reasons": reasons,
                "source": self.source}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_run(steps, warmup, sample_pairs=1, threads=None):
    """The CPU path (oracle port): one registration per step, OpenMP over queries."""
    from oracle import binding as ob
    cfg = ob.config_from_dict(c2_config())
    pairs = gen_pairs(range(1000, 1000 + sample_pairs))
    clouds = [(ob.Cloud(rd), ob.Cloud(rf)) for rd, rf in pairs]
    # all the host threads the box has (torchrun exports OMP_NUM_THREADS=1 to its workers)
    ob.lib().orc_set_num_threads(threads or os.cpu_count() or 1)
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
                       f"OpenMP over queries on {threads} host thread(s); about 0.14 s of each registration is "
                       f"serial (copies, quantile, minimizer)")


def resolve_batching(args, world):
    if args.batch_streams <= 0:
        # every worker stream has a host thread (mostly asleep in its waits); 8 per rank measured best at
        # N = 1, 4 per rank is what the 16-core box carried at N = 8 in round 1
        args.batch_streams = 8 if world * 8 <= 2 * (os.cpu_count() or 16) else 4
    if args.batch_chunk <= 0:
        args.batch_chunk = 96 // args.batch_streams


def make_config(args, world):
    """the `config` object of the JSON line: identical for the GPU arm and the reference arm"""
    per = -(-args.pool // world)
    in_bytes = 2 * per * BEAMS * AZ * 16
    return {"workload": WORKLOAD, "pool_pairs": args.pool, "pairs_per_gpu_per_step": per,
            "points_per_scan": BEAMS * AZ, "batch_streams": args.batch_streams, "batch_chunk": args.batch_chunk,
            "l2_policy": f"inputs larger than L2: {in_bytes / 1e6:.0f} MB of distinct clouds per GPU per step",
            "parallelism": f"{world} GPU(s), contiguous blocks of the pool per rank, NCCL all_gather of the result records",
            "pool_generator": "pgslam_b200/synth_torch.py on the device, pair i <- seed i"}


def emit(line: dict, fd: int):
    os.write(fd, (json.dumps(line) + "\n").encode())


# ------------------------------------------------------------------------------------------------
def leg_c3(pm, ctx, util, dev, scans=2000):
    """BASELINE C3: a 2000-scan odometry sequence through ICPSequence (Localizer.hpp:103-126,148,254):
    every scan is uploaded from pinned host memory, run through the input filters, and registered
    against a local map of up to 3 keyframes seeded with the previous pose; every 10th scan becomes
    a keyframe (local map rebuilt on the device, setMap).  One sequence is sequential: one GPU."""
    import torch
    from pgslam_b200 import synth, synth_torch
    poses = synth.trajectory(scans + 1, step=0.25, turn_deg=1.0)
    dev_scans = synth_torch.trajectory_scans(77, poses, dev, beams=BEAMS, az_steps=AZ)
    n_pts = BEAMS * AZ
    host = torch.empty((scans + 1, n_pts, 4), dtype=torch.float32).pin_memory()
    for i, t in enumerate(dev_scans):
        host[i].copy_(t, non_blocking=True)
    torch.cuda.synchronize()
    del dev_scans
    filt = pm.DataPointsFilters(util.to_yaml(util.INPUT_FILTERS), ctx=ctx)
    seq = pm.ICPSequence(ctx)
    seq.loadFromYaml(util.to_yaml(util.C2))

    def cloud(i):
        c = pm.DataPoints(ctx=ctx, pinned_host_ptr=host[i].data_ptr(), n=n_pts)
        filt.apply(c)
        return c
    keyframes = [(cloud(0), np.eye(4))]
    seq.setMap(keyframes[0][0])
    T = np.eye(4)
    lat, its, fails = [], [], 0
    t_all = time.perf_counter()
    for i in range(1, scans + 1):
        t0 = time.perf_counter()
        c = cloud(i)
        try:
            T = seq(c, T)
        except pm.PointMatcherError:
            fails += 1
        its.append(seq.last["iterations"])
        if i % 10 == 0:  # keyframe: local map = last 3 keyframes in the newest one's frame
            keyframes.append((c, T.copy()))
            keyframes = keyframes[-3:]
            Tn = np.linalg.inv(keyframes[-1][1])
            m = pm.assemble_local_map([keyframes[-1][0]] + [k for k, _ in keyframes[:-1]],
                                      [np.eye(4)] + [Tn @ Tk for _, Tk in keyframes[:-1]])
            seq.setMap(m)
            keyframes = [(k, Tn @ Tk) for k, Tk in keyframes]
            T = np.eye(4)
        lat.append(1e3 * (time.perf_counter() - t0))
    dt = time.perf_counter() - t_all
    lat.sort()
    return {"scans": scans, "scans_per_s": scans / dt, "latency_ms_p50": lat[len(lat) // 2],
            "latency_ms_p99": lat[min(len(lat) - 1, int(0.99 * len(lat)))], "iterations_mean": statistics.mean(its),
            "failed_registrations": fails,
            "note": "distinct 120k-pt scans along a trajectory, each uploaded from pinned host memory, input filters "
                    "(SurfaceNormal knn=10 + 3 more), ICPSequence against a local map of up to 3 keyframes (360k pts), "
                    "keyframe + local-map rebuild + setMap every 10 scans"}


def leg_c5(pm, ctx, util, reps=5):
    """BASELINE C5: dense 1M-point scan-to-map registration, voxel-subsampled reading, trimmed 0.75."""
    from pgslam_b200 import synth
    rd, rf, _ = synth.scan_pair(9, beams=128, az_steps=7813)  # 1 000 064 points
    icp = pm.ICP(ctx)
    icp.loadFromYaml(util.to_yaml(util.C5))
    a, b = pm.DataPoints(rd, ctx=ctx), pm.DataPoints(rf, ctx=ctx)
    icp(a, b)
    t0 = time.perf_counter()
    for _ in range(reps):
        icp(a, b)
    dt = (time.perf_counter() - t0) / reps
    return {"points": int(rd.shape[1]), "ms_per_registration": 1e3 * dt, "registrations_per_s": 1.0 / dt,
            "iterations": icp.last["iterations"], "n_reading_after_voxel_filter": icp.last["n_reading"]}


def leg_dropin():
    """pgslam's own call sites (LoopCloser::ProcessVertex + CheckIcpResult + ComputeResidualError,
    Localizer::ComputeOverlapWith) compiled against the C++ adapter, timed end to end at 120k points."""
    try:
        from tests.test_cpp_adapter import build_callsites
        exe = build_callsites()
        out = subprocess.run([exe, "--bench", "120000"], capture_output=True, text=True, timeout=300)
        for ln in out.stdout.splitlines():
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": (out.stdout + out.stderr)[-300:]}
    except Exception as e:  # the adapter bench is an extra: never fail the bench line for it
        return {"error": repr(e)[:300]}


# ------------------------------------------------------------------------------------------------
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
    ap.add_argument("--pool", type=int, default=4096, help="distinct candidate pairs in the job (BASELINE C4)")
    ap.add_argument("--cpu-steps", type=int, default=0, help="registrations for the cpu_baseline leg (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the C3 / C5 / drop-in legs")
    ap.add_argument("--parity-pairs", type=int, default=16)
    ap.add_argument("--batch-streams", type=int, default=0,
                    help="worker streams the library splits a batch over (0 = 8, or 4 when the ranks of the job "
                         "would otherwise outnumber the host cores)")
    ap.add_argument("--batch-chunk", type=int, default=0, help="pairs per chunk a worker pulls (0 = 96 / streams)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    resolve_batching(args, args.gpus if args.impl == "reference" else world)

    # ------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(max(args.steps, 1), args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / max(args.steps, 1),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 points / f64 reductions",
                "data": "synthetic",
                "config": make_config(args, max(args.gpus, 1)),
                "reference_step": "one registration of the workload per step (bounded sample of the pool), all host threads",
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "iterations": r["iterations"]}
        emit(line, out_fd)
        return

    # --------------------------------------------------------------------- ours
    import torch
    import torch.distributed as dist
    from pgslam_b200 import build, pm, synth_torch
    from pgslam_b200 import dist as pdist
    from tests import util

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    build.build()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    multi = world > 1
    if multi:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream()
    ctx = pm.Context(local_rank, stream.cuda_stream)
    ctx.set_batch_streams(args.batch_streams)
    ctx.set_option("batch_chunk", args.batch_chunk)
    icp = pm.ICP(ctx)
    icp.loadFromYaml(util.to_yaml(c2_config()))

    # ---- the pool: this rank's contiguous block, generated on the device, pair i <- seed i ----------
    POOL = args.pool
    mine = pdist.shard_range(POOL, rank, world)
    B = len(mine)
    n_pts = BEAMS * AZ
    t_setup = time.perf_counter()
    host = torch.empty((max(B, 1), 2, n_pts, 4), dtype=torch.float32).pin_memory()  # e2e leg: pinned host copies
    dev_rd, dev_rf = [], []
    for c0 in range(0, B, 32):
        seeds = [mine[j] for j in range(c0, min(B, c0 + 32))]
        data, _ = synth_torch.scan_pairs(seeds, dev, beams=BEAMS, az_steps=AZ)
        # the library copies on ITS stream (non-blocking: no implicit ordering with torch's): the generator's
        # kernels must have finished before the clouds are handed over
        torch.cuda.synchronize()
        for j, (rd, rf) in enumerate(data):
            dev_rd.append(pm.DataPoints(ctx=ctx, device_ptr=rd.data_ptr(), n=n_pts))
            dev_rf.append(pm.DataPoints(ctx=ctx, device_ptr=rf.data_ptr(), n=n_pts))
            host[c0 + j, 0].copy_(rd, non_blocking=True)
            host[c0 + j, 1].copy_(rf, non_blocking=True)
        ctx.synchronize()
        torch.cuda.synchronize()
        del data
    torch.cuda.empty_cache()
    handles = pm.batch_handles(dev_rd, dev_rf)
    host_rd = pm.host_clouds([(host[j, 0].data_ptr(), n_pts) for j in range(B)])
    host_rf = pm.host_clouds([(host[j, 1].data_ptr(), n_pts) for j in range(B)])
    in_bytes = 2 * B * n_pts * 16
    setup_s = time.perf_counter() - t_setup

    def gather(rec):
        """the path's only exchange step: per-pair result records to every rank (pair order)"""
        return pdist.gather_records(rec, POOL, dev if multi else None)

    resident_wall, e2e_wall = [], []

    def step_resident():
        t0 = time.perf_counter()
        rec = icp.compute_batch_array(dev_rd, dev_rf, handles=handles) if B else pm.empty_records()
        allrec = gather(rec)
        resident_wall.append(round(1e3 * (time.perf_counter() - t0), 2))
        return rec, allrec

    def step_e2e():
        t0 = time.perf_counter()
        rec = pm.compute_batch_multi([icp], host_rd, host_rf, pinned=True) if B else pm.empty_records()
        allrec = gather(rec)
        e2e_wall.append(round(1e3 * (time.perf_counter() - t0), 2))
        return rec, allrec

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

    # a generation-2 Python garbage collection in a process that has imported torch takes ~35 ms;
    # the steps allocate nothing cyclic, so the collector is parked for the measurement
    import gc
    gc.collect()
    gc.freeze()
    gc.disable()

    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("BENCH_NO_SAMPLER"):
        sampler.start()
    for _ in range(args.warmup):
        step_resident()
    launches0 = ctx.launch_count
    ms, (rec, allrec) = timed(step_resident, args.steps)
    launches = ctx.launch_count - launches0

    for _ in range(args.warmup):
        step_e2e()
    ms_e2e, (rec_e2e, allrec_e2e) = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # per-kernel CUDA-event timings of one 96-pair batch, taken on ONE stream: with chunks
    # overlapping on several streams an event pair around a kernel would also time whatever
    # the other streams run in between
    nprof = min(96, B)
    stage = None
    prof_rec = None
    if nprof:
        ctx.set_profiling(True)
        prof_rec = icp.compute_batch_array(dev_rd[:nprof], dev_rf[:nprof])
        stage = ctx.stage_times()
        ctx.set_profiling(False)
        # single-pair latency (batch of 1), for the record
        for _ in range(3):
            icp.compute_batch_array(dev_rd[:1], dev_rf[:1])
        ms_one, _ = timed(lambda: icp.compute_batch_array(dev_rd[:1], dev_rf[:1]), 10)
    else:
        ms_one = 0.0
        timed(lambda: None, 1)

    if rank != 0:
        if multi:
            dist.destroy_process_group()
        return

    value = POOL * args.steps / (ms / 1e3)
    e2e_value = POOL * args.steps / (ms_e2e / 1e3)
    ok = int((allrec[:, 53] == 0).sum())
    iters = allrec[:, 52]
    sha = hashlib.sha256(np.ascontiguousarray(allrec).tobytes()).hexdigest()

    # ---- in-run parity: pairs of this rank's block re-registered by the CPU oracle -------------------
    parity = None
    if args.parity_pairs > 0 and B:
        from oracle import binding as ob
        ob.lib().orc_set_num_threads(os.cpu_count() or 1)
        cfg = ob.config_from_dict(c2_config())
        pick = sorted(set(int(x) for x in np.linspace(0, B - 1, min(args.parity_pairs, B))))
        max_dt = max_dr = 0.0
        it_equal = 0
        worst = None
        for j in pick:
            rd = np.asfortranarray(host[j, 0].numpy().T)
            rf = np.asfortranarray(host[j, 1].numpy().T)
            want = ob.icp_run(cfg, ob.Cloud(rd), ob.Cloud(rf))
            got_T = rec[j]["T"].reshape(4, 4).T
            dt = float(np.abs(got_T[:3, 3] - want["T"][:3, 3]).max())
            dr = util.rot_angle(got_T, want["T"])
            it_equal += int(int(rec[j]["iterations"]) == int(want["iterations"]))
            if dt >= max_dt:
                worst = int(mine[j])
            max_dt, max_dr = max(max_dt, dt), max(max_dr, dr)
        parity = {"pairs_checked": len(pick), "pool_indices": [int(mine[j]) for j in pick], "iterations_equal": it_equal,
                  "max_dT_m": max_dt, "max_dR_rad": max_dr, "worst_pair": worst, "tolerance": "1e-5 m / 1e-5 rad, equal iteration counts",
                  "ok": bool(it_equal == len(pick) and max_dt <= 1e-5 and max_dr <= 1e-5),
                  "checker": "oracle/ (CPU port), same arrays the GPU registered"}

    # ---- roofline of the dominant kernel (match): algorithmic bytes = per query 16 B read + 8 B match
    # written, plus one pass over the reference (16 B/pt) per launch and pair
    peak, peak_src = peaks()
    roofline = None
    if stage:
        alg_bytes = float(sum(int(r["iterations"]) * (24.0 * int(r["n_reading"]) + 16.0 * int(r["n_reference"])) for r in prof_rec))
        per_launch_all = float(sum(24.0 * int(r["n_reading"]) + 16.0 * int(r["n_reference"]) for r in prof_rec))
        match_s = stage["match_ms"] / 1e3
        achieved = alg_bytes / match_s / 1e9 if match_s > 0 else 0.0
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r2_match_kernel_ncu_full.json")) as f:
                cap = json.load(f)
            if cap.get("pairs") == nprof:
                traffic = cap["dram_traffic_bytes_per_launch"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": "match_kernel (fused rigid transform + exact k=1 NN traversal)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "algorithmic_bytes_per_launch_all_pairs_active": per_launch_all,
                    "peak_source": peak_src, "launches": stage["iterations_launched"], "pairs_in_profiled_batch": nprof,
                    "avg_launch_ms": stage["match_ms"] / max(stage["iterations_launched"], 1),
                    "algorithmic_bytes_per_step": alg_bytes,
                    "timing": "CUDA events around every launch of one 96-pair batch after the timed region, whole batch on one stream",
                    "stage_ms_profiled_batch": {k: round(float(v), 3) for k, v in stage.items()}}
        n_ref = int(prof_rec[0]["n_reference"])
        reg_bytes = 68.0 * n_ref + 32.0 * n_pts + float(iters.mean()) * ((32 + 32 * 0.85) * n_pts + 16.0 * n_ref)
        roofline["whole_registration"] = {"algorithmic_bytes": reg_bytes, "achieved_gbs": reg_bytes * value / world / 1e9,
                                          "frac": reg_bytes * value / world / 1e9 / peak}
        knn_qps = float(sum(int(r["iterations"]) * int(r["n_reading"]) for r in prof_rec)) / match_s if match_s > 0 else 0.0
    else:
        knn_qps = 0.0

    cpu = None
    extras = {}
    if world == 1:
        if not args.no_cpu_baseline:  # reported on rank 0 at N = 1 only
            n = args.cpu_steps or 64  # ~10 s of CPU work on this class of host
            r = cpu_reference_run(n, 1, sample_pairs=min(4, n))
            r1 = cpu_reference_run(4, 1, sample_pairs=1, threads=1)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
                   "one_thread": {"value": r1["value"], "cores": 1, "sample": r1["sample"]}}
        if not args.no_extras:
            for name, fn in (("c3", lambda: leg_c3(pm, ctx, util, dev)), ("c5", lambda: leg_c5(pm, ctx, util)), ("dropin", leg_dropin)):
                try:
                    extras[name] = fn()
                except Exception as e:  # extras never cost the headline line
                    extras[name] = {"error": repr(e)[:300]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 points / f64 reductions", "data": "synthetic",
            "config": make_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes * world,
                    "d2h_bytes_per_step": int(POOL * 480), "ms_per_step": ms_e2e / args.steps,
                    "entry": "pgs_icp_run_batch_multi (host-resident pairs, pinned), chunked uploads inside the call",
                    "results_equal_resident": bool(np.array_equal(allrec, allrec_e2e)),
                    "wall_ms_each_step_incl_warmup": e2e_wall},
            "resident_wall_ms_each_step_incl_warmup": resident_wall,
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "bench_parity": parity, "results_sha256": sha, "knn_queries_per_s": knn_qps,
            "iterations_mean": float(iters.mean()), "pairs_ok": ok, "single_pair_latency_ms": ms_one / 10.0,
            "setup_s": round(setup_s, 1)}
    if extras:
        line["c3"] = extras.get("c3")
        line["c5"] = extras.get("c5")
        dp = extras.get("dropin") or {}
        line["dropin_ms"] = dp.get("dropin_ms")
        line["dropin"] = dp
    emit(line, out_fd)
    if multi:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
