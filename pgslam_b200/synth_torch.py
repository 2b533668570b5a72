"""The synthetic Velodyne-like scan generator of `synth.py`, evaluated with torch on the GPU.

BASELINE config C4 needs 4096 DISTINCT 120k-point scan pairs (15.7 GB of clouds): the numpy
generator takes about a second per pair on one host core, this one a few seconds for the whole
pool.  Same scene model, same sensor model, same formulas (float64, cast to float32 at the end);
the random numbers come from a per-scan `torch.Generator` on the device, so the clouds are NOT
bit-identical to `synth.scan_pair`'s — they do not have to be: the parity checks always hand the
CPU checker the very arrays the CUDA path was given.  A pair depends only on its seed (never on the
batch it was generated in or on the rank), which is what lets bench.py compare results across
1/2/4/8 GPUs bit for bit.

Test / bench infrastructure: nothing in the product path imports this module.
"""
from __future__ import annotations

import numpy as np
import torch

from . import synth


def _elevations(beams: int) -> np.ndarray:
    if beams == 64:
        return np.deg2rad(np.linspace(-24.8, 2.0, beams))
    if beams == 16:
        return np.deg2rad(np.linspace(-15.0, 15.0, beams))
    return np.deg2rad(np.linspace(-25.0, 15.0, beams))


def _scene_tensors(seeds, dev):
    """boxes (B, nb, 2, 3) and cylinders (B, nc, 4) of the pairs' scenes (synth.make_scene)."""
    bl, cl = [], []
    for s in seeds:
        boxes, cyls = synth.make_scene(int(s))
        bl.append(np.stack([np.stack([lo, hi]) for lo, hi in boxes]))
        cl.append(np.array(cyls, dtype=np.float64))
    return (torch.tensor(np.stack(bl), dtype=torch.float64, device=dev),
            torch.tensor(np.stack(cl), dtype=torch.float64, device=dev))


def _raycast(o, d, boxes, cyls):
    """synth._raycast for a batch: o (B,3), d (B,N,3), boxes (B,nb,2,3), cyls (B,nc,4) -> (B,N)."""
    room = torch.tensor(synth.ROOM, dtype=torch.float64, device=d.device)
    inv = 1.0 / d
    ob = o[:, None, :]
    t_hi = (room[:, 1] - ob) * inv
    t_lo = (room[:, 0] - ob) * inv
    best = torch.minimum(torch.maximum(t_hi, t_lo).amin(dim=2), torch.full_like(d[..., 0], float("inf")))
    for k in range(boxes.shape[1]):  # objects one at a time: the update order is synth.py's
        lo = boxes[:, k, 0][:, None, :]
        hi = boxes[:, k, 1][:, None, :]
        ta = (lo - ob) * inv
        tb = (hi - ob) * inv
        tn = torch.minimum(ta, tb).amax(dim=2)
        tf = torch.maximum(ta, tb).amin(dim=2)
        hit = (tn <= tf) & (tn > 0)
        best = torch.where(hit & (tn < best), tn, best)
    for k in range(cyls.shape[1]):
        cx, cy, r, h = (cyls[:, k, j][:, None] for j in range(4))
        ox, oy = o[:, 0:1] - cx, o[:, 1:2] - cy
        a = d[..., 0] ** 2 + d[..., 1] ** 2
        b = 2 * (ox * d[..., 0] + oy * d[..., 1])
        c = ox * ox + oy * oy - r * r
        disc = b * b - 4 * a * c
        sq = torch.sqrt(torch.clamp(disc, min=0.0))
        for t in ((-b - sq) / (2 * a), (-b + sq) / (2 * a)):
            z = o[:, 2:3] + t * d[..., 2]
            hit = (disc > 0) & (t > 0) & (z >= 0) & (z <= h) & (a > 1e-12)
            best = torch.where(hit & (t < best), t, best)
    return best


def _grids(dev, beams, az_steps):
    el = torch.tensor(_elevations(beams), dtype=torch.float64, device=dev)
    az = torch.tensor(np.linspace(0.0, 2 * np.pi, az_steps, endpoint=False), dtype=torch.float64, device=dev)
    return az[:, None].expand(az_steps, beams), el[None, :].expand(az_steps, beams)  # azimuth-major, like a spinning sensor


def _scans(dev, grids, rng_keys, Ts, boxes, cyls, beams, az_steps, range_noise, az_jitter_deg):
    """One scan per entry: rng_keys[i] seeds scan i's jitter and noise, Ts[i] is its T_world_sensor,
    boxes / cyls hold its scene.  Returns (B, N, 4) float32, {x,y,z,1} per point, sensor frame."""
    azg0, elg = grids
    n = beams * az_steps
    jit, noise = [], []
    for key in rng_keys:
        g = torch.Generator(device=dev)
        g.manual_seed(int(key) & 0x7FFFFFFFFFFFFFFF)
        jit.append(torch.rand((az_steps, beams), generator=g, dtype=torch.float64, device=dev))
        noise.append(torch.randn((n,), generator=g, dtype=torch.float64, device=dev))
    jitter = (torch.stack(jit) * 2.0 - 1.0) * np.deg2rad(az_jitter_deg)
    azg = azg0[None] + jitter
    ce = torch.cos(elg)[None]
    ds = torch.stack([ce * torch.cos(azg), ce * torch.sin(azg), torch.sin(elg)[None].expand_as(azg)],
                     dim=-1).reshape(len(rng_keys), n, 3)
    Tt = torch.tensor(np.stack(Ts), dtype=torch.float64, device=dev)
    # explicit sums, not a batched GEMM: the result must not depend on the batch size
    R = Tt[:, :3, :3]
    dw = torch.stack([ds[..., 0] * R[:, i, 0:1] + ds[..., 1] * R[:, i, 1:2] + ds[..., 2] * R[:, i, 2:3]
                      for i in range(3)], dim=-1)
    rng = _raycast(Tt[:, :3, 3], dw, boxes, cyls)
    rng = torch.clamp(rng + torch.stack(noise) * range_noise, 1.0, 80.0)
    pts = torch.ones((len(rng_keys), n, 4), dtype=torch.float32, device=dev)
    pts[..., :3] = (ds * rng[..., None]).to(torch.float32)
    return pts


def scan_pairs(seeds, device, beams: int = 64, az_steps: int = 1875, range_noise: float = 0.02,
               az_jitter_deg: float = 0.02, chunk: int = 16):
    """[(reading, reference)] as (N, 4) float32 device tensors ({x,y,z,1} per point: the C ABI's
    layout) and the list of true T_ref_reading (numpy 4x4), one pair per seed."""
    dev = torch.device(device)
    seeds = [int(s) for s in seeds]
    grids = _grids(dev, beams, az_steps)
    out, truths = [], []
    T_ref = synth.pose_matrix([0.0, 0.0, synth.SENSOR_HEIGHT])
    for c0 in range(0, len(seeds), chunk):
        cs = seeds[c0:c0 + chunk]
        boxes, cyls = _scene_tensors(cs, dev)
        poses = []
        for s in cs:
            g = synth._rng(s, 0xBEEF)
            dt = g.uniform(-0.5, 0.5, size=3) * np.array([1.0, 1.0, 0.2])
            yaw = np.deg2rad(g.uniform(-5, 5))
            roll, pitch = np.deg2rad(g.uniform(-1, 1, size=2))
            T_rd = synth.pose_matrix(np.array([0.0, 0.0, synth.SENSOR_HEIGHT]) + dt, yaw, pitch, roll)
            poses.append(T_rd)
            truths.append(np.linalg.inv(T_ref) @ T_rd)
        clouds = {}
        for scan, Ts in ((0, [T_ref] * len(cs)), (1, poses)):
            keys = [s * 2654435761 + 97 * scan + 12345 for s in cs]
            clouds[scan] = _scans(dev, grids, keys, Ts, boxes, cyls, beams, az_steps, range_noise, az_jitter_deg)
        for j in range(len(cs)):
            out.append((clouds[1][j].contiguous(), clouds[0][j].contiguous()))
    return out, truths


def trajectory_scans(seed, poses, device, beams: int = 64, az_steps: int = 1875, range_noise: float = 0.02,
                     az_jitter_deg: float = 0.02, chunk: int = 16):
    """One scan per pose (T_world_sensor, numpy 4x4) of ONE scene (`seed`): the sequential-odometry
    input of BASELINE config C3.  Returns a list of (N, 4) float32 device tensors, sensor frame."""
    dev = torch.device(device)
    grids = _grids(dev, beams, az_steps)
    out = []
    for c0 in range(0, len(poses), chunk):
        Ts = poses[c0:c0 + chunk]
        boxes, cyls = _scene_tensors([seed] * len(Ts), dev)
        keys = [int(seed) * 2654435761 + 7919 * (c0 + j) + 777 for j in range(len(Ts))]
        pts = _scans(dev, grids, keys, Ts, boxes, cyls, beams, az_steps, range_noise, az_jitter_deg)
        out += [pts[j].contiguous() for j in range(len(Ts))]
    return out
