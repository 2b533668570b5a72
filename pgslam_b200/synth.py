"""Synthetic Velodyne-like scans (SURVEY.md §8d "Synthetic inputs").

The reference ships no data (29 files, none are clouds), so every parity and
bench input is generated here, ONCE, on the host in float64 and cast to
float32; the oracle and the CUDA path are always fed the same arrays.

Scene: a closed 60 m x 40 m x 12 m room (so every ray returns and point counts
are exact) with axis-aligned boxes and vertical cylinders placed from `seed`.
Sensor: `beams` elevations x `az_steps` azimuths, per-ray azimuth jitter and
Gaussian range noise, counter-based RNG (Philox keyed by (seed, scan)).
"""
from __future__ import annotations

import numpy as np

ROOM = np.array([[-30.0, 30.0], [-20.0, 20.0], [0.0, 12.0]])
SENSOR_HEIGHT = 1.73


def _rng(seed: int, stream: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream)]))


def make_scene(seed: int, n_boxes: int = 12, n_cyl: int = 8):
    """Boxes (lo, hi) and cylinders (cx, cy, r, h); kept clear of the sensor area."""
    g = _rng(seed, 0xC0FFEE)
    boxes = []
    while len(boxes) < n_boxes:
        c = np.array([g.uniform(-26, 26), g.uniform(-17, 17)])
        if np.hypot(*c) < 6.0:
            continue
        half = g.uniform(0.5, 2.5, size=2)
        h = g.uniform(0.8, 6.0)
        boxes.append((np.array([c[0] - half[0], c[1] - half[1], 0.0]),
                      np.array([c[0] + half[0], c[1] + half[1], h])))
    cyls = []
    while len(cyls) < n_cyl:
        c = np.array([g.uniform(-26, 26), g.uniform(-17, 17)])
        if np.hypot(*c) < 6.0:
            continue
        cyls.append((c[0], c[1], g.uniform(0.2, 1.0), g.uniform(2.0, 10.0)))
    return boxes, cyls


def pose_matrix(t, yaw=0.0, pitch=0.0, roll=0.0) -> np.ndarray:
    """4x4 T_world_sensor = Rz(yaw) Ry(pitch) Rx(roll), translation t."""
    cy, sy = np.cos(yaw), np.sin(yaw)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cr, sr = np.cos(roll), np.sin(roll)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1.0]])
    Ry = np.array([[cp, 0, sp], [0, 1.0, 0], [-sp, 0, cp]])
    Rx = np.array([[1.0, 0, 0], [0, cr, -sr], [0, sr, cr]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = t
    return T


def _raycast(o: np.ndarray, d: np.ndarray, scene) -> np.ndarray:
    """Range along unit directions d (N,3) from origin o to the closest surface."""
    boxes, cyls = scene
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        # room: inside a closed box -> exit distance
        t_hi = (ROOM[:, 1] - o) * inv
        t_lo = (ROOM[:, 0] - o) * inv
        t_room = np.min(np.maximum(t_hi, t_lo), axis=1)
        best = t_room
        for lo, hi in boxes:
            ta = (lo - o) * inv
            tb = (hi - o) * inv
            tn = np.max(np.minimum(ta, tb), axis=1)
            tf = np.min(np.maximum(ta, tb), axis=1)
            hit = (tn <= tf) & (tn > 0)
            best = np.where(hit & (tn < best), tn, best)
        for cx, cy, r, h in cyls:
            ox, oy = o[0] - cx, o[1] - cy
            a = d[:, 0] ** 2 + d[:, 1] ** 2
            b = 2 * (ox * d[:, 0] + oy * d[:, 1])
            c = ox * ox + oy * oy - r * r
            disc = b * b - 4 * a * c
            sq = np.sqrt(np.maximum(disc, 0.0))
            for t in ((-b - sq) / (2 * a), (-b + sq) / (2 * a)):
                z = o[2] + t * d[:, 2]
                hit = (disc > 0) & (t > 0) & (z >= 0) & (z <= h) & (a > 1e-12)
                best = np.where(hit & (t < best), t, best)
    return best


def velodyne_scan(seed: int, scan: int, T_world_sensor: np.ndarray, beams: int = 64,
                  az_steps: int = 1875, range_noise: float = 0.02,
                  az_jitter_deg: float = 0.02, scene=None) -> np.ndarray:
    """One scan in the SENSOR frame as a 4 x N float32 array (PM `features`).

    64-beam: elevations -24.8..+2.0 deg (HDL-64E-like); 16-beam: -15..+15 deg;
    otherwise -25..+15 deg.  N = beams * az_steps exactly.
    """
    if scene is None:
        scene = make_scene(seed)
    g = _rng(seed, 1 + scan)
    if beams == 64:
        el = np.deg2rad(np.linspace(-24.8, 2.0, beams))
    elif beams == 16:
        el = np.deg2rad(np.linspace(-15.0, 15.0, beams))
    else:
        el = np.deg2rad(np.linspace(-25.0, 15.0, beams))
    az = np.linspace(0.0, 2 * np.pi, az_steps, endpoint=False)
    azg, elg = np.meshgrid(az, el, indexing="ij")  # azimuth-major, like a spinning sensor
    azg = azg + np.deg2rad(g.uniform(-az_jitter_deg, az_jitter_deg, size=azg.shape))
    ds = np.stack([np.cos(elg) * np.cos(azg), np.cos(elg) * np.sin(azg), np.sin(elg)], axis=-1).reshape(-1, 3)
    R = T_world_sensor[:3, :3]
    o = T_world_sensor[:3, 3]
    dw = ds @ R.T
    rng = _raycast(o, dw, scene)
    rng = rng + g.normal(0.0, range_noise, size=rng.shape)
    rng = np.clip(rng, 1.0, 80.0)
    pts = ds * rng[:, None]
    out = np.ones((4, pts.shape[0]), dtype=np.float32, order="F")
    out[:3, :] = pts.T.astype(np.float32)
    return out


def scan_pair(seed: int, beams: int = 64, az_steps: int = 1875, scene_seed: int | None = None):
    """(reading, reference, T_ref_reading_truth): same scene from two poses.

    Reference pose: origin at sensor height.  Reading pose perturbed by
    t ~ U(-0.5,0.5)^3 (z x0.2), yaw U(-5,5) deg, roll/pitch U(-1,1) deg.
    ICP (init = identity) should recover T_ref_reading = pose_ref^-1 pose_reading.
    """
    scene = make_scene(seed if scene_seed is None else scene_seed)
    g = _rng(seed, 0xBEEF)
    T_ref = pose_matrix([0.0, 0.0, SENSOR_HEIGHT])
    dt = g.uniform(-0.5, 0.5, size=3) * np.array([1.0, 1.0, 0.2])
    yaw = np.deg2rad(g.uniform(-5, 5))
    roll, pitch = np.deg2rad(g.uniform(-1, 1, size=2))
    T_rd = pose_matrix(np.array([0.0, 0.0, SENSOR_HEIGHT]) + dt, yaw, pitch, roll)
    reference = velodyne_scan(seed, 0, T_ref, beams, az_steps, scene=scene)
    reading = velodyne_scan(seed, 1, T_rd, beams, az_steps, scene=scene)
    truth = np.linalg.inv(T_ref) @ T_rd
    return reading, reference, truth


def trajectory(n: int, step: float = 0.5, turn_deg: float = 2.0):
    """Closed-ish loop of sensor poses for the sequential odometry config (C3)."""
    poses = []
    x, y, yaw = 8.0, 0.0, np.pi / 2
    for _ in range(n):
        poses.append(pose_matrix([x, y, SENSOR_HEIGHT], yaw))
        yaw += np.deg2rad(turn_deg)
        x += step * np.cos(yaw) * 0.56
        y += step * np.sin(yaw) * 0.56
    return poses
