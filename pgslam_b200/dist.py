"""Sharding of independent scan pairs across the GPUs of one box.

Pairs (loop-closure candidates, multi-sequence odometry) share nothing, so the
data path has NO collective: each rank registers its contiguous block of pairs
on its own GPU.  The only exchange is the gather of the per-pair results
(4x4 pose, iterations, status, overlap ...) at the end of a batch, over
torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

RESULT_WIDTH = 16 + 36 + 6  # T (col-major), covariance, iterations, status, max_iter, overlap, residual, ratio


def shard_range(n_pairs: int, rank: int, world: int) -> range:
    """Contiguous block partition: pair i -> rank i // ceil(P / G)."""
    per = -(-n_pairs // world)
    return range(min(rank * per, n_pairs), min((rank + 1) * per, n_pairs))


def pack_results(results: list[dict]) -> np.ndarray:
    out = np.zeros((len(results), RESULT_WIDTH), np.float64)
    for i, r in enumerate(results):
        out[i, :16] = np.asarray(r["T"]).ravel(order="F")
        out[i, 16:52] = np.asarray(r.get("covariance", np.zeros((6, 6)))).ravel(order="F")
        out[i, 52:58] = (r["iterations"], r["status"], float(r.get("max_iterations_reached", False)),
                         r.get("overlap", 0.0), r.get("residual", 0.0), r.get("weighted_point_used_ratio", 0.0))
    return out


def pack_records(rec: np.ndarray) -> np.ndarray:
    """pack_results for the structured pgs_icp_result array of pm.ICP.compute_batch_array
    (no per-pair Python objects: a 4096-pair batch is packed in a few vector operations)."""
    out = np.zeros((len(rec), RESULT_WIDTH), np.float64)
    if len(rec):
        out[:, :16] = rec["T"]
        out[:, 16:52] = rec["covariance"]
        out[:, 52] = rec["iterations"]
        out[:, 53] = rec["status"]
        out[:, 54] = rec["max_iterations_reached"] != 0
        out[:, 55] = rec["overlap"]
        out[:, 56] = rec["residual"]
        out[:, 57] = rec["weighted_point_used_ratio"]
    return out


def unpack_results(arr: np.ndarray) -> list[dict]:
    out = []
    for row in np.asarray(arr):
        out.append(dict(T=row[:16].reshape(4, 4).T.copy(), covariance=row[16:52].reshape(6, 6).T.copy(),
                        iterations=int(row[52]), status=int(row[53]), max_iterations_reached=bool(row[54]),
                        overlap=float(row[55]), residual=float(row[56]), weighted_point_used_ratio=float(row[57])))
    return out


def gather_results(local: list[dict], n_pairs: int, device=None) -> list[dict]:
    """all_gather of the per-pair result records; every rank gets all P of them,
    in pair order.  Works under NCCL (device tensors) and gloo (CPU tensors)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return list(local)
    per = -(-n_pairs // world)
    buf = np.zeros((per, RESULT_WIDTH), np.float64)
    buf[:, 53] = -1.0  # padding rows
    if local:
        buf[:len(local)] = pack_results(local)
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    out = torch.empty((world * per, RESULT_WIDTH), dtype=torch.float64, device=t.device)
    dist.all_gather_into_tensor(out, t)
    rows = out.cpu().numpy()
    keep = [rows[r * per + j] for r in range(world) for j in range(len(shard_range(n_pairs, r, world)))]
    return unpack_results(np.asarray(keep))


def gather_rows(local: np.ndarray, n_pairs: int, device=None) -> np.ndarray:
    """all_gather of packed result rows (this rank's block, in pair order) -> (n_pairs, RESULT_WIDTH)
    on every rank.  NCCL with `device`, gloo without."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return np.ascontiguousarray(local)
    per = -(-n_pairs // world)
    buf = np.zeros((per, RESULT_WIDTH), np.float64)
    buf[:, 53] = -1.0  # padding rows
    buf[:len(local)] = local
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    out = torch.empty((world * per, RESULT_WIDTH), dtype=torch.float64, device=t.device)
    dist.all_gather_into_tensor(out, t)
    rows = out.cpu().numpy()
    keep = [rows[r * per: r * per + len(shard_range(n_pairs, r, world))] for r in range(world)]
    return np.ascontiguousarray(np.concatenate(keep, axis=0))


def gather_records(rec: np.ndarray, n_pairs: int, device=None) -> np.ndarray:
    """gather_rows of a structured pgs_icp_result array."""
    return gather_rows(pack_records(rec), n_pairs, device)


def register_sharded(run_batch, pairs: list, device=None) -> list[dict]:
    """run_batch(list of (reading, reference)) -> list of result dicts, applied
    to this rank's block; results of all ranks are gathered in pair order."""
    import torch.distributed as dist

    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    mine = [pairs[i] for i in shard_range(len(pairs), rank, world)]
    local = run_batch(mine) if mine else []
    return gather_results(local, len(pairs), device)
