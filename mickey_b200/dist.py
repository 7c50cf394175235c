"""Multi-GPU plumbing: pairs are independent (no cross-pair term anywhere on the path; BatchNorm is
folded), so a batch is sharded contiguously by rank, every rank holds a full weight replica and runs the
whole pipeline locally, and the only data-path collective is ONE all-gather of the packed poses
([B_local, 13] fp32 = R row-major 9 | t 3 | inliers 1 -> 52 B per pair).
Works with backend 'nccl' (GPU) and 'gloo' (CPU tests, world_size 2)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_pairs: int, rank: int = None, world_size: int = None) -> Tuple[int, int]:
    """Contiguous [start, end) of the global batch owned by `rank` (remainder spread over the first ranks)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(n_pairs, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(data: dict, rank: int = None, world_size: int = None) -> dict:
    """Slice every batched tensor of a MicKey data dict to this rank's pairs."""
    n = data["image0"].shape[0]
    s, e = shard_range(n, rank, world_size)
    return {k: (v[s:e] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n else v) for k, v in data.items()}


def pack_pose(R: torch.Tensor, t: torch.Tensor, inliers: torch.Tensor) -> torch.Tensor:
    B = R.shape[0]
    return torch.cat([R.reshape(B, 9), t.reshape(B, 3), inliers.reshape(B, 1)], dim=1).contiguous()


def unpack_pose(packed: torch.Tensor):
    B = packed.shape[0]
    return packed[:, :9].reshape(B, 3, 3), packed[:, 9:12].reshape(B, 1, 3), packed[:, 12:13]


def gather_poses(packed: torch.Tensor, n_pairs: int = None) -> torch.Tensor:
    """All-gather of [B_local, 13] -> [B_global, 13] in rank order -- still ONE collective.  With `n_pairs` (the
    global batch) the shards may be ragged (shard_range spreads the remainder over the first ranks): every rank
    pads its block to ceil(n_pairs / world) rows and the padding rows are dropped after the gather.  Without it
    every rank must hold the same number of pairs."""
    rank, ws = world()
    if ws == 1:
        return packed
    packed = packed.contiguous()
    if n_pairs is None:
        out = torch.empty(ws * packed.shape[0], packed.shape[1], dtype=packed.dtype, device=packed.device)
        dist.all_gather_into_tensor(out, packed)
        return out
    rows = -(-n_pairs // ws)
    block = packed
    if packed.shape[0] != rows:
        block = packed.new_zeros(rows, packed.shape[1])
        block[:packed.shape[0]] = packed
    out = torch.empty(ws * rows, packed.shape[1], dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, block)
    if n_pairs == ws * rows:
        return out
    keep = [out[r * rows:r * rows + (e - s)] for r in range(ws) for s, e in [shard_range(n_pairs, r, ws)]]
    return torch.cat(keep, dim=0)


def step_plan(n_pairs: int, world_size: int, batch_size: int):
    """For a pair list sharded contiguously by rank (shard_range) and consumed in batches of `batch_size`: the number of
    steps every rank has to join (the longest shard decides) and, per step, how many real rows each rank contributes to
    the step's one all-gather (0 once its shard is exhausted).  Every rank computes the same plan, so the collective
    count matches without any extra communication (tools/run_submission.py)."""
    shard_len = [e - s for s, e in (shard_range(n_pairs, r, world_size) for r in range(world_size))]
    n_steps = max(-(-l // batch_size) for l in shard_len) if n_pairs > 0 else 0
    rows = [[max(0, min(batch_size, l - step * batch_size)) for l in shard_len] for step in range(n_steps)]
    return n_steps, rows


def forward_sharded(model, data: dict, return_inliers: bool = False):
    """Run `model` on this rank's shard of a global batch and return the globally gathered (R, t, inliers)."""
    local = shard_batch(data)
    n = data["image0"].shape[0]
    if local["image0"].shape[0] == 0:
        # more ranks than pairs: this rank contributes an empty block (it still joins the collective)
        dev = data["image0"].device
        if dist.get_backend() == "nccl" and dev.type != "cuda":
            dev = torch.device("cuda", torch.cuda.current_device())
        mine = torch.zeros(0, 13, dtype=torch.float32, device=dev)
    else:
        R, t = model(local, return_inliers=return_inliers)
        mine = pack_pose(R, t, local["inliers"])
    return unpack_pose(gather_poses(mine, n_pairs=n))
