"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: contiguous pair sharding and the single
all-gather of packed poses.  The model is replaced by a deterministic stand-in (no CUDA here)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mickey_b200 import dist as mkdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeModel:
    """Maps image content to a pose deterministically so the gather order can be verified."""

    def __call__(self, data, return_inliers=False):
        B = data["image0"].shape[0]
        key = data["image0"].reshape(B, -1)[:, 0]
        R = torch.eye(3)[None].repeat(B, 1, 1) * key.view(B, 1, 1)
        t = torch.stack([key, key + 1, key + 2], dim=1).view(B, 1, 3)
        data["inliers"] = (key * 10).view(B, 1)
        return R, t


def _worker(rank, world, port, n_pairs, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = {"image0": torch.arange(n_pairs, dtype=torch.float32).view(n_pairs, 1, 1, 1).repeat(1, 3, 2, 2),
                "image1": torch.zeros(n_pairs, 3, 2, 2), "K_color0": torch.eye(3)[None].repeat(n_pairs, 1, 1),
                "K_color1": torch.eye(3)[None].repeat(n_pairs, 1, 1), "down_factor": 14}
        s, e = mkdist.shard_range(n_pairs)
        local = mkdist.shard_batch(data)
        assert local["image0"].shape[0] == e - s and local["down_factor"] == 14
        R, t, inl = mkdist.forward_sharded(_FakeModel(), data)
        if rank == 0:
            torch.save({"R": R, "t": t, "inl": inl}, out)
    finally:
        dist.destroy_process_group()


def test_shard_ranges_cover_batch():
    for n in (1, 2, 7, 32, 256):
        for w in (1, 2, 4, 8):
            spans = [mkdist.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1


def test_pack_unpack_roundtrip():
    R, t, i = torch.randn(5, 3, 3), torch.randn(5, 1, 3), torch.randn(5, 1)
    R2, t2, i2 = mkdist.unpack_pose(mkdist.pack_pose(R, t, i))
    assert torch.equal(R, R2) and torch.equal(t, t2) and torch.equal(i, i2)


@pytest.mark.parametrize("n_pairs", [6, 5, 1])
def test_two_rank_gloo_gather_matches_single_process(tmp_path, n_pairs):
    world = 2
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, _free_port(), n_pairs, out), nprocs=world, join=True)
    got = torch.load(out)
    data = {"image0": torch.arange(n_pairs, dtype=torch.float32).view(n_pairs, 1, 1, 1).repeat(1, 3, 2, 2)}
    R, t = _FakeModel()(data)
    assert torch.equal(got["R"], R) and torch.equal(got["t"], t) and torch.equal(got["inl"], data["inliers"])


def test_step_plan_covers_every_pair_once():
    """The per-step row counts of the sharded submission run: every pair appears exactly once, in rank order per step,
    and all ranks agree on the number of collectives."""
    from mickey_b200.dist import shard_range, step_plan
    for n_pairs, world, bs in [(192, 8, 12), (5, 2, 12), (64, 2, 12), (7, 8, 4), (0, 4, 8), (100, 3, 7)]:
        n_steps, rows = step_plan(n_pairs, world, bs)
        assert len(rows) == n_steps and all(len(r) == world for r in rows)
        per_rank = [sum(r[k] for r in rows) for k in range(world)]
        assert per_rank == [e - s for s, e in (shard_range(n_pairs, k, world) for k in range(world))]
        assert sum(per_rank) == n_pairs and all(0 <= x <= bs for r in rows for x in r)
        for k in range(world):                      # a rank's rows are consumed front to back: no gaps
            col = [r[k] for r in rows]
            assert all(col[i] == bs or sum(col[i + 1:]) == 0 for i in range(len(col)))
