"""2-rank NCCL check of the sharded forward (run under torchrun on a multi-GPU box).

Every rank builds the same global batch, runs its contiguous shard through the CUDA path and all-gathers the
packed poses.  Checks: every rank holds the identical gathered result; a rank's slice of it equals what its own
local forward returned; descriptors of the shard equal the same pairs extracted in a full-batch run (the pose
draws depend on the local pair index, so poses are compared through the deterministic stages)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mickey_b200 import dist as mkdist                      # noqa: E402
from mickey_b200.config import mickey_cfg                   # noqa: E402
from mickey_b200.model import MickeyRelativePose            # noqa: E402
from mickey_b200.weights import synthetic_state_dict        # noqa: E402
from tests.common import synthetic_pair                     # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    dev = torch.device("cuda")
    cfg = mickey_cfg("vits", 8, 64)
    model = MickeyRelativePose(cfg).to(dev)
    model.load_state_dict(synthetic_state_dict(cfg, seed=1))
    B = 2 * world + 1                                        # ragged: ranks get unequal shards
    data = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in synthetic_pair(B, 224, 196, seed=5).items()}
    torch.manual_seed(100 + rank)
    local = mkdist.shard_batch(data)
    R_l, t_l = model(local)
    packed = mkdist.gather_poses(mkdist.pack_pose(R_l, t_l, local["inliers"]), n_pairs=B)
    R, t, inl = mkdist.unpack_pose(packed)
    assert R.shape == (B, 3, 3) and t.shape == (B, 1, 3) and inl.shape[0] == B, (R.shape, t.shape, inl.shape)
    s, e = mkdist.shard_range(B)
    assert torch.equal(R[s:e], R_l) and torch.equal(t[s:e], t_l)
    ref = packed.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(ref, packed), "ranks disagree on the gathered poses"
    full = dict(data)
    model(full)
    d_full = full["dsc0"][s:e].float()
    d_loc = local["dsc0"].float()
    err = ((d_full - d_loc).norm() / d_full.norm()).item()
    assert err < 1e-6, err
    dist.barrier()
    if rank == 0:
        print(f"dist_check ok: world={world} B={B} shard_desc_err={err:.2e} finite={bool(torch.isfinite(packed).all())}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
