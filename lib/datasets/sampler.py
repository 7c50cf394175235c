"""Samplers for multi-GPU inference over a pair dataset: pairs are independent, so every rank takes
a contiguous slice of the pair list (the same split as mickey_b200.dist.shard_range) and no pair is dropped or
duplicated; with world_size 1 it is a plain sequential sampler (the reference's DataLoader(sampler=None))."""
from torch.utils.data import Sampler

from mickey_b200.dist import shard_range, world


class ShardedSequentialSampler(Sampler):
    def __init__(self, data_source, rank=None, world_size=None):
        if rank is None or world_size is None:
            rank, world_size = world()
        self.start, self.end = shard_range(len(data_source), rank, world_size)

    def __iter__(self):
        return iter(range(self.start, self.end))

    def __len__(self):
        return self.end - self.start
