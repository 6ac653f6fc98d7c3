"""Sequence sharding across the GPUs of one box.

The path has no exchange step: `detect` is per frame, `matchV` needs two frames of the SAME sequence.
So sequences are dealt round-robin (`seq % world`), every rank keeps its sequences device-resident and
no data-path collective exists.  torch.distributed is used for the launch plumbing only: a barrier
around the timed region, MAX over ranks of the device time, and a gather of the small per-rank result
summaries (counts, checksums) onto rank 0.
"""
import hashlib

import numpy as np


def sequences_for_rank(n_sequences, rank, world):
    """Round-robin ownership: rank r owns sequences r, r+world, ...  (config 5: 64 sequences, 8 per GPU)."""
    if not (0 <= rank < world):
        raise ValueError("rank outside [0, world)")
    return list(range(rank, n_sequences, world))


def owner_of(seq, world):
    return seq % world


def digest(*arrays):
    """Order-sensitive 64-bit checksum of result arrays (keypoints, descriptors, matches)."""
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return int.from_bytes(h.digest()[:8], "little") >> 1


def max_over_ranks(value, device=None):
    """MAX over ranks of a scalar (the timed region's device milliseconds)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_summaries(summary):
    """Gather {seq: (n_frames, n_keypoints, n_matches, digest)} dicts onto every rank and merge them.
    Raises if two ranks claim the same sequence (a sharding bug)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(summary)
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, summary)
    merged = {}
    for r, part in enumerate(parts):
        for seq, val in part.items():
            if seq in merged:
                raise RuntimeError(f"sequence {seq} processed by two ranks")
            if owner_of(seq, dist.get_world_size()) != r:
                raise RuntimeError(f"sequence {seq} processed by rank {r}, owner is {owner_of(seq, dist.get_world_size())}")
            merged[seq] = val
    return merged
