"""Multi-GPU execution of the hot path: independent frames, one process per GPU.

A single frame does not shard (hash/bitmap index, rulebook and first-come voxel order are
global to the frame; RoI sampling needs the whole BEV map), exactly like the reference,
whose multi-GPU inference is one model replica per rank over a partition of the dataset
(tools/test.py:227-233).  So: REPLICAS ONLY -- frame i runs on rank i mod world, weights are
replicated, and there is no collective on the data path.  The only communication is the
optional gather of per-frame results (tiny) for the caller.
"""
import torch
import torch.distributed as dist


def frames_of_rank(n_frames, rank, world):
    """Round-robin partition: frame i -> rank i % world."""
    return list(range(rank, n_frames, world))


def run_partitioned(n_frames, fn, rank=None, world=None):
    """Run fn(frame_index) for this rank's frames; returns {frame_index: result}."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    return {i: fn(i) for i in frames_of_rank(n_frames, rank, world)}


def gather_results(local, n_frames):
    """All ranks receive the full ordered list of per-frame results (CPU tensors / picklable).
    Uses all_gather_object: control-plane sized messages only."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [local[i] for i in range(n_frames)]
    bucket = [None] * dist.get_world_size()
    dist.all_gather_object(bucket, local)
    merged = {}
    for part in bucket:
        merged.update(part)
    return [merged[i] for i in range(n_frames)]


def max_over_ranks(value, device):
    """MAX all-reduce of a scalar timing (ms) across ranks (device-side, NCCL or gloo)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
