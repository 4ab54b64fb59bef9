"""Multi-GPU plumbing: one process per GPU, cameras sharded by rank, `torch.distributed` only for
(i) the optional frame scatter from the ingest rank and (ii) timing reductions.

The reference's multi-accelerator model is N independent detector processes pulling from one queue
(watsor/detection/detector.py:40-50); there is no collective on its data path.  Sharding cameras by
rank keeps that property: the detection path itself never communicates.  The scatter is the
"engine frame scatter" of BASELINE.json's north star: rank `src` owns every camera's frame of a
tick and sends each rank its `[C, H, W, 3]` uint8 slab (NCCL over NVLink on GPUs, gloo in tests).
"""
import torch
import torch.distributed as dist


def camera_shard(rank, world, cameras_per_rank):
    """Global camera ids served by `rank` (camera c -> rank c // cameras_per_rank)."""
    assert 0 <= rank < world
    return list(range(rank * cameras_per_rank, (rank + 1) * cameras_per_rank))


def scatter_frames(recv, per_rank_frames, src=0):
    """recv: this rank's `[C,H,W,3]` uint8 tensor; per_rank_frames: list of such tensors on `src`."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        recv.copy_(per_rank_frames[0])
        return
    dist.scatter(recv, per_rank_frames if dist.get_rank() == src else None, src=src)


def max_over_ranks(value, device='cpu'):
    """Every multi-GPU time is the max over ranks (the slowest rank defines the step)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device='cpu'):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
