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


def init_engine_comm(engine, rank=None, world=None, root=0):
    """Create the engine's own NCCL communicator (`wb_comm_init`, include/watsor_b200.h): the root makes the 128-byte
    id through the C-ABI and `torch.distributed` -- any backend, it is only the host-side rendezvous channel --
    carries it to the other ranks.  After this, `engine_scatter_frames` never touches torch."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    box = [engine.comm_unique_id() if rank == root else None]
    dist.broadcast_object_list(box, src=root)
    engine.comm_init(rank, world, box[0])


def engine_scatter_frames(engine, recv, per_rank_frames, root=0, cuda_stream=0):
    """`scatter_frames` through the library's own collective (`wb_scatter_frames`): tensors are only pointer carriers."""
    send = [int(t.data_ptr()) for t in per_rank_frames] if per_rank_frames is not None else None
    engine.scatter_frames(root, send, int(recv.data_ptr()), recv.numel() * recv.element_size(), cuda_stream)


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


def gpu_cpu_affinity(device_index):
    """CPUs on the NUMA node next to GPU `device_index` (the `nvidia-smi topo -m` "CPU Affinity" column), or
    None when it cannot be determined.  NVML first, then sysfs `local_cpulist` of the GPU's PCI function."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByIndex(_physical_index(device_index))
            words = (len(__import__('os').sched_getaffinity(0)) + 4096) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
            cpus = {w * 64 + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
            if cpus:
                return cpus
        finally:
            pynvml.nvmlShutdown()
    except Exception:
        pass
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = '%04x:%02x:%02x.0' % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open('/sys/bus/pci/devices/%s/local_cpulist' % bdf) as f:
            return parse_cpulist(f.read())
    except Exception:
        return None


def _physical_index(device_index):
    """NVML enumerates physical GPUs; CUDA_VISIBLE_DEVICES may renumber them for this process."""
    import os
    vis = os.environ.get('CUDA_VISIBLE_DEVICES')
    if vis:
        ids = [v.strip() for v in vis.split(',') if v.strip()]
        if device_index < len(ids) and ids[device_index].isdigit():
            return int(ids[device_index])
    return device_index


def parse_cpulist(text):
    """'0-31,64-95' -> {0..31, 64..95}"""
    cpus = set()
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa(device_index):
    """Pin the calling process to the CPUs local to its GPU, so that the pinned ingest ring is allocated on (and
    the submitting thread runs on) the NUMA node the GPU's PCIe root hangs off.  Eight unbound ranks doing 7.4 MB
    of H2D per 0.4 ms from the wrong socket cost 23 % of the end-to-end rate at 8 GPUs in round 1.  Returns a
    short description for the bench record; never raises."""
    import os
    try:
        cpus = gpu_cpu_affinity(device_index)
        allowed = os.sched_getaffinity(0)
        if not cpus:
            return 'unbound (no affinity information)'
        use = cpus & allowed
        if not use or use == allowed:
            return 'unbound (GPU-local CPUs = all allowed CPUs, %d)' % len(allowed)
        os.sched_setaffinity(0, use)
        return 'bound to %d GPU-local CPUs (%d..%d)' % (len(use), min(use), max(use))
    except Exception as e:           # affinity is an optimisation, never a failure
        return 'unbound (%s)' % type(e).__name__
