"""The N>1 host logic on CPU: 2 ranks over gloo (rendezvous on 127.0.0.1) -- camera sharding,
frame scatter from the ingest rank, max-over-ranks timing, whole-job frame count."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from watsor_b200.parallel import (camera_shard, engine_scatter_frames, init_engine_comm, max_over_ranks,
                                      scatter_frames, sum_over_ranks)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    C, H, W = 3, 8, 12
    cams = camera_shard(rank, world, C)
    per_rank = None
    if rank == 0:     # the ingest rank owns every camera's frame: value = global camera id
        per_rank = [torch.stack([torch.full((H, W, 3), r * C + c, dtype=torch.uint8) for c in range(C)])
                    for r in range(world)]
    recv = torch.empty((C, H, W, 3), dtype=torch.uint8)
    scatter_frames(recv, per_rank, src=0)
    ok = all(int(recv[c].min()) == cams[c] == int(recv[c].max()) for c in range(C))
    # rendezvous of the library's own communicator (wb_comm_init): the id is made by the C-ABI on the root and must
    # reach every rank unchanged; the engine here records the calls (the collective itself needs GPUs:
    # tests/test_gpu_scatter.py)
    class RecordingEngine:
        def __init__(self):
            from watsor_b200 import _lib
            self.lib, self.calls = _lib.load(), []

        def comm_unique_id(self):
            import ctypes
            buf = (ctypes.c_uint8 * 128)()
            assert self.lib.wb_comm_unique_id(buf) == 0
            return bytes(buf)

        def comm_init(self, rank, world, unique_id):
            self.calls.append(('init', rank, world, bytes(unique_id)))

        def scatter_frames(self, root, send, recv, nbytes, cuda_stream=0):
            self.calls.append(('scatter', root, None if send is None else len(send), nbytes))

    eng = RecordingEngine()
    init_engine_comm(eng, rank, world)
    engine_scatter_frames(eng, recv, per_rank, root=0)
    ident = eng.calls[0][3]
    gathered = [None] * world
    dist.all_gather_object(gathered, ident)
    ok = ok and eng.calls[0][:3] == ('init', rank, world) and len(ident) == 128 and any(ident) and \
        all(g == ident for g in gathered) and \
        eng.calls[1] == ('scatter', 0, world if rank == 0 else None, C * H * W * 3)
    slowest = max_over_ranks(1.0 + rank)
    frames = sum_over_ranks(C * 10)
    dist.barrier()
    out.put((rank, cams, ok, slowest, frames))
    dist.destroy_process_group()


def test_two_ranks_shard_scatter_and_reduce():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [[0, 1, 2], [3, 4, 5]]          # disjoint, complete shards
    assert all(r[2] for r in res)                                  # every rank received its own cameras
    assert all(r[3] == 2.0 for r in res)                           # max over ranks
    assert all(r[4] == 60.0 for r in res)                          # whole-job frame count


def test_single_process_paths_need_no_process_group():
    from watsor_b200.parallel import camera_shard, max_over_ranks, scatter_frames
    assert camera_shard(0, 1, 8) == list(range(8))
    assert max_over_ranks(3.5) == 3.5
    dst = torch.zeros((2, 4, 4, 3), dtype=torch.uint8)
    scatter_frames(dst, [torch.ones((2, 4, 4, 3), dtype=torch.uint8)])
    assert int(dst.sum()) == 96
