"""Multi-GPU plumbing on CPU: the channel partition and the optional gather of
packed spectra, exercised with world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest

from fpga_real_time_fft_analyzer_b200.sharding import channel_range


def test_channel_range_partitions_exactly():
    for total in (1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            spans = [channel_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        channel_range(8, 2, 2)


def _worker(rank, world, port, total, width, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fpga_real_time_fft_analyzer_b200.sharding import channel_range, gather_frames
    a, b = channel_range(total, rank, world)
    ch = torch.arange(a, b, dtype=torch.int64)[:, None]
    local = ((ch * 131 + torch.arange(width)[None, :]) % 251).to(torch.uint8)
    got = gather_frames(local, total, dst=0)
    dist.barrier()
    if rank == 0:
        q.put(got.numpy())
    else:
        assert got is None
    dist.destroy_process_group()


def test_gather_frames_world2_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    total, width, world = 7, 64, 2          # ragged: 4 + 3 channels
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, width, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = ((np.arange(total)[:, None] * 131 + np.arange(width)[None, :]) % 251).astype(np.uint8)
    assert np.array_equal(got, want)
