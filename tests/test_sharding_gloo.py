"""CPU, world_size 2 over gloo: the N>1 path's host logic -- contiguous batch shards, scatter from and
gather to rank 0 in batch order, per-rank plan replicas -- with the test-only emulation library doing
the transforms."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import ROOT

import rustfft_b200 as rb


def test_shard_range_partitions_every_batch():
    for batch in [0, 1, 7, 8, 9, 4096, 65536, 65537]:
        for world in [1, 2, 3, 4, 8]:
            spans = [rb.shard_range(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)  # remainder to low ranks
    assert rb.shard_range(65536, 3, 8) == (3 * 8192, 4 * 8192)  # BASELINE config 5: 8192 per GPU


def _worker(rank, world, port, n, batch, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from util import emu_library, signal, truth
        from rustfft_b200.sharded import ShardedFft

        planner = rb.FftPlanner(np.complex64, lib=emu_library())
        sh = ShardedFft(planner, n, rb.FftDirection.Forward)
        full = torch.from_numpy(signal(n * batch, np.complex64, seed=11)) if rank == 0 else None
        mine = sh.scatter(full, batch, root=0, device="cpu", dtype=torch.complex64)
        lo, hi = sh.my_range(batch)
        assert mine.numel() == (hi - lo) * n
        sh.process_local(mine)
        out = sh.gather(mine, batch, root=0)
        if rank == 0:
            x = signal(n * batch, np.complex64, seed=11)
            ref = truth(x, n, False)
            err = np.linalg.norm(out.numpy() - ref) / np.linalg.norm(ref)
            # every transform equals the single-process result bit for bit (same kernels, same order)
            solo = x.copy()
            planner.plan_fft_forward(n).process(solo)
            q.put((float(err), bool(np.array_equal(out.numpy(), solo))))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,batch", [(256, 9), (1000, 4), (8192, 3)])
def test_two_ranks_scatter_process_gather(n, batch):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, batch, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(timeout=180) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    err, same = q.get(timeout=5)
    assert err < 4 * 5.96e-8 * np.log2(n) and same
