"""Helper launched under torchrun (one rank per GPU, NCCL) by tests/test_gpu_parity.py::test_sharded_scatter_fft_gather_two_gpus:
BASELINE config 5's data path at a small size -- the batch starts on rank 0, NCCL point-to-point scatter of contiguous
shards, every rank transforms its shard on its own GPU, gather back to rank 0, result checked on rank 0 against the oracle
and against the single-GPU result (bit for bit: the same kernels run on every GPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

import oracle
import rustfft_b200 as rb
from rustfft_b200.sharded import ShardedFft
from util import rel_l2, signal, strict_bound, truth


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    planner = rb.FftPlanner(np.complex64, device=local)
    for n, batch in [(1 << 16, 67), (1000, 1001), (65537, 5), (1 << 10, 1)]:  # ragged shards; batch 1 = an empty shard on rank 1
        sh = ShardedFft(planner, n)
        x = signal(n * batch, np.complex64, seed=n)
        full = torch.from_numpy(x).to(dev) if rank == 0 else None
        shard = sh.scatter(full, batch, root=0, device=dev, dtype=torch.complex64)
        lo, hi = sh.my_range(batch)
        assert shard.numel() == (hi - lo) * n
        sh.process_local(shard)
        out = sh.gather(shard, batch, root=0)
        torch.cuda.synchronize()
        if rank == 0:
            got = out.cpu().numpy()
            single = torch.from_numpy(x).to(dev)
            sh.fft.process_device(single)
            assert np.array_equal(got, single.cpu().numpy()), n  # same kernels on every GPU: bit-identical
            assert rel_l2(got, truth(x, n, False)) <= strict_bound(n, np.complex64), n
            b = batch - 1
            assert rel_l2(got[b * n:(b + 1) * n], oracle.fft(x[b * n:(b + 1) * n], n)) <= 2 * strict_bound(n, np.complex64), n
        dist.barrier()
    if rank == 0:
        print("SHARDED-OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
