"""The N>1 path on CPU (gloo, world_size 2): contiguous image shards, gather of per-image
sizes, max/sum reductions -- the host-side logic bench.py uses under torchrun.  The codec
itself is stood in for by the oracle here (tests may use it); the GPU box runs the same
logic over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from felics_b200 import sharding


def test_shard_range_partitions_every_batch():
    for n in (0, 1, 2, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_images, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from oracle import felics_oracle as fo

        rng = np.random.default_rng(1234)                       # same batch on every rank
        batch = rng.integers(90, 140, (n_images, 24, 40), dtype=np.uint8)
        first, count = sharding.shard_range(n_images, rank, world)
        fels = [fo.compress(batch[i]) for i in range(first, first + count)]
        sizes = sharding.gather_sizes([len(f) for f in fels])
        offsets = sharding.global_offsets(sizes)
        t_max = sharding.reduce_scalar(float(rank + 1), "max")
        px_sum = sharding.reduce_scalar(float(count), "sum")
        if rank == 0:
            np.savez(out_path, offsets=offsets, t_max=t_max, px_sum=px_sum)
        # every rank can place its own images in the global stream
        mine = offsets[first:first + count + 1]
        assert all(int(mine[i + 1] - mine[i]) == len(fels[i]) for i in range(count))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [7, 8])
def test_two_ranks_gather_sizes(tmp_path, n_images):
    from oracle import felics_oracle as fo
    fo.lib()
    port = _free_port()
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, port, n_images, out), nprocs=2, join=True)
    got = np.load(out)
    rng = np.random.default_rng(1234)
    batch = rng.integers(90, 140, (n_images, 24, 40), dtype=np.uint8)
    want = sharding.global_offsets([len(fo.compress(img)) for img in batch])
    assert np.array_equal(got["offsets"], want)
    assert float(got["t_max"]) == 2.0 and float(got["px_sum"]) == float(n_images)
