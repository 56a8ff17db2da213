"""The N>1 path on CPU: two gloo ranks exchange their instance totals and derive global label
offsets; ids must be disjoint and dense across ranks."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from classpose_b200 import distributed as cdist
    rng = np.random.default_rng(100 + rank)
    n_tiles = 37
    a, b = cdist.shard_range(n_tiles, rank, world)
    counts = torch.from_numpy(rng.integers(0, 150, size=b - a).astype(np.int32))
    offs, total, base = cdist.global_label_offsets(counts)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.stack([offs.numpy(), counts.numpy().astype(np.int64)]))
    dist.destroy_process_group()


def test_global_label_offsets_two_ranks(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    offs, counts = [], []
    for r in range(world):
        a = np.load(tmp_path / f"r{r}.npy")
        offs.append(a[0]); counts.append(a[1])
    offs, counts = np.concatenate(offs), np.concatenate(counts)
    assert len(offs) == 37
    np.testing.assert_array_equal(offs, np.cumsum(counts) - counts)   # dense, disjoint, in tile order


def test_single_process_offsets_need_no_group():
    from classpose_b200 import distributed as cdist
    offs, total, base = cdist.global_label_offsets(torch.tensor([3, 0, 5], dtype=torch.int32))
    assert offs.tolist() == [0, 3, 3] and total == 8 and base == 0
