"""World-size-2 (gloo, CPU) coverage of the N > 1 path: contiguous sharding by transient, no
data-path collective, max-over-ranks timing reduction, ragged gather.  The per-rank compute
runs the real device code through the CPU thread emulator (tests/emu)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hiddenpose_b200 import sharding


def test_shard_bounds_cover_without_overlap():
    for total in (1, 3, 8, 64, 65):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [sharding.shard_bounds(64, w, 0)[1] for w in (2, 4, 8)] == [32, 16, 8]   # BASELINE configs[2]
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import lct_oracle as O
        from tests.emu.emu import EmuPlan
        M, N, B = 32, 8, 3                                   # ragged: rank 0 gets 2 transients, rank 1 gets 1
        x = torch.from_numpy(np.random.RandomState(5).rand(B, 1, M, N, N).astype(np.float32))
        mine = sharding.shard_batch(x, world, rank)
        plan = EmuPlan(N, M, 0.16)
        y_local, _, _ = plan.run(mine.numpy().reshape(-1, M, N, N), 1, M, [0] * mine.shape[0])
        y_local = torch.from_numpy(y_local).view(mine.shape[0], 1, M, N, N)
        y = sharding.gather_batch(y_local, B)                # test-only collective; the data path has none
        slowest = sharding.max_over_ranks(10.0 + rank)
        if rank == 0:
            yo = O.LctOracle(N, M, 0.16).forward(x, [0] * B, [M] * B)
            np.save(os.path.join(out_dir, "res.npy"), np.array([O.rel_l2(y, yo), slowest, y.shape[0]]))
    finally:
        dist.destroy_process_group()


def test_two_ranks_shard_by_transient(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    err, slowest, n = np.load(tmp_path / "res.npy")
    assert n == 3 and err <= 1e-5
    assert slowest == 11.0                                   # max over ranks, not rank 0's own value
