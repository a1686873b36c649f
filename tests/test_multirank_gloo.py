"""N > 1 host logic on CPU: two gloo ranks shard the residues exactly as bench.py / dispatch do
(LPT over N_r, no data-path collective) and combine their timings with a MAX all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from basicrta_b200.plan import build_plan, shard_chains


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, sizes, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mine = shard_chains(sizes, world)[rank]
    plan = build_plan((np.asarray(sizes)[mine] + 3) // 4, 64, 3000)
    owned = torch.zeros(len(sizes), dtype=torch.int64)
    owned[torch.from_numpy(mine)] = 1
    dist.all_reduce(owned, op=dist.ReduceOp.SUM)
    load = torch.tensor([float(np.asarray(sizes)[mine].sum())])
    tmax = torch.tensor([float(rank + 1)])                    # stands for the per-rank elapsed ms
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    loads = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(loads, load)
    np.save(os.path.join(out_dir, f'r{rank}.npy'),
            np.array([owned.min().item(), owned.max().item(), tmax.item(), plan.est_efficiency] +
                     [x.item() for x in loads]))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_sharding(tmp_path):
    rng = np.random.default_rng(0)
    sizes = np.round(10 ** rng.uniform(4, 5, 40)).astype(np.int64).tolist()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, sizes, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        owned_min, owned_max, tmax, eff, l0, l1 = np.load(tmp_path / f'r{rank}.npy')
        assert owned_min == 1 and owned_max == 1              # every residue on exactly one rank
        assert tmax == 2.0                                    # max over ranks
        assert abs(l0 - l1) / (l0 + l1) < 0.05                # LPT balance
        assert eff > 0.5
