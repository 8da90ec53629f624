"""world_size-2 gloo test of the N>1 host logic (no GPU): contiguous op sharding, per-rank tally of
the shard (the oracle stands in for the device here — this tests the plumbing, not the kernels),
gather into disjoint ranges, max-over-ranks timing."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from bftkv_b200 import shard
    from oracle import c_oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)                      # same data on every rank
    M, R = 1001, 16
    op_off = (np.arange(M + 1) * R).astype(np.uint32)
    key_idx = np.tile(np.arange(R, dtype=np.uint64), M)
    status = rng.choice([0, 0, 0, 1, 6], M * R).astype(np.uint8)
    qcs = [(5, 16, 6, 11, list(range(16)))]
    lo, hi = shard.op_range(M, world, rank)
    off, t0, t1 = shard.slice_ops(op_off, lo, hi)
    mine = c_oracle.tally_batch(qcs, off, key_idx[t0:t1], status[t0:t1])
    full = torch.zeros(M, dtype=torch.uint8)
    full[lo:hi] = torch.from_numpy(mine)
    dist.all_reduce(full, op=dist.ReduceOp.SUM)          # disjoint ranges: sum == gather
    ms = shard.max_over_ranks(10.0 + rank, dist)
    if rank == 0:
        ref = c_oracle.tally_batch(qcs, op_off, key_idx, status)
        np.save(os.path.join(out_dir, "ok.npy"), np.array([int(np.array_equal(full.numpy(), ref)), int(ms == 10.0 + world - 1)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path, built):
    from bftkv_b200 import shard
    for M in (0, 1, 7, 65536, 1048576):
        for G in (1, 2, 4, 8):
            edges = [shard.op_range(M, G, r) for r in range(G)]
            assert edges[0][0] == 0 and edges[-1][1] == M and all(edges[i][1] == edges[i + 1][0] for i in range(G - 1))
            assert max(e[1] - e[0] for e in edges) - min(e[1] - e[0] for e in edges) <= 1
    mp.spawn(_worker, args=(2, 29500 + os.getpid() % 2000, str(tmp_path)), nprocs=2, join=True)
    ok = np.load(tmp_path / "ok.npy")
    assert ok.tolist() == [1, 1]
