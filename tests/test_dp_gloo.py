"""CPU, world_size 2 over gloo: the N>1 host logic (request sharding, sum-of-units / max-over-ranks timing, token gather)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from crane_b200 import dp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = dp.shard_indices(7, world, rank)
    thr, units, secs = dp.job_throughput(units_local=len(mine) * 256.0, seconds_local=1.0 + rank)
    toks = dp.gather_tokens(torch.tensor([10 * rank + 1, 10 * rank + 2], dtype=torch.int64))
    # config 4's batch sharding (crane_b200_decode_batch_gather): contiguous blocks, the rank-major gather restores the batch order
    import crane_b200
    block = crane_b200.shard_sequences(8, world, rank)
    order = dp.gather_tokens(torch.tensor(block, dtype=torch.int64))
    q.put((rank, mine, thr, units, secs, toks.tolist(), block, order.tolist()))
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29611, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5]
    assert res[0][6] == [0, 1, 2, 3] and res[1][6] == [4, 5, 6, 7]
    for r in res:
        assert r[7] == list(range(8))
    for _, _, thr, units, secs, toks, _, _ in res:
        assert units == 7 * 256.0 and secs == 2.0 and abs(thr - 7 * 256.0 / 2.0) < 1e-9   # sum units / MAX seconds
        assert toks == [1, 2, 11, 12]


def test_single_process_passthrough():
    assert dp.job_throughput(100.0, 4.0)[0] == 25.0
    assert dp.shard_indices(5, 1, 0) == [0, 1, 2, 3, 4]
    import crane_b200
    import pytest
    assert crane_b200.shard_sequences(32, 8, 3) == [12, 13, 14, 15]
    with pytest.raises(ValueError):
        crane_b200.shard_sequences(30, 8, 0)
