"""CPU (gloo, world_size 2) checks of the sharded-bounds plumbing in beast_tokenizer_b200/_dist.py:
MIN/MAX all-reduce of the column bounds and the uneven row gather in front of the exact quantile select."""
import os

import numpy as np
import torch


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from beast_tokenizer_b200 import _dist
    rng = np.random.default_rng(100 + rank)
    rows = torch.from_numpy(rng.normal(size=(5 + 4 * rank, 6)).astype(np.float32))      # uneven shards
    lo, hi = rows.min(0).values.clone(), rows.max(0).values.clone()
    _dist.allreduce_minmax(lo, hi)
    allrows = _dist.gather_rows(rows)
    local = _dist.gather_rows(rows, process_group=False)
    lo_b, hi_b = rows.min(0).values.clone(), rows.max(0).values.clone()
    _dist.allreduce_minmax(lo_b, hi_b, None, implicit=False)                              # per-batch default: local
    lo_w, hi_w = rows.min(0).values.clone(), rows.max(0).values.clone()
    _dist.allreduce_minmax(lo_w, hi_w, _dist.WORLD, implicit=False)
    empty = _dist.gather_rows(torch.zeros((0, 6)) if rank == 0 else rows)                 # a rank without rows
    q.put((rank, lo.numpy(), hi.numpy(), allrows.numpy(), local.shape[0], lo_b.numpy(), lo_w.numpy(), empty.shape[0]))
    dist.destroy_process_group()


def test_sharded_bounds_plumbing_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = [np.random.default_rng(100 + r).normal(size=(5 + 4 * r, 6)).astype(np.float32) for r in range(2)]
    full = np.concatenate(shards)
    for rank, lo, hi, allrows, n_local, lo_b, lo_w, n_empty in res:
        assert np.array_equal(lo, full.min(0)) and np.array_equal(hi, full.max(0))
        assert np.array_equal(allrows, full)                     # rank order, padding stripped
        assert n_local == shards[rank].shape[0]
        assert np.array_equal(lo_b, shards[rank].min(0))         # implicit=False keeps a bare call local
        assert np.array_equal(lo_w, full.min(0))
        assert n_empty == shards[1].shape[0]
    # the quantile of the gathered rows is the quantile of the whole data set, whatever the sharding
    assert np.array_equal(np.quantile(res[0][3], 0.01, axis=0), np.quantile(full, 0.01, axis=0))


def test_no_group_is_a_no_op():
    from beast_tokenizer_b200 import _dist
    x = torch.arange(12, dtype=torch.float32).reshape(4, 3)
    assert _dist.gather_rows(x) is x
    lo, hi = x.min(0).values, x.max(0).values
    assert _dist.allreduce_minmax(lo, hi) == (lo, hi)
    assert _dist.resolve() == (None, None) and _dist.resolve(False) == (None, None)
