"""CPU tests of the multi-GPU host logic (world_size 2, gloo): partition plan, synapse split by
post range, and the bitmap all-gather that the exchange callback performs."""
import os

import numpy as np
import pytest

from rl_stdp_b200 import synth
from rl_stdp_b200.partition import plan, split_by_post


def test_plan_alignment_and_cover():
    for n, world in [(4096, 2), (1_000_000, 8), (1000, 3), (128, 4)]:
        pl = plan(n, world)
        assert pl[0][0] == 0 and pl[-1][1] == n
        for (lo, hi), (lo2, _) in zip(pl, pl[1:]):
            assert hi == lo2
        for lo, hi in pl:
            assert lo % 128 == 0 and (hi % 128 == 0 or hi == n)
    eq = plan(1_048_576, 8)
    assert len({hi - lo for lo, hi in eq}) == 1  # equal shards when n % (128*world) == 0


def test_split_by_post_is_a_partition():
    n, ne = 1024, 800
    rowptr, post, w, delay = synth.fixed_fanout(n, ne, 30, 7, seed=3)
    seen = np.zeros(len(post), int)
    for lo, hi in plan(n, 4):
        rp, po, ww, dd, idx = split_by_post(rowptr, post, w, delay, lo, hi)
        assert rp[0] == 0 and rp[-1] == len(po) == len(idx) and len(rp) == n + 1
        assert np.all((po >= lo) & (po < hi))
        assert np.array_equal(po, post[idx]) and np.array_equal(ww, w[idx]) and np.array_equal(dd, delay[idx])
        pre = np.repeat(np.arange(n), np.diff(rp))
        pre_full = np.repeat(np.arange(n), np.diff(rowptr))
        assert np.array_equal(pre, pre_full[idx])  # rows keep their synapses, in order
        seen[idx] += 1
    assert np.all(seen == 1)


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rl_stdp_b200.partition import TorchDistExchange
    n = 128 * world * 3
    rng = np.random.Generator(np.random.Philox(key=5))
    truth = rng.integers(0, 2**31 - 1, size=(6, n // 32), dtype=np.int64).astype(np.int32)  # 6 steps of bitmaps
    lo, hi = plan(n, world)[rank]
    ex = TorchDistExchange()
    ok = True
    for step in range(6):
        local = torch.from_numpy(truth[step, lo // 32:hi // 32].copy())
        full = ex.gather_words(local)
        ok = ok and np.array_equal(full.numpy(), truth[step])
    # the bench protocol's reductions: max of times, sum of events
    vals = torch.tensor([10.0 + rank, 100.0 * (rank + 1)], dtype=torch.float64)
    mx, sm = vals.clone(), vals.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    ok = ok and mx[0].item() == 10.0 + world - 1 and sm[1].item() == 100.0 * world * (world + 1) / 2
    # the peer-exchange set-up (connect_dist): every rank must receive all blobs in rank order
    from rl_stdp_b200.partition import connect_dist
    got = {}
    connect_dist(object(), _export=lambda net: bytes([rank]) * 192,
                 _connect=lambda net, r, blobs: got.update(rank=r, blobs=blobs))
    ok = ok and got["rank"] == rank and got["blobs"] == [bytes([r]) * 192 for r in range(world)]
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_bitmap_allgather_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, 29533, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(r for r, _ in res) == [0, 1] and all(ok for _, ok in res), res
