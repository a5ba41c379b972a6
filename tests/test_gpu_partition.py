"""Neuron-partitioned runs: the union of the ranks must reproduce the single network.
On a 1-GPU box the ranks are emulated as several handles of one process (threads + device-to-device
bitmap copies, rl_stdp_b200.partition.LocalExchange); with >= 2 GPUs the NCCL path is run too."""
import numpy as np
import pytest

from tests.helpers import rel_err, run_oracle_sparse

pytestmark = pytest.mark.gpu


def _case(n=4096, ne=3328, fan=60, dmax=9, t=300, seed=81):
    from rl_stdp_b200 import synth
    rowptr, post, w, delay = synth.fixed_fanout(n, ne, fan, dmax, seed=seed)
    noise = synth.thalamic_noise(t, n, seed=seed + 1)
    train = np.array([(k + 1) % 10 == 0 for k in range(t)], np.uint8)
    return n, ne, dmax, t, rowptr, post, w, delay, noise, train


@pytest.mark.parametrize("world", [2, 4])
def test_partition_exact_equals_oracle_single_gpu(world):
    from oracle import py_oracle as po
    from rl_stdp_b200 import _lib as L
    from rl_stdp_b200.partition import LocalExchange, PartitionedRank, plan, run_local_ranks
    n, ne, dmax, t, rowptr, post, w, delay, noise, train = _case()
    da = [(99, 0, 0.5), (200, 0, 0.25)]
    ora = po.Sparse(n, ne, rowptr, post, w, delay, max_delay=dmax)
    ref = run_oracle_sparse(ora, noise, train, da)
    ex = LocalExchange(world)
    nets = [PartitionedRank(n, ne, rowptr, post, w, delay, lo, hi, ex.callback(r), dtype="f64", mode="exact",
                            max_delay=dmax) for r, (lo, hi) in enumerate(plan(n, world))]
    outs = run_local_ranks(nets, t, noise, train, da_events=da)
    w_ref, sd_ref = ora.get(4), ora.get(5)
    plastic = np.repeat(np.arange(n), np.diff(rowptr)) < ne
    ev_total = 0
    for net, (spikes, stats) in zip(nets, outs):
        for k in range(t):  # every rank returns the full (gathered) spike record
            assert np.array_equal(spikes[k], ref[k]), (net.lo, k)
        idx = net.local_index
        assert rel_err(net.get(L.FIELD_W), w_ref[idx]) <= 1e-9
        pl = plastic[idx]
        assert rel_err(net.get(L.FIELD_SD)[pl], sd_ref[idx][pl]) <= 1e-9
        v, u = net.get(L.FIELD_V), net.get(L.FIELD_U)
        assert np.array_equal(v[net.lo:net.hi], ora.get(0)[net.lo:net.hi])
        assert np.array_equal(u[net.lo:net.hi], ora.get(1)[net.lo:net.hi])
        assert np.array_equal(net.get(L.FIELD_TRACE), ora.get(2))  # replicated on every rank
        assert net.get(L.FIELD_DA)[0] == ora.get(6)[0]
        ev_total += stats.n_syn_events
    assert ev_total == ora.counters()[0]  # deliveries are split between the ranks, none lost
    assert sum(len(r) for r in ref) > 1000


def test_partition_fast_mode_frozen_weights():
    """Production kernels (atomic pushes, delay-1 push of REMOTE spikers in remote_kernel) in a
    2-rank partition; weights frozen so the result is order independent."""
    from oracle import py_oracle as po
    from rl_stdp_b200 import _lib as L
    from rl_stdp_b200.partition import LocalExchange, PartitionedRank, plan, run_local_ranks
    n, ne, dmax, t, rowptr, post, w, delay, noise, train = _case(seed=91)
    ora = po.Sparse(n, ne, rowptr, post, w, delay, max_delay=dmax, learn_base=0.0)
    ref = run_oracle_sparse(ora, noise, train)
    ex = LocalExchange(2)
    nets = [PartitionedRank(n, ne, rowptr, post, w, delay, lo, hi, ex.callback(r), dtype="f64", mode="fast",
                            max_delay=dmax, learn_base=0.0) for r, (lo, hi) in enumerate(plan(n, 2))]
    outs = run_local_ranks(nets, t, noise, train)
    for net, (spikes, _) in zip(nets, outs):
        for k in range(t):
            assert np.array_equal(spikes[k], ref[k]), (net.lo, k)
        sd_o = ora.get(5)[net.local_index]
        assert np.abs(net.get(L.FIELD_SD) - sd_o).max() <= 1e-12 * max(1.0, np.abs(sd_o).max())
        assert np.array_equal(net.get(L.FIELD_V)[net.lo:net.hi], ora.get(0)[net.lo:net.hi])


def _p2p_worker(rank, world, port, q, mode="exact", same_device=False):
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = 0 if same_device else rank
    torch.cuda.set_device(dev)
    if same_device:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    from oracle import py_oracle as po
    from rl_stdp_b200 import _lib as L
    from rl_stdp_b200.partition import PartitionedRank, connect_dist, plan
    n, ne, dmax, t, rowptr, post, w, delay, noise, train = _case(seed=91 if mode == "fast" else 81)
    kw = dict(learn_base=0.0) if mode == "fast" else {}
    ora = po.Sparse(n, ne, rowptr, post, w, delay, max_delay=dmax, **kw)
    ref = run_oracle_sparse(ora, noise, train)
    lo, hi = plan(n, world)[rank]
    net = PartitionedRank(n, ne, rowptr, post, w, delay, lo, hi, None, dtype="f64", mode=mode,
                          max_delay=dmax, device=dev, **kw)
    connect_dist(net)
    spikes = []
    for lo_t in range(0, t, 100):  # three windows: the captured graph of the production path is reused
        sp, _ = net.run(100, noise[lo_t:lo_t + 100], train[lo_t:lo_t + 100])
        spikes += sp
    ok = all(np.array_equal(spikes[k], ref[k]) for k in range(t))
    if mode == "exact":
        ok = ok and rel_err(net.get(L.FIELD_W), ora.get(4)[net.local_index]) <= 1e-9
    sd_o = ora.get(5)[net.local_index]
    ok = ok and np.abs(net.get(L.FIELD_SD) - sd_o).max() <= 1e-9 * max(1.0, np.abs(sd_o).max())
    ok = ok and np.array_equal(net.get(L.FIELD_V)[net.lo:net.hi], ora.get(0)[net.lo:net.hi])
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def _run_workers(fn, world, port, *extra):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=fn, args=(r, world, port, q) + extra) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=420) for _ in procs]
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
    return res


@pytest.mark.parametrize("mode,port", [("exact", 29575), ("fast", 29577)])
def test_partition_p2p_ipc_two_processes_one_gpu(mode, port):
    """The in-library exchange end to end on ONE GPU: two processes (one rank each) map each other's
    bitmap ring and step counters through CUDA IPC; neuron_kernel stores its spike words into the
    peer's ring and raises the peer's counter, xwait_kernel waits on the device.  The two contexts
    time-slice the GPU, so a waiting rank is preempted until its peer has published."""
    res = _run_workers(_p2p_worker, 2, port, mode, True)
    assert all(ok for _, ok in res), res


def test_partition_p2p_two_gpus():
    """CUDA-IPC mapped peer rings over NVLink, one process per GPU."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    res = _run_workers(_p2p_worker, 2, 29573, "fast", False)
    assert all(ok for _, ok in res), res
    res = _run_workers(_p2p_worker, 2, 29579, "exact", False)
    assert all(ok for _, ok in res), res


def _nccl_worker(rank, world, port, q):
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oracle import py_oracle as po
    from rl_stdp_b200 import _lib as L
    from rl_stdp_b200.partition import PartitionedRank, TorchDistExchange, plan
    n, ne, dmax, t, rowptr, post, w, delay, noise, train = _case()
    ora = po.Sparse(n, ne, rowptr, post, w, delay, max_delay=dmax)
    ref = run_oracle_sparse(ora, noise, train)
    ex = TorchDistExchange()
    lo, hi = plan(n, world)[rank]
    net = PartitionedRank(n, ne, rowptr, post, w, delay, lo, hi, ex.callback(), dtype="f64", mode="exact",
                          max_delay=dmax, device=rank)
    spikes, _ = net.run(t, noise, train)
    ok = all(np.array_equal(spikes[k], ref[k]) for k in range(t))
    ok = ok and rel_err(net.get(L.FIELD_W), ora.get(4)[net.local_index]) <= 1e-9
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_partition_nccl_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, 29571, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
