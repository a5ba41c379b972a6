#!/usr/bin/env python
"""Kernel timeline of one captured window (no nsys in the image): every kernel stamps
%globaltimer at its first block's start and last block's end (snn_config.reserved[1] & 2).

  python tools/timeline.py [--workload c5_1M_100M] [--window-ms 100] [--out gpurun_out/timeline.json]
"""
import argparse, ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from rl_stdp_b200 import _lib as L, synth
from rl_stdp_b200.network import SparseNetwork

CLS = ["neuron", "synapse", "ltp", "learn", "remote/replay", "push", "n.wait", "n.cur", "n.euler", "n.spk", "n.fence", "s.A", "s.B", "l.A", "l.B", "-"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c5_1M_100M")
    ap.add_argument("--window-ms", type=int, default=100)
    ap.add_argument("--warm", type=int, default=4)
    ap.add_argument("--exp-flags", type=int, default=0)
    ap.add_argument("--learn-blocks", type=int, default=0)
    ap.add_argument("--learn-async", default="auto")
    ap.add_argument("--learn-lazy", type=int, default=0, help="experimental event-driven learning (DESIGN 11.1)")
    ap.add_argument("--chain", default=None)
    ap.add_argument("--chain-lookahead", type=int, default=0)
    ap.add_argument("--scheduler", default="auto", choices=["auto", "chain", "graph"])
    ap.add_argument("--waves", action="store_true")
    ap.add_argument("--solo", action="store_true", help="chain phases (classes 6-12) of ONE neuron block instead of the envelope over all")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline.json"))
    a = ap.parse_args()
    if a.solo:
        a.exp_flags |= 1024
    wl = bench.WORKLOADS[a.workload]
    n, ne, fan, dmax = wl[:4]
    n_inst = wl[4] if len(wl) > 4 else 1
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    rowptr, post, w, delay = synth.fixed_fanout(n, ne, fan, dmax, seed=9, n_instances=n_inst)
    if world > 1:  # torchrun: ONE network neuron-partitioned over the ranks, in-library exchange
        import torch, torch.distributed as dist
        from rl_stdp_b200.partition import PartitionedRank, connect_dist, plan
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        lo, hi = plan(n, world)[rank]
        net = PartitionedRank(n, ne, rowptr, post, w, delay, lo, hi, None, dtype="f32", mode="fast", noise="philox",
                              device=local, max_delay=dmax, noise_seed=10, exp_flags=2 | a.exp_flags,
                              learn_async=a.learn_async)
        connect_dist(net)
    else:
        net = SparseNetwork(n, ne, rowptr, post, w, delay, n_instances=n_inst, dtype="f32", mode="fast", noise="philox",
                        max_delay=dmax, noise_seed=10, exp_flags=2 | a.exp_flags, learn_blocks=a.learn_blocks,
                            learn_async=a.learn_async, learn_lazy=a.learn_lazy, chain=a.chain, persistent={'auto': None, 'chain': True, 'graph': False}[a.scheduler], chain_lookahead=a.chain_lookahead, wave_kernels=a.waves)
    W = a.window_ms
    for kw in range(a.warm):
        tr, da = bench.window_inputs(kw, W)
        _, st = net.run(W, None, tr, da_events=da, record=False)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        if rank != 0:
            # every rank keeps its raw stamps (absolute %globaltimer ns: the GPUs of a box share the time base closely
            # enough to line the ranks up to a few us) next to rank 0's report
            S, K = L.SNN_TL_STEPS, L.SNN_TL_CLASSES
            buf = np.zeros(2 * K * S, np.uint64)
            L.check(L.load().snn_debug_timeline(net._h, buf.ctypes.data_as(C.c_void_p), buf.size))
            np.save(a.out.replace(".json", f"_rank{rank}.npy"), buf.reshape(2, K, S))
            dist.destroy_process_group()
            return
    S, K = L.SNN_TL_STEPS, L.SNN_TL_CLASSES
    buf = np.zeros(2 * K * S, np.uint64)
    L.check(L.load().snn_debug_timeline(net._h, buf.ctypes.data_as(C.c_void_p), buf.size))
    tl = buf.reshape(2, K, S)
    if world > 1:
        np.save(a.out.replace(".json", "_rank0.npy"), tl)
    valid = tl[1] > 0
    valid[13:16, :] = False   # classes 13-15 carry per-SM diagnostics of the chain, not (begin, end) stamps
    t0 = int(tl[0][valid].min())
    rows = []
    for c in range(K):
        for k in range(S):
            if valid[c, k]:
                rows.append({"cls": CLS[c], "k": k - 1, "start_us": (int(tl[0, c, k]) - t0) / 1e3,
                             "end_us": (int(tl[1, c, k]) - t0) / 1e3})
    # persistent chain: blocks of the synapse / learn role per SM and the neuron block's duration of two probe steps
    if int(tl[1, 13].sum()) > 0:
        syn, lrn = tl[1, 13].astype(np.int64), tl[1, 14].astype(np.int64)
        da, db = tl[0, 15].astype(np.float64), tl[1, 15].astype(np.float64)
        da[tl[0, 15] == np.uint64(0xFFFFFFFFFFFFFFFF)] = np.nan
        sms = [i for i in range(S) if syn[i] or lrn[i] or not np.isnan(da[i])]
        print("per SM: synapse blocks, learn blocks, neuron-block time of step 27 / 33 (us)")
        import collections
        grp = collections.defaultdict(list)
        for i in sms:
            grp[(int(syn[i]), int(lrn[i]))].append((da[i] / 1e3, db[i] / 1e3))
        for key in sorted(grp):
            v = np.array(grp[key])
            print(f"  syn={key[0]} learn={key[1]}: {len(v):3d} SMs  step27 mean {np.nanmean(v[:,0]):6.1f} max {np.nanmax(v[:,0]):6.1f}   step33 mean {np.nanmean(v[:,1]):6.1f} max {np.nanmax(v[:,1]):6.1f}")
        valid[13:16, :] = False
    rows.sort(key=lambda r: r["start_us"])
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump({"workload": a.workload, "window_ms": W, "device_ms": st.device_ms, "rows": rows}, open(a.out, "w"))
    # summary: per class mean duration, neuron(k)->neuron(k+1) start-to-start
    for c in CLS:
        d = [r["end_us"] - r["start_us"] for r in rows if r["cls"] == c]
        if d:
            print(f"{c:8s} n={len(d):4d} mean={np.mean(d):8.2f} us  median={np.median(d):8.2f}  max={np.max(d):8.2f}")
    ns = {r["k"]: r for r in rows if r["cls"] == "neuron"}
    ks = sorted(ns)
    gaps = [(k, ns[k]["start_us"] - ns[k - 1]["end_us"], ns[k]["start_us"] - ns[k - 1]["start_us"]) for k in ks if k - 1 in ns]
    print("neuron end->next start gap (us) median", np.median([g[1] for g in gaps]), " start->start median", np.median([g[2] for g in gaps]))
    print("window device_ms", st.device_ms, " span_us", max(r["end_us"] for r in rows))
    # per learn period: pass duration vs the neuron chain beside it
    for lr in [r for r in rows if r["cls"] == "learn"][:4]:
        k0 = lr["k"]
        ch = [ns[k] for k in range(k0 + 1, k0 + 11) if k in ns]
        if ch:
            print(f"pass k={k0}: {lr['start_us']:.0f}->{lr['end_us']:.0f} ({lr['end_us'] - lr['start_us']:.0f} us); neuron chain "
                  f"{ch[0]['start_us']:.0f}->{ch[-1]['end_us']:.0f} ({ch[-1]['end_us'] - ch[0]['start_us']:.0f} us, "
                  f"{len(ch)} steps, mean kernel {np.mean([c['end_us'] - c['start_us'] for c in ch]):.1f} us)")
    print("first 40 rows:")
    for r in [r for r in rows if 26 <= r["k"] <= 31]:
        print(f"  {r['cls']:8s} k={r['k']:3d}  {r['start_us']:9.2f} -> {r['end_us']:9.2f}  ({r['end_us'] - r['start_us']:7.2f})")


if __name__ == "__main__":
    main()
