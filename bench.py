#!/usr/bin/env python
"""bench.py — headline benchmark of the RL-STDP hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json target): N = 1 000 000 Izhikevich neurons (800k exc / 200k
inh), fan-out 100 -> 100 000 000 synapses, exc delays U{1..20} ms, inh delay 1 ms, DA-STDP on
(learn! every 10th ms, dopamine +0.5 once per simulated second), fp32 production mode, thalamic
noise 13*(U-0.5) from on-device Philox.  One bench "step" = one snn_step() window of
--window-ms simulated milliseconds (default 250).

Metric: synaptic events/s (one spike delivered across one synapse), whole job.  Also reported:
simulated-ms per wall-s (x real time), per-kernel roofline fractions, and the CPU oracle port
timed on the box's host cores (a reported baseline, not the target).

N > 1 (torchrun, one rank per GPU): every rank simulates an independent instance of the workload
(instance sharding) — no data-path collective, "scaling": "weak".
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) off it
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

WORKLOADS = {
    # name: (neurons, n_exc, fanout, max_delay)
    "c5_1M_100M": (1_000_000, 800_000, 100, 20),
    "c3_10k_1M": (10_000, 8_000, 100, 1),
    "c1_single": (1_000, 800, 100, 1),   # BASELINE config 1 as the reference runs it: ONE 1000-neuron net
    # batched small nets: (neurons per instance, n_exc, fanout, max_delay, instances)
    "c1_x1024": (1_000, 800, 100, 1, 1024),
    "c3_x64": (10_000, 8_000, 100, 1, 64),
    "tiny": (20_000, 16_000, 100, 20),
    "c4_actors": None,  # 4096 x LIF [8,8,8] CartPole actors, whole episodes on the device (per GPU: 4096/N)
}
LEARN_PERIOD = int(os.environ.get("BENCH_LEARN_PERIOD", "10"))  # the reference learns every 10th ms (newnet_DASTDP.jl:48-51); the override is a diagnostic
DA_PERIOD = 1000
# algorithmic bytes (fp32), SURVEY §8(d)
B_NEURON = 36      # per neuron per ms
B_EVENT = 17 + 12  # per synaptic event: propagate 17 + LTD 12
B_LTP = 20         # per incoming plastic synapse of a spiker
B_LEARN = 16       # per plastic synapse per learn pass


def learn_traffic(workload):
    """DRAM bytes per learn_kernel launch from the committed `ncu --set full` capture (same workload)."""
    if workload != "c5_1M_100M":
        return None
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r1i_learn_pass_ncu_full.json")))
        v = [(float(x["dram__bytes_read.sum"]) + float(x["dram__bytes_write.sum"])) * 1e6 for x in d["launches"]]
        return sum(v) / len(v)
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, f"/tmp/bench_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                sm.append(float(f[0])); mx = float(f[1])
                for nme, val in zip(names, f[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nme)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}
        return out


def window_inputs(k_window: int, window_ms: int):
    """train flags + dopamine events of window k (absolute ms = k*window_ms + step)."""
    t0 = k_window * window_ms
    train = np.array([((t0 + s + 1) % LEARN_PERIOD) == 0 for s in range(window_ms)], np.uint8)
    da = [(s, 0, 0.5) for s in range(window_ms) if ((t0 + s + 1) % DA_PERIOD) == 0]
    return train, da


def run_b200(args):
    import torch
    import torch.distributed as dist
    from rl_stdp_b200 import _lib as L
    from rl_stdp_b200 import synth
    from rl_stdp_b200.network import SparseNetwork

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libsnn_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = WORKLOADS[args.workload]
    n, ne, fan, dmax = wl[:4]
    n_inst = wl[4] if len(wl) > 4 else 1
    W = args.window_ms
    t_build = time.perf_counter()
    if args.partition and world > 1:
        # ONE network, neuron-partitioned over the ranks, spike bitmap all-gathered per ms (NCCL)
        from rl_stdp_b200.partition import PartitionedRank, TorchDistExchange, connect_dist, plan
        rowptr, post, w, delay = synth.fixed_fanout(n, ne, fan, dmax, seed=9)
        lo, hi = plan(n, world)[rank]
        cb = TorchDistExchange().callback() if args.partition == "nccl" else None
        net = PartitionedRank(n, ne, rowptr, post, w, delay, lo, hi, cb, dtype="f32",
                              mode="fast", noise="philox", device=local, max_delay=dmax, noise_seed=10,
                              learn_async=args.learn_async)
        if cb is None:
            connect_dist(net)  # in-library exchange: peer stores over NVLink, device-side wait
        n_plastic = int(np.count_nonzero(np.repeat(np.arange(n), np.diff(rowptr))[net.local_index] < ne))
    else:
        rowptr, post, w, delay = synth.fixed_fanout(n, ne, fan, dmax, seed=9 + rank, n_instances=n_inst)
        net = SparseNetwork(n, ne, rowptr, post, w, delay, n_instances=n_inst, dtype="f32", mode="fast", noise="philox",
                            device=local, max_delay=dmax, noise_seed=10 + rank, exp_flags=args.exp_flags,
                            graph=not args.no_graph, learn_async=args.learn_async, learn_blocks=args.learn_blocks, learn_tile_k=args.learn_tile_k, learn_stages=args.learn_stages, side_blocks=args.side_blocks,
                            learn_lazy=args.learn_lazy, learn_lazy_da=args.learn_lazy_da)
        n_plastic = int(rowptr[ne]) * n_inst
    del post, w, delay
    t_build = time.perf_counter() - t_build
    lib = L.load()
    stream = torch.cuda.Stream()  # the library launches on this stream, so torch events bracket its kernels
    torch.cuda.set_stream(stream)
    lib.snn_set_stream(net._h, C.c_void_p(stream.cuda_stream))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also brings the network to its stationary firing regime)
    kw = 0
    # events around the learn pass (roofline of the dominant kernel) — not in partitioned runs, where the
    # pass is 1/N of the work and the extra event nodes were seen to cost 20 % at 8 ranks
    prof = False if (args.partition and world > 1) else 'learn'
    for _ in range(args.warmup):
        tr, da = window_inputs(kw, W)
        net.run(W, None, tr, da_events=da, record=False, profile=prof)  # same graph as the timed region
        kw += 1

    # ---- timed region 1: inputs resident, no spike record -> `value`
    sampler = ClockSampler(local)
    sampler.start()
    tot = dict(spikes=0, ev=0, ltp=0, learn_ms=0.0, neuron_ms=0.0, syn_ms=0.0, nl=0, nn=0, ns=0, npass=0)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        tr, da = window_inputs(kw, W)
        _, st = net.run(W, None, tr, da_events=da, record=False, profile=prof)
        kw += 1
        tot["spikes"] += st.n_spikes; tot["ev"] += st.n_syn_events; tot["ltp"] += st.n_ltp_events
        tot["learn_ms"] += st.learn_ms; tot["neuron_ms"] += st.neuron_ms; tot["syn_ms"] += st.synapse_ms
        tot["nl"] += st.learn_launches; tot["nn"] += st.neuron_launches; tot["ns"] += st.synapse_launches
        tot["npass"] += st.pass_launches
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()

    # ---- timed region 2: end to end through the C ABI with host buffers (spike record D2H)
    cap = int(max(1 << 20, 4 * tot["spikes"] / max(1, args.steps)))
    e2e_spikes = e2e_ev = 0
    h2d = d2h = 0
    for _ in range(min(2, args.warmup)):  # untimed: the recording variant of the window graph is captured once
        tr, da = window_inputs(kw, W)
        net.run(W, None, tr, da_events=da, record=True, spikes_cap=cap)
        kw += 1
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tr, da = window_inputs(kw, W)
        sp, st = net.run(W, None, tr, da_events=da, record=True, spikes_cap=cap)
        kw += 1
        e2e_spikes += st.n_spikes; e2e_ev += st.n_syn_events
        h2d += tr.nbytes + 16 * len(da)
        d2h += 4 * st.n_spikes + 4 * (W + 1) + 64
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- diagnostics (outside the timed regions): per-kernel-class CUDA-event times with plain
    # stream-ordered launches, to show where a simulated millisecond goes
    diag = dict(learn_ms=0.0, neuron_ms=0.0, syn_ms=0.0, nl=0, nn=0, ns=0, total=0.0)
    for _ in range(2):
        tr, da = window_inputs(kw, W)
        _, st = net.run(W, None, tr, da_events=da, record=False, profile='kernels')
        kw += 1
        diag["learn_ms"] += st.learn_ms; diag["neuron_ms"] += st.neuron_ms; diag["syn_ms"] += st.synapse_ms
        diag["nl"] += st.learn_launches; diag["nn"] += st.neuron_launches; diag["ns"] += st.synapse_launches
        diag["total"] += st.device_ms

    vals = torch.tensor([dev_ms, e2e_s * 1e3, float(tot["ev"]), float(e2e_ev), float(tot["spikes"])],
                        dtype=torch.float64, device="cuda")
    if world > 1:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms, e2e_ms = float(mx[0]), float(mx[1])
        ev_all, e2e_ev_all, spikes_all = float(sm[2]), float(sm[3]), float(sm[4])
    else:
        e2e_ms = e2e_s * 1e3
        ev_all, e2e_ev_all, spikes_all = float(tot["ev"]), float(e2e_ev), float(tot["spikes"])

    if rank == 0:
        peak, peak_src = peaks()
        sim_ms = args.steps * W
        value = ev_all / (dev_ms * 1e-3)
        learn_bytes = B_LEARN * n_plastic
        learn_avg_ms = tot["learn_ms"] / max(1, tot["nl"])
        learn_gbs = learn_bytes / (learn_avg_ms * 1e-3) / 1e9 if learn_avg_ms > 0 else 0.0
        path_bytes = (sim_ms * n * n_inst * B_NEURON + tot["ev"] * B_EVENT + tot["ltp"] * B_LTP + tot["nl"] * learn_bytes)
        line = {
            "metric": "synaptic events/s", "value": value, "unit": "events/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if (args.partition and world > 1) else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "neurons": n * n_inst, "synapses": int(rowptr[-1]), "max_delay_ms": dmax,
                       "instances_per_gpu": n_inst,
                       "window_ms_per_step": W, "learn_period_ms": LEARN_PERIOD,
                       **({"learn_rule": f"EXPERIMENTAL event-driven, re-basing every {args.learn_lazy} learn steps (approximates the clamp)"} if args.learn_lazy else {}), "noise": "on-device Philox4x32-10",
                       "instances": 1 if (args.partition and world > 1) else world,
                       "sharding": ("one network, neurons partitioned over the ranks; per ms the spike bitmap is exchanged "
                                    + ("by NCCL all-gather from a host callback" if args.partition == "nccl" else
                                       "by peer stores over NVLink + device-side wait (in-graph)"))
                                   if (args.partition and world > 1) else
                                   "one independent instance per GPU, no collective",
                       "l2": "working set ~4 GB per instance >> 126 MB L2 (no flush needed)"},
            "sim_ms_per_wall_s": sim_ms * n_inst * (1 if (args.partition and world > 1) else world) / (dev_ms * 1e-3),
            "x_realtime_per_instance": sim_ms / dev_ms,
            "firing_rate_hz": spikes_all / world / (n * n_inst) / (sim_ms * 1e-3),
            "e2e": {"value": e2e_ev_all / (e2e_ms * 1e-3), "unit": "events/s", "h2d_bytes_per_step": h2d // args.steps,
                    "d2h_bytes_per_step": d2h // args.steps,
                    "note": "snn_step() with host buffers: train flags + dopamine events in, sorted spike record out"},
            "gpu_launches": (tot["npass"] if tot["npass"] else tot["nl"]) + tot["nn"] + tot["ns"],
            "roofline": {"bound": "hbm",
                         "kernel": ("learn_pass_tma_kernel<float> (asynchronous pass, timed live beside the step kernels)"
                                    if tot["npass"] else "learn_kernel<float>"),
                         "achieved": learn_gbs, "peak": peak,
                         "unit": "GB/s", "frac": learn_gbs / peak, "traffic": learn_traffic(args.workload), "peak_source": peak_src,
                         "bytes_per_launch": learn_bytes, "avg_launch_ms": learn_avg_ms,
                         "share_of_step_time": tot["learn_ms"] / dev_ms if dev_ms else None},
            "roofline_path": {"algorithmic_bytes": path_bytes, "achieved": path_bytes / (dev_ms * 1e-3) / 1e9,
                              "unit": "GB/s", "frac": path_bytes / (dev_ms * 1e-3) / 1e9 / peak},
            "kernels_serial_diag": {"note": "2 extra windows, plain launches, CUDA events around each kernel class "
                                            "(includes launch gaps); not part of the timed region",
                                    "per_launch_us": {"neuron": 1e3 * diag["neuron_ms"] / max(1, diag["nn"]),
                                                      "arrival+ltp": 1e3 * diag["syn_ms"] / max(1, diag["ns"]),
                                                      "learn": 1e3 * diag["learn_ms"] / max(1, diag["nl"])},
                                    "window_ms": diag["total"] / 2},
            "learn_ms_in_timed_region": tot["learn_ms"],
            "clocks": clocks, "build_s": t_build,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = (cpu_baseline_dense_c1(args.cpu_seconds) if args.workload == "c1_single"
                                    else cpu_baseline(args.cpu_seconds))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_actors(args):
    """Config 4: 4096 independent CartPole LIF actors (cartpole_fitness.jl play_episode), sharded by
    instance over the ranks; a bench step = one full episode of every agent (<= 200 decisions x 40 ms)."""
    import torch
    import torch.distributed as dist
    from rl_stdp_b200.lif import ActorPopulation, borne
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    arch, total = [8, 8, 8], 4096
    A = total // world
    rng = np.random.Generator(np.random.Philox(key=4 + rank))
    pop = ActorPopulation(arch, A, device=local)
    for l in range(2):
        pop.set_weights(l, np.clip(rng.standard_normal((A, 8, 8)) * 0.3 + 0.8, 0, 1.1))
    matalpha = borne(rng.random((A, 8, 4)) * 20 - 10, -1.0, 1.0)

    def episode():
        s0 = rng.random((A, 4)) * 0.1 - 0.05
        ties = rng.integers(0, 2, size=(A, 200), dtype=np.int32)
        return pop.play_episodes(matalpha, s0, ties, n_steps=200), s0.nbytes + ties.nbytes + matalpha.nbytes

    for _ in range(args.warmup):
        episode()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    dev_ms = 0.0; sim_ms = 0; h2d = 0
    for _ in range(args.steps):
        out, nb = episode()
        dev_ms += out["device_ms"]; sim_ms += int(out["n_decisions"].sum()) * 40; h2d += nb
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    vals = torch.tensor([dev_ms, wall * 1e3, float(sim_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms, wall_ms, sim_all = float(mx[0]), float(mx[1]), float(sm[2])
    else:
        wall_ms, sim_all = wall * 1e3, float(sim_ms)
    if rank == 0:
        line = {"metric": "simulated agent-ms per wall-s", "value": sim_all / (dev_ms * 1e-3), "unit": "agent-ms/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "c4_actors", "agents": total, "arch": arch, "window_ms": 40, "max_decisions": 200,
                           "env": "CartPole-v1 restated, on device"},
                "e2e": {"value": sim_all / (wall_ms * 1e-3), "unit": "agent-ms/s", "h2d_bytes_per_step": h2d // args.steps,
                        "d2h_bytes_per_step": A * (8 + 4 + 4 + 32)},
                "gpu_launches": args.steps}
        if world == 1 and not args.no_cpu:
            from oracle import py_oracle as po
            o = po.Layered(arch, "lif")
            for l in range(2):
                o.e(l)[:, :] = np.clip(rng.standard_normal((8, 8)) * 0.3 + 0.8, 0, 1.1)
            t1 = time.perf_counter(); ms = 0
            while time.perf_counter() - t1 < min(args.cpu_seconds, 5.0):
                _, nd, _, _ = o.play_episode(matalpha[0], rng.random(4) * 0.1 - 0.05, rng.integers(0, 2, 200), 200)
                ms += nd * 40
            dt = time.perf_counter() - t1
            line["cpu_baseline"] = {"value": ms / dt, "unit": "agent-ms/s", "cores": 1, "kind": "port",
                                    "sample": f"one agent, sequential episodes for {dt:.1f} s (oracle/lif_oracle.c)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline_dense_c1(budget_s: float = 12.0):
    """Config 1 exactly as the reference computes it: the DENSE N x N algorithm of new_net.jl
    (oracle/snn_oracle.c ora_dense_*, fp64, 1 thread) on the full 1000-neuron net."""
    from oracle import py_oracle as po
    from rl_stdp_b200 import synth
    n, ne = 1000, 800
    conn = synth.dense_connection(n, 0.1, 1)
    ora = po.DenseA(n, ne, conn)
    fan = conn.sum(axis=1)
    rng = np.random.Generator(np.random.Philox(key=2))
    ev = steps = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        sp = ora.step(13.0 * (rng.random(n) - 0.5), (steps + 1) % LEARN_PERIOD == 0)
        ev += int(fan[sp].sum())
        steps += 1
    dt = time.perf_counter() - t0
    return {"value": ev / dt, "unit": "events/s", "cores": 1, "kind": "port",
            "sample": f"the full config-1 net (N=1000, dense N x N matrices as in new_net.jl), {steps} ms simulated in {dt:.1f} s",
            "sim_ms_per_wall_s": steps / dt, "host_cpus": os.cpu_count()}


def cpu_baseline(budget_s: float = 12.0, n=50_000, ne=40_000, fan=100, dmax=20):
    """The oracle port (oracle/snn_oracle.c sparse restatement, fp64, 1 thread — the reference is
    single-threaded Julia) on a 1/20-scale sample of the same workload.  Its cost is per synapse
    visited, so events/s carries over to the full size."""
    from oracle import py_oracle as po
    from rl_stdp_b200 import synth
    rowptr, post, w, delay = synth.fixed_fanout(n, ne, fan, dmax, seed=9)
    ora = po.Sparse(n, ne, rowptr, post, w, delay, max_delay=dmax)
    warm = 60
    for t in range(warm):
        ora.step(po.philox_noise(10, t, n), (t + 1) % LEARN_PERIOD == 0)
    ev0, _ = ora.counters()
    t0 = time.perf_counter()
    steps = 0
    while time.perf_counter() - t0 < budget_s:
        t = warm + steps
        noise = po.philox_noise(10, t, n)
        ora.step(noise, (t + 1) % LEARN_PERIOD == 0)
        steps += 1
    dt = time.perf_counter() - t0
    ev1, _ = ora.counters()
    return {"value": (ev1 - ev0) / dt, "unit": "events/s", "cores": 1, "kind": "port",
            "sample": f"N={n} neurons / {n * fan} synapses (1/20 of the workload, same fan-out, delays, DA-STDP), "
                      f"{steps} ms simulated in {dt:.1f} s",
            "sim_ms_per_wall_s": steps / dt, "host_cpus": os.cpu_count()}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Julia original cannot run
    here) on the host cores, same metric/unit/config, each step a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, ne, fan, dmax = WORKLOADS[args.workload if args.workload != "c4_actors" else "c5_1M_100M"][:4]
    # the reference itself is single-threaded Julia; this arm gives its algorithm every host core
    # (OpenMP build of the same oracle sources, bit-identical results)
    os.environ.setdefault("ORACLE_OMP", "1")
    from oracle import py_oracle as po
    from rl_stdp_b200 import synth
    ns = min(n, 50_000)
    nes = ns * ne // n
    rowptr, post, w, delay = synth.fixed_fanout(ns, nes, fan, dmax, seed=9)
    ora = po.Sparse(ns, nes, rowptr, post, w, delay, max_delay=dmax)
    ms_per_step = max(1, args.ref_ms_per_step)
    t_abs = 0

    def one_step():
        nonlocal t_abs
        for _ in range(ms_per_step):
            ora.step(po.philox_noise(10, t_abs, ns), (t_abs + 1) % LEARN_PERIOD == 0)
            if (t_abs + 1) % DA_PERIOD == 0:
                ora.add_da(0, 0.5)
            t_abs += 1

    for _ in range(args.warmup):
        one_step()
    ev0, _ = ora.counters()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step()
    dt = time.perf_counter() - t0
    ev1, _ = ora.counters()
    value = (ev1 - ev0) / dt
    cores = po.omp_threads()
    sample = (f"N={ns} neurons / {ns * fan} synapses sample of {args.workload}, {ms_per_step} simulated ms per step, "
              f"fp64, {cores} OpenMP thread(s) over the independent loops of the step (the Julia reference itself is "
              f"single-threaded)")
    line = {"impl": "reference", "metric": "synaptic events/s", "value": value, "unit": "events/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "neurons": n, "synapses": n * fan, "max_delay_ms": dmax,
                       "sampled_as": sample},
            "sim_ms_per_wall_s": args.steps * ms_per_step / dt,
            "cpu_baseline": {"value": value, "unit": "events/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5_1M_100M", choices=sorted(WORKLOADS))
    ap.add_argument("--window-ms", type=int, default=250, help="simulated ms per snn_step() call (one bench step)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-ms-per-step", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--exp-flags", type=int, default=0, help="snn_config.reserved[1] diagnostics (2 timeline, 4 serialise the learn pass, 512 LSU pass); 0 for reported numbers")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--learn-async", default="auto", choices=["auto", "off", "force"], help="learn pass beside the following steps (auto: large nets) or in place")
    ap.add_argument("--learn-lazy", type=int, default=0, help="EXPERIMENTAL event-driven learning: dense re-basing pass every M learn steps (0 = the reference's rule; not for reported numbers)")
    ap.add_argument("--learn-lazy-da", type=float, default=0.0, help="with --learn-lazy: re-base at every learn step while da exceeds this (0 = never)")
    ap.add_argument("--learn-tile-k", type=int, default=0, help="async learn pass: tile size in units of 1024 synapses (0 = 32)")
    ap.add_argument("--learn-stages", type=int, default=0, help="async learn pass: TMA pipeline depth (0 = 6)")
    ap.add_argument("--side-blocks", type=int, default=0, help="10*synapse + ltp kernel blocks per SM (0 = default)")
    ap.add_argument("--learn-blocks", type=int, default=0, help="async learn pass: blocks per SM (0 = default)")
    ap.add_argument("--partition", nargs="?", const="p2p", default=None, choices=["p2p", "nccl"],
                    help="N>1: ONE network neuron-partitioned over the ranks (strong scaling) instead of one instance "
                         "per rank; p2p = in-library exchange over peer memory, nccl = host callback + all-gather")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c4_actors":
        run_actors(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
