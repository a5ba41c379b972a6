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
    "c2_teacher": None,  # BASELINE config 2: the Teacher loop (reward-driven dopamine, train = true), 256 learners per GPU
}
LEARN_PERIOD = int(os.environ.get("BENCH_LEARN_PERIOD", "10"))  # the reference learns every 10th ms (newnet_DASTDP.jl:48-51); the override is a diagnostic
DA_PERIOD = 1000
# algorithmic bytes (fp32), SURVEY §8(d)
B_NEURON = 36      # per neuron per ms
B_EVENT = 17 + 12  # per synaptic event: propagate 17 + LTD 12
B_LTP = 20         # per incoming plastic synapse of a spiker
B_LEARN = 16       # per plastic synapse per learn pass


def learn_traffic(workload):
    """DRAM bytes per learn-pass launch from the committed `ncu --set full` capture (same kernel, same workload):
    (bytes, file) or (None, None)."""
    if workload != "c5_1M_100M":
        return None, None
    try:
        name = "r2_step_kernels_ncu_full.json"
        rows = json.load(open(os.path.join(ROOT, "profiles", name)))
        v = [(float(x["dram__bytes_read.sum"].split()[0]) + float(x["dram__bytes_write.sum"].split()[0])) * 1e6
             for x in rows if "learn_pass_tma_kernel" in x["Kernel Name"]]
        if v:
            return sum(v) / len(v), name
    except Exception:
        pass
    try:
        name = "r1i_learn_pass_ncu_full.json"
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        v = [(float(x["dram__bytes_read.sum"]) + float(x["dram__bytes_write.sum"])) * 1e6 for x in d["launches"]]
        return sum(v) / len(v), name
    except Exception:
        return None, None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampled every 20 ms from before the warm-up; only the samples whose timestamp falls
    inside [mark_begin(), mark_end()] — the timed region — are reported (a sampler started at the timed
    region's first instant delivers its first line only after nvidia-smi has started up, which in a
    multi-rank run was after the region had ended)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, f"/tmp/bench_clocks_{os.getpid()}.csv"
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    @staticmethod
    def _ts(txt):
        import datetime
        try:
            return datetime.datetime.strptime(txt.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except Exception:
            return None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        if self.t1 is not None:  # let the sampler deliver the lines of the region's last instants
            time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows, mx = [], None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 8:
                    continue
                try:
                    rows.append((self._ts(f[0]), float(f[1]), [n for n, v in zip(names, f[4:8]) if v.lower().startswith("active")]))
                    mx = float(f[2])
                except ValueError:
                    continue
            os.remove(self.path)
        except Exception:
            pass
        inside = [r for r in rows if r[0] is not None and self.t0 is not None and self.t1 is not None
                  and self.t0 - 0.02 <= r[0] <= self.t1 + 0.02]
        scope = "timed region"
        if not inside and rows and self.t0 is not None:
            # a region shorter than the sampling period: the nearest samples on either side of it
            rows_t = [r for r in rows if r[0] is not None]
            rows_t.sort(key=lambda r: abs(r[0] - 0.5 * (self.t0 + (self.t1 or self.t0))))
            inside, scope = rows_t[:3], "nearest samples around the timed region"
        if inside:
            out = {"sm_mhz": statistics.median(r[1] for r in inside), "sm_max_mhz": mx,
                   "reasons": sorted({n for r in inside for n in r[2]}), "samples": len(inside), "scope": scope}
        return out


def window_inputs(k_window: int, window_ms: int):
    """train flags + dopamine events of window k (absolute ms = k*window_ms + step)."""
    t0 = k_window * window_ms
    train = np.array([((t0 + s + 1) % LEARN_PERIOD) == 0 for s in range(window_ms)], np.uint8)
    da = [(s, 0, 0.5) for s in range(window_ms) if ((t0 + s + 1) % DA_PERIOD) == 0]
    return train, da


def bench_config(args, world, partitioned=False):
    """`config` of the JSON line — ONE function for the product arm and the reference arm, so that the
    driver's same-config check compares like with like."""
    wl = WORKLOADS[args.workload]
    n, ne, fan, dmax = wl[:4]
    n_inst = wl[4] if len(wl) > 4 else 1
    cfg = {"workload": args.workload, "neurons": n * n_inst, "synapses": n * n_inst * fan, "max_delay_ms": dmax,
           "instances_per_gpu": n_inst, "window_ms_per_step": args.window_ms, "learn_period_ms": LEARN_PERIOD,
           "noise": "Philox4x32-10 keyed by (seed, step, neuron): the device generates it in the neuron kernel, the CPU arm replays the same stream",
           "topology": "clustered (20 blocks, 80 % of the targets inside the own block)" if args.workload.startswith("c3")
                       else "uniform fixed fan-out",
           "instances": 1 if partitioned else world,
           "sharding": ("one network, neurons partitioned over the ranks; per ms the spike bitmap is exchanged "
                        + ("by NCCL all-gather from a host callback" if args.partition == "nccl" else
                           "by peer stores over NVLink + device-side wait")) if partitioned
                       else "one independent instance per GPU, no collective",
           "l2": "working set ~4 GB per instance >> 126 MB L2 (no flush needed)"}
    if getattr(args, "learn_lazy", 0):
        cfg["learn_rule"] = f"EXPERIMENTAL event-driven, re-basing every {args.learn_lazy} learn steps (approximates the clamp)"
    return cfg


def synth_workload(args, seed, n_instances=None):
    from rl_stdp_b200 import synth
    wl = WORKLOADS[args.workload]
    n, ne, fan, dmax = wl[:4]
    n_inst = (wl[4] if len(wl) > 4 else 1) if n_instances is None else n_instances
    if args.workload.startswith("c3"):  # MatrixCluster: block-clustered topology
        parts = [synth.clustered(n, ne, fan, 20, 0.8, seed=seed + 1000 * b) for b in range(n_inst)]
        rowptr = np.arange(n * n_inst + 1, dtype=np.int64) * fan
        post = np.concatenate([p[1] + b * n for b, p in enumerate(parts)]).astype(np.int32)
        w = np.concatenate([p[2] for p in parts]); delay = np.concatenate([p[3] for p in parts])
        return rowptr, post, w, delay
    return synth.fixed_fanout(n, ne, fan, dmax, seed=seed, n_instances=n_inst)


def run_b200(args):
    import hashlib
    import torch
    import torch.distributed as dist
    from rl_stdp_b200 import _lib as L
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
    lib = L.load()
    stream = torch.cuda.Stream()  # the library launches on this stream, so torch events bracket its kernels
    torch.cuda.set_stream(stream)
    only_partition = bool(args.partition) and world > 1
    # N > 1: the headline stays instance sharding; the SAME invocation also times ONE network partitioned
    # over the ranks (BASELINE config 5) and checks it spike for spike against a 1-GPU run of that network
    with_partition_block = world > 1 and not only_partition and not args.no_partition_block and n_inst == 1
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    common = dict(dtype=args.dtype, mode="fast", noise="philox", device=local, max_delay=dmax, learn_async=args.learn_async,
                  chain=args.chain, persistent={'auto': None, 'chain': True, 'graph': False}[args.scheduler], chain_lookahead=args.chain_lookahead, wave_kernels=args.waves)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_partitioned():
        from rl_stdp_b200.partition import PartitionedRank, TorchDistExchange, connect_dist, plan
        rowptr, post, w, delay = synth_workload(args, seed=9)
        lo, hi = plan(n, world)[rank]
        cb = TorchDistExchange().callback() if args.partition == "nccl" else None
        pn = PartitionedRank(n, ne, rowptr, post, w, delay, lo, hi, cb, noise_seed=10, exp_flags=args.exp_flags, **common)
        if cb is None:
            connect_dist(pn)  # in-library exchange: peer stores over NVLink, device-side wait
        npl = int(np.count_nonzero(np.repeat(np.arange(n), np.diff(rowptr))[pn.local_index] < ne))
        lib.snn_set_stream(pn._h, C.c_void_p(stream.cuda_stream))
        return pn, npl

    t_build = time.perf_counter()
    pnet = None
    if only_partition:
        net, n_plastic = make_partitioned()
    else:
        rowptr, post, w, delay = synth_workload(args, seed=9 + rank)
        net = SparseNetwork(n, ne, rowptr, post, w, delay, n_instances=n_inst, noise_seed=10 + rank, exp_flags=args.exp_flags,
                            graph=not args.no_graph, learn_blocks=args.learn_blocks, learn_tile_k=args.learn_tile_k,
                            learn_stages=args.learn_stages, side_blocks=args.side_blocks, learn_lazy=args.learn_lazy,
                            learn_lazy_da=args.learn_lazy_da, **common)
        n_plastic = int(rowptr[ne]) * n_inst
        del post, w, delay
        lib.snn_set_stream(net._h, C.c_void_p(stream.cuda_stream))
    t_build = time.perf_counter() - t_build

    # ---- partition parity (N > 1): 100 ms with frozen weights from t = 0; every rank's spike record of the
    # partitioned network must equal, bit for bit, the record of the same network on ONE GPU (rank 0's instance)
    parity = None
    if with_partition_block:
        pnet, p_plastic = make_partitioned()
        PM = 100
        frozen = np.zeros(PM, np.uint8)
        sp_p, _ = pnet.run(PM, None, frozen, record=True, spikes_cap=1 << 24)
        digest = hashlib.sha1(b"".join(np.asarray(x, np.int32).tobytes() + b"|" for x in sp_p)).hexdigest()
        n_sp = int(sum(len(x) for x in sp_p))
        single = None
        if rank == 0:
            sp_1, _ = net.run(PM, None, frozen, record=True, spikes_cap=1 << 24)
            single = hashlib.sha1(b"".join(np.asarray(x, np.int32).tobytes() + b"|" for x in sp_1)).hexdigest()
        allp = [None] * world
        dist.all_gather_object(allp, (digest, n_sp))
        if rank == 0:
            parity = {"ms": PM, "weights": "frozen (train flags 0), fp32 production mode", "spikes": n_sp,
                      "ranks_agree": len({d for d, _ in allp}) == 1,
                      "identical_to_single_gpu": all(d == single for d, _ in allp)}

    def measure(net_, prof_, kw_, clock=None):
        tot_ = dict(spikes=0, ev=0, ltp=0, learn_ms=0.0, neuron_ms=0.0, syn_ms=0.0, nl=0, nn=0, ns=0, npass=0)
        for _ in range(args.warmup):  # also brings the network to its stationary firing regime
            tr, da = window_inputs(kw_, W)
            net_.run(W, None, tr, da_events=da, record=False, profile=prof_)  # same graph as the timed region
            kw_ += 1
        barrier()
        if clock is not None:
            clock.mark_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            tr, da = window_inputs(kw_, W)
            _, st = net_.run(W, None, tr, da_events=da, record=False, profile=prof_)
            kw_ += 1
            tot_["spikes"] += st.n_spikes; tot_["ev"] += st.n_syn_events; tot_["ltp"] += st.n_ltp_events
            tot_["learn_ms"] += st.learn_ms; tot_["neuron_ms"] += st.neuron_ms; tot_["syn_ms"] += st.synapse_ms
            tot_["nl"] += st.learn_launches; tot_["nn"] += st.neuron_launches; tot_["ns"] += st.synapse_launches
            tot_["npass"] += st.pass_launches
        e1.record()
        barrier()
        if clock is not None:
            clock.mark_end()
        return tot_, e0.elapsed_time(e1), kw_

    # ---- timed region 1: inputs resident, no spike record -> `value`
    # events around the learn pass (roofline of the dominant kernel) — not in partitioned runs, where the
    # pass is 1/N of the work and the extra event nodes were seen to cost 20 % at 8 ranks
    prof = False if only_partition else 'learn'
    tot, dev_ms, kw = measure(net, prof, 0, sampler if rank == 0 else None)
    clocks = sampler.stop() if rank == 0 else None

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
    Wd = min(W, 250)
    for _ in range(2):
        tr, da = window_inputs(kw, Wd)
        _, st = net.run(Wd, None, tr, da_events=da, record=False, profile='kernels')
        kw += 1
        diag["learn_ms"] += st.learn_ms; diag["neuron_ms"] += st.neuron_ms; diag["syn_ms"] += st.synapse_ms
        diag["nl"] += st.learn_launches; diag["nn"] += st.neuron_launches; diag["ns"] += st.synapse_launches
        diag["total"] += st.device_ms

    # ---- N > 1: the partitioned network, timed the same way (strong scaling of ONE network)
    part = None
    if with_partition_block:
        psampler = ClockSampler(local)
        if rank == 0:
            psampler.start()
        ptot, p_ms, _ = measure(pnet, False, 0, psampler if rank == 0 else None)
        pclocks = psampler.stop() if rank == 0 else None
        pv = torch.tensor([p_ms, float(ptot["ev"]), float(ptot["spikes"])], dtype=torch.float64, device="cuda")
        pmx = pv.clone(); dist.all_reduce(pmx, op=dist.ReduceOp.MAX)
        psm = pv.clone(); dist.all_reduce(psm, op=dist.ReduceOp.SUM)
        part = dict(ms=float(pmx[0]), ev=float(psm[1]), spikes=float(pmx[2]), clocks=pclocks)

    vals = torch.tensor([dev_ms, e2e_s * 1e3, float(tot["ev"]), float(e2e_ev), float(tot["spikes"])],
                        dtype=torch.float64, device="cuda")
    if world > 1:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms, e2e_ms = float(mx[0]), float(mx[1])
        ev_all, e2e_ev_all, spikes_all = float(sm[2]), float(sm[3]), float(sm[4])
        if only_partition:  # every rank's spike counter sees the whole network
            spikes_all = float(mx[4]) * world
    else:
        e2e_ms = e2e_s * 1e3
        ev_all, e2e_ev_all, spikes_all = float(tot["ev"]), float(e2e_ev), float(tot["spikes"])

    if rank == 0:
        peak, peak_src = peaks()
        sim_ms = args.steps * W
        value = ev_all / (dev_ms * 1e-3)
        fpw = 2 if args.dtype == "f64" else 1   # every floating-point quantity doubles; indices and bitmaps do not (path-level bytes below: upper bound for f64)
        learn_bytes = B_LEARN * fpw * n_plastic
        learn_avg_ms = tot["learn_ms"] / max(1, tot["nl"])
        learn_gbs = learn_bytes / (learn_avg_ms * 1e-3) / 1e9 if learn_avg_ms > 0 else 0.0
        path_bytes = (fpw * (sim_ms * n * n_inst * B_NEURON + tot["ev"] * B_EVENT + tot["ltp"] * B_LTP) + tot["nl"] * learn_bytes)
        traffic, traffic_file = learn_traffic(args.workload)
        if args.dtype != "f32":
            traffic = traffic_file = None   # the capture is of the fp32 kernel
        line = {
            "metric": "synaptic events/s", "value": value, "unit": "events/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if only_partition else "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": bench_config(args, world, only_partition),
            "sim_ms_per_wall_s": sim_ms * n_inst * (1 if only_partition else world) / (dev_ms * 1e-3),
            "x_realtime_per_instance": sim_ms / dev_ms,
            "simulated_s_in_timed_region": sim_ms * 1e-3,
            "firing_rate_hz": spikes_all / world / (n * n_inst) / (sim_ms * 1e-3),
            "e2e": {"value": e2e_ev_all / (e2e_ms * 1e-3), "unit": "events/s", "h2d_bytes_per_step": h2d // args.steps,
                    "d2h_bytes_per_step": d2h // args.steps,
                    "note": "snn_step() with host buffers: train flags + dopamine events in, sorted spike record out"},
            "gpu_launches": (tot["npass"] if tot["npass"] else tot["nl"]) + tot["nn"] + tot["ns"],
            "roofline": {"bound": "hbm",
                         "kernel": ("learn_pass_tma_kernel<%s> (asynchronous pass, timed live beside the step kernels)"
                                    if tot["npass"] else "learn_kernel<%s>") % ("double" if args.dtype == "f64" else "float"),
                         "achieved": learn_gbs, "peak": peak,
                         "unit": "GB/s", "frac": learn_gbs / peak, "traffic": traffic,
                         "traffic_source": (None if traffic is None else
                                            "dram__bytes_read+write per launch from the committed ncu --set full capture "
                                            "profiles/%s (same kernel, same workload, fp32), not measured in this run" % traffic_file),
                         "peak_source": peak_src,
                         "bytes_per_launch": learn_bytes, "avg_launch_ms": learn_avg_ms,
                         "note": ("timed live: the pass shares HBM with the step roles that run beside it for its whole "
                                  "duration (persistent chain); the same kernel alone streams the same bytes in 198 us "
                                  "= 0.99 of the measured peak (ncu, profiles/r2_step_kernels_ncu_full.json); "
                                  "roofline_path is the whole step path against the same peak"),
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
        if part is not None:
            x1 = sim_ms / dev_ms                 # the same network alone on one GPU, measured in this run (rank 0's instance)
            xp = sim_ms / part["ms"]
            line["partition"] = {
                "what": "ONE network of this workload, neurons partitioned over the ranks (BASELINE config 5), timed like the headline",
                "value": part["ev"] / (part["ms"] * 1e-3), "unit": "events/s", "ms_per_step": part["ms"] / args.steps,
                "x_realtime": xp, "x_realtime_one_gpu_same_run": x1, "strong_speedup_vs_n1": xp / x1,
                "strong_efficiency_vs_n1": xp / x1 / world,
                "exchange": "peer stores of (bitmap word, step) pairs over NVLink/NVSwitch + device-side wait, fused into the neuron kernels",
                "firing_rate_hz": part["spikes"] / n / (sim_ms * 1e-3), "clocks": part["clocks"]}
            line["partition_parity"] = parity
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = (cpu_baseline_dense_c1(args.cpu_seconds) if args.workload == "c1_single"
                                    else cpu_baseline(args.cpu_seconds))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_actors(args):
    """Configs 2 and 4, the LIF CartPole actor (Modules/Networks/net_arch.jl variant F), whole episodes on the device.

    c4_actors  4096 independent actors (cartpole_fitness.jl play_episode, train = false), sharded by instance over
               the ranks; a bench step = one full episode of every agent (<= 200 decisions x 40 ms).
    c2_teacher BASELINE config 2: the Teacher loop (teacher_sim.jl:61-80,101-144) — train = true, after every
               decision da += da_inc if the action matches the teacher's draw, -200 da_inc otherwise — run for 500
               episodes (--steps) on 256 independent learners per GPU (seeds); a bench step = one episode of each.
    The tie stream is the reference's default (`rng_seed = true`: every episode draws from a fresh
    MersenneTwister(0)), so it is one shared 200-entry stream; per step the host uploads the start states (and,
    c4: the population's observation matrices, which an evolution strategy rewrites every generation)."""
    import torch
    import torch.distributed as dist
    from rl_stdp_b200.lif import ActorPopulation, borne
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    teacher = args.workload == "c2_teacher"
    arch, total = [8, 8, 8], (256 * world if teacher else 4096)
    A = total // world
    rng = np.random.Generator(np.random.Philox(key=4 + rank))
    pop = ActorPopulation(arch, A, device=local, **(dict(da_inc=0.0005) if teacher else {}))
    for l in range(2):
        pop.set_weights(l, np.clip(rng.standard_normal((A, 8, 8)) * 0.3 + 0.8, 0, 1.1))
    matalpha = borne(rng.random((A, 8, 4)) * 20 - 10, -1.0, 1.0)

    def episode():
        s0 = rng.random((A, 4)) * 0.1 - 0.05
        if teacher:
            teach = rng.integers(0, 2, size=(A, 200), dtype=np.int32)   # the teacher's rand(0:1), teacher_sim.jl:132
            ties = rng.integers(0, 2, size=(A, 200), dtype=np.int32)
            out = pop.play_episodes(matalpha, s0, ties, n_steps=200, train=True, teach_actions=teach)
            return out, s0.nbytes + teach.nbytes + ties.nbytes + matalpha.nbytes
        return pop.play_episodes(matalpha, s0, None, n_steps=200), s0.nbytes + matalpha.nbytes + 4 * 200

    for _ in range(args.warmup):
        episode()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    dev_ms = 0.0; sim_ms = 0; h2d = 0; first = last = None
    for _ in range(args.steps):
        out, nb = episode()
        dev_ms += out["device_ms"]; sim_ms += int(out["n_decisions"].sum()) * 40; h2d += nb
        first = first if first is not None else float(out["rewards"].mean())
        last = float(out["rewards"].mean())
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
        # bound of actor_episode_kernel: one warp per agent, fp64, the whole episode in registers / shared memory —
        # neither HBM nor tensor cores; it is issue-latency bound (a dependent chain per LIF step), so what counts is
        # how many of the SMs' warp slots the population fills
        warp_slots = 148 * 64
        line = {"metric": "simulated agent-ms per wall-s", "value": sim_all / (dev_ms * 1e-3), "unit": "agent-ms/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
                "higher_is_better": True, "scaling": "weak" if teacher else "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": args.workload, "agents": total, "arch": arch, "window_ms": 40, "max_decisions": 200,
                           "env": "CartPole-v1 restated, on device",
                           **({"train": "teacher dopamine after every decision (teacher_sim.jl:132-133), learn! every ms"}
                              if teacher else {})},
                "e2e": {"value": sim_all / (wall_ms * 1e-3), "unit": "agent-ms/s", "h2d_bytes_per_step": h2d // args.steps,
                        "d2h_bytes_per_step": A * (8 + 4 + 4 + 32)},
                "gpu_launches": args.steps,
                "occupancy": {"agents_per_gpu": A, "warps": A, "warp_slots_per_gpu": warp_slots, "fill": min(1.0, A / warp_slots),
                              "bound": "issue latency of one warp per agent (no HBM traffic inside an episode)"},
                "mean_reward_first_last_step": [first, last]}
        if world == 1 and not args.no_cpu:
            from oracle import py_oracle as po
            o = po.Layered(arch, "lif")
            for l in range(2):
                o.e(l)[:, :] = np.clip(rng.standard_normal((8, 8)) * 0.3 + 0.8, 0, 1.1)
            t1 = time.perf_counter(); ms = 0
            while time.perf_counter() - t1 < min(args.cpu_seconds, 5.0):
                if teacher:
                    _, nd, _, _ = o.play_episode(matalpha[0], rng.random(4) * 0.1 - 0.05, rng.integers(0, 2, 200), 200,
                                                 train=True, teach_actions=rng.integers(0, 2, 200), da_inc=0.0005)
                else:
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
    """--impl reference: the reference's CPU algorithm (oracle port; the Julia original cannot run here:
    no Julia in the image) on every host core, on the SAME network as the product arm — the full
    workload, same topology seed, same Philox noise stream, same learn / dopamine schedule.  A step is a
    bounded sample: --ref-ms-per-step simulated ms of that network (a product-arm step is a whole
    window), after an untimed spin-up that takes the network out of its start-up transient so that
    events per simulated ms are those of the stationary regime."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload in ("c4_actors", "c2_teacher"):
        args.workload = "c5_1M_100M"
    wl = WORKLOADS[args.workload]
    n, ne, fan, dmax = wl[:4]
    n_inst = wl[4] if len(wl) > 4 else 1
    # the reference itself is single-threaded Julia; this arm gives its algorithm every host core
    # (OpenMP build of the same oracle sources, bit-identical results)
    os.environ.setdefault("ORACLE_OMP", "1")
    from oracle import py_oracle as po
    cores = po.use_all_cores()  # torchrun exports OMP_NUM_THREADS=1: ask for the cores, then ask what we got
    t_build = time.perf_counter()
    rowptr, post, w, delay = synth_workload(args, seed=9)
    ora = po.Sparse(n * n_inst, ne, rowptr, post, w, delay, max_delay=dmax, n_per_instance=n) if n_inst > 1 else \
        po.Sparse(n, ne, rowptr, post, w, delay, max_delay=dmax)
    del post, w, delay
    t_build = time.perf_counter() - t_build
    ms_per_step = max(1, args.ref_ms_per_step)
    t_abs = 0
    ntot = n * n_inst

    def one_ms():
        nonlocal t_abs
        sp = ora.step(po.philox_noise(10, t_abs, ntot), (t_abs + 1) % LEARN_PERIOD == 0)
        if (t_abs + 1) % DA_PERIOD == 0:
            for b in range(n_inst):
                ora.add_da(b, 0.5)
        t_abs += 1
        return len(sp)

    t_spin = time.perf_counter()
    for _ in range(max(0, args.ref_spinup_ms)):
        one_ms()
    t_spin = time.perf_counter() - t_spin
    for _ in range(args.warmup):
        for _ in range(ms_per_step):
            one_ms()
    ev0, _ = ora.counters()
    spikes = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for _ in range(ms_per_step):
            spikes += one_ms()
    dt = time.perf_counter() - t0
    ev1, _ = ora.counters()
    value = (ev1 - ev0) / dt
    sim_ms = args.steps * ms_per_step
    sample = (f"the full {args.workload} network ({ntot} neurons / {int(rowptr[-1])} synapses), {ms_per_step} simulated ms "
              f"per step ({sim_ms} ms timed) after {args.ref_spinup_ms} ms of untimed spin-up, fp64, {cores} OpenMP "
              f"thread(s) over the independent loops of the step (the Julia reference itself is single-threaded)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {"impl": "reference", "metric": "synaptic events/s", "value": value, "unit": "events/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": bench_config(args, world),
            "sampled_as": sample,
            "sim_ms_per_wall_s": sim_ms / dt,
            "firing_rate_hz": spikes / ntot / (sim_ms * 1e-3),
            "build_s": t_build, "spinup_s": t_spin,
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
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"], help="arithmetic type of the production path (f64: the fp64 build of the same kernels, non-contracted arithmetic; not the headline)")
    ap.add_argument("--window-ms", type=int, default=500, help="simulated ms per snn_step() call (one bench step)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-ms-per-step", type=int, default=4, help="--impl reference: simulated ms of the full network per step")
    ap.add_argument("--ref-spinup-ms", type=int, default=120, help="--impl reference: untimed simulated ms before the warm-up steps")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--exp-flags", type=int, default=0, help="snn_config.reserved[1] diagnostics (2 timeline, 4 serialise the learn pass, 512 LSU pass); 0 for reported numbers")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--chain", default=None, help="persistent chain footprints 'nthr,syn,ltp,ltp_thr,learn,neuron_regs,stages' (sweeps; default: the library's)")
    ap.add_argument("--chain-lookahead", type=int, default=0, help="persistent chain: steps of push lookahead (0 = default)")
    ap.add_argument("--waves", action="store_true", help="captured graph: the chain's wave kernels instead of the group-per-item side kernels")
    ap.add_argument("--scheduler", default="auto", choices=["auto", "chain", "graph"],
                    help="production-mode windows: the library's automatic choice, the persistent chain of resident role kernels, or the captured CUDA graph of per-step kernels")
    ap.add_argument("--learn-async", default="auto", choices=["auto", "off", "force"], help="learn pass beside the following steps (auto: large nets) or in place")
    ap.add_argument("--learn-lazy", type=int, default=0, help="EXPERIMENTAL event-driven learning: dense re-basing pass every M learn steps (0 = the reference's rule; not for reported numbers)")
    ap.add_argument("--learn-lazy-da", type=float, default=0.0, help="with --learn-lazy: re-base at every learn step while da exceeds this (0 = never)")
    ap.add_argument("--learn-tile-k", type=int, default=0, help="async learn pass: tile size in units of 1024 synapses (0 = 32)")
    ap.add_argument("--learn-stages", type=int, default=0, help="async learn pass: TMA pipeline depth (0 = 4)")
    ap.add_argument("--side-blocks", type=int, default=0, help="10*synapse + ltp kernel blocks per SM (0 = default)")
    ap.add_argument("--learn-blocks", type=int, default=0, help="async learn pass: blocks per SM (0 = default)")
    ap.add_argument("--no-partition-block", action="store_true",
                    help="N>1: skip the partitioned-network measurement that is otherwise added to the line")
    ap.add_argument("--partition", nargs="?", const="p2p", default=None, choices=["p2p", "nccl"],
                    help="N>1: ONE network neuron-partitioned over the ranks (strong scaling) instead of one instance "
                         "per rank; p2p = in-library exchange over peer memory, nccl = host callback + all-gather")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in ("c4_actors", "c2_teacher"):
        run_actors(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
