"""bench.py's contract on a machine without a GPU: the product arm fails loudly (no CPU fallback), the reference
arm prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


def test_product_arm_needs_a_gpu():
    if not _no_gpu():
        return  # on the GPU box this is what the gpu tests and the driver's bench run cover
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--no-cpu"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stdout + r.stderr)
    assert not any(l.startswith("{") for l in r.stdout.splitlines())   # and no number is printed


def test_reference_arm_json_line():
    # launched the way torchrun launches a worker: OMP_NUM_THREADS=1 in the environment.  The arm must still
    # use (and report) the cores it really has, on the workload's full network, with the product arm's config.
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                        "--warmup", "1", "--ref-ms-per-step", "3", "--ref-spinup-ms", "4", "--workload", "tiny"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "synaptic events/s" and d["unit"] == "events/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["sample"]
    try:
        have = len(os.sched_getaffinity(0))
    except AttributeError:
        have = os.cpu_count() or 1
    assert cb["cores"] == have, (cb["cores"], have)   # not the 1 that OMP_NUM_THREADS asked for
    assert f"{have} OpenMP thread" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # same config object as the product arm prints for this workload (the driver's same-config check)
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    a = argparse.Namespace(workload="tiny", window_ms=500, impl="b200", partition=None, learn_lazy=0)
    assert d["config"] == bench.bench_config(a, 1)
    assert d["config"]["neurons"] == 20000 and d["config"]["synapses"] == 2000000  # the full network, not a sample
