"""How long does a dependent kernel node of a CUDA graph take to start on this GPU?  (context for the
5-8 us gaps between neuron(k) and neuron(k+1) in tools/timeline.py)"""
import torch, time
x = torch.zeros(1024, device="cuda")
big = torch.zeros(250_000 * 4, device="cuda")
for name, t, n in (("tiny (1 block)", x, 200), ("1M-element (977 blocks)", big, 200)):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            t.add_(1.0)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                t.add_(1.0)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) * 1e3 / (10 * n):.2f} us per dependent kernel node")
    # the same chain as plain stream launches
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10 * n):
        t.add_(1.0)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) * 1e3 / (10 * n):.2f} us per plain stream launch")

# ---- the window graph's shape with trivial kernels: chain(k) -> chain(k+1) on s1; side(k) on s2 after
# chain(k); optionally chain(k+1) also waits for side(k-1) (fan-in 2), as neuron(k+1) waits for synapse(k-1)
def shape(fan_in, n=200, two_sides=False):
    s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    a, b, c = (torch.zeros(1024, device="cuda") for _ in range(3))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s1):
        a.add_(1.0); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s1):
            side_ev = []
            for k in range(n):
                if fan_in and k >= 2:
                    s1.wait_event(side_ev[k - 2])
                a.add_(1.0)
                ev = torch.cuda.Event(); ev.record(s1)
                s2.wait_event(ev)
                with torch.cuda.stream(s2):
                    b.add_(1.0)
                    e2 = torch.cuda.Event(); e2.record(s2)
                side_ev.append(e2)
                if two_sides:
                    s3.wait_event(ev)
                    with torch.cuda.stream(s3):
                        c.add_(1.0)
            s1.wait_stream(s2)
            if two_sides:
                s1.wait_stream(s3)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (10 * n)

print(f"chain + one side branch per step (fan-out 2): {shape(False):.2f} us per step")
print(f"  ... and chain(k+1) waits for side(k-1) (fan-in 2): {shape(True):.2f} us per step")
print(f"  ... with two side branches (fan-out 3, fan-in 2): {shape(True, two_sides=True):.2f} us per step")
