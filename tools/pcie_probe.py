"""Host<->device copy bandwidth of this box with pinned buffers: H2D alone, D2H alone, both directions at once.
The floor for the `e2e` numbers of bench.py (which move 0.8-1.6 GB each way per step)."""
import json
import torch

n = 100_000_000   # doubles = 800 MB
h1 = torch.empty(n, dtype=torch.float64, pin_memory=True)
h2 = torch.empty(n, dtype=torch.float64, pin_memory=True)
d1 = torch.empty(n, dtype=torch.float64, device="cuda")
d2 = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def h2d():
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


def both():
    h2d()
    d2h()


gb = n * 8 / 1e9
out = {"bytes_each_way": n * 8}
for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", both)):
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    ms = timed(fn)
    out[name] = {"ms": ms, "GB/s_per_direction": gb / ms * 1e3}
print(json.dumps(out))
