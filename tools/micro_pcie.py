#!/usr/bin/env python
"""PCIe ceiling of the box: pinned H2D alone, D2H alone, both at once (torch copies, CUDA-event timed)."""
import torch
n = 512 * 1024 * 1024
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n // 4, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n // 4, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, pieces=16, reps=5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_event(e0); s2.wait_event(e0)
    for _ in range(reps):
        for k in range(pieces):
            if h2d:
                with torch.cuda.stream(s1):
                    a = n // pieces; d_in[k * a:(k + 1) * a].copy_(h_in[k * a:(k + 1) * a], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    a = n // 4 // pieces; h_out[k * a:(k + 1) * a].copy_(d_out[k * a:(k + 1) * a], non_blocking=True)
    ev1, ev2 = torch.cuda.Event(), torch.cuda.Event()
    ev1.record(s1); ev2.record(s2)
    torch.cuda.current_stream().wait_event(ev1); torch.cuda.current_stream().wait_event(ev2)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return (n * reps / ms / 1e6 if h2d else 0.0), (n // 4 * reps / ms / 1e6 if d2h else 0.0)
run(True, True, reps=1)
for name, a, b in (("h2d only", True, False), ("d2h only", False, True), ("both (d2h = h2d / 4)", True, True)):
    for pieces in (1, 16, 64):
        g = run(a, b, pieces)
        print(f"{name:24s} pieces {pieces:3d}: h2d {g[0]:6.1f} GB/s  d2h {g[1]:6.1f} GB/s")
