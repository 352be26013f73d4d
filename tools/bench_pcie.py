#!/usr/bin/env python
"""Raw host<->device copy rates of this box (pinned memory, 1 GiB), to put the e2e figure in context."""
import json, time, torch
n = 1 << 30
h_a = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
h2d = t(lambda: d_a.copy_(h_a, non_blocking=True))
d2h = t(lambda: h_b.copy_(d_b, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d_a.copy_(h_a, non_blocking=True)
    with torch.cuda.stream(s2): h_b.copy_(d_b, non_blocking=True)
bi = t(both)
print(json.dumps({"h2d_GBps": round(n / h2d / 1e9, 1), "d2h_GBps": round(n / d2h / 1e9, 1),
                  "simultaneous_each_direction_GBps": round(n / bi / 1e9, 1)}))
