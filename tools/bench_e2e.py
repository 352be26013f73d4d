#!/usr/bin/env python
"""Host-buffer (e2e) throughput of GcmEngine.encrypt for a few pipeline granules."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_in.random_(0, 256)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
for mb in sys.argv[1:] or ["4", "8", "16", "32"]:
    os.environ["AGCM_CHUNK_MB"] = mb
    import aesgcm_b200
    eng = aesgcm_b200.GcmEngine(0); eng.set_key(bytes(32))
    eng.encrypt(bytes(12), b"a" * 16, h_in, out=h_out)
    t0 = time.perf_counter()
    for _ in range(5):
        eng.encrypt(bytes(12), b"a" * 16, h_in, out=h_out)
    dt = (time.perf_counter() - t0) / 5
    print("chunk %s MiB: %.2f ms  %.1f GB/s" % (mb, dt * 1e3, n / dt / 1e9), flush=True)
    eng.close()
