#!/usr/bin/env python
"""Probe: one AAD-heavy batch point of config 5 (16 KiB payload + 256 KiB AAD per message, ~1 GiB) through
agcm_batch_crypt_uniform, for an ncu capture of the lane-group kernel on GHASH-dominated work."""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch, aesgcm_b200
eng = aesgcm_b200.GcmEngine(0); eng.set_key(bytes(range(32)))
size, alen = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (16384, 262144)
n = (1 << 30) // (size + alen)
d_in = torch.randint(0, 256, (n * size,), dtype=torch.uint8, device="cuda"); d_out = torch.empty_like(d_in)
d_aad = torch.randint(0, 256, (max(1, n * alen),), dtype=torch.uint8, device="cuda")
d_iv = torch.randint(0, 256, (n * 12,), dtype=torch.uint8, device="cuda"); d_tags = torch.zeros(16 * n, dtype=torch.uint8, device="cuda")
for lanes in [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["0"])]:
    for _ in range(3):
        eng.batch_crypt_uniform_device(0, d_iv, d_aad if alen else None, alen, alen, d_in, d_out, size, size, d_tags, n_msgs=n, lanes=lanes)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.batch_crypt_uniform_device(0, d_iv, d_aad if alen else None, alen, alen, d_in, d_out, size, size, d_tags, n_msgs=n, lanes=lanes)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("size", size, "aad", alen, "lanes", lanes, "n", n, "ms", round(ms, 4), "total GB/s", round(n * (size + alen) / ms / 1e6, 1), flush=True)
