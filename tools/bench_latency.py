#!/usr/bin/env python
"""Per-call latency of one small message (AES-256, 16 B AAD) through the device API
(agcm_stream_crypt, data in HBM) and the host API (agcm_stream_crypt_host)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import aesgcm_b200
eng = aesgcm_b200.GcmEngine(0); eng.set_key(bytes(range(32)))
iv, aad = bytes(12), bytes(16)
d_aad = torch.zeros(16, dtype=torch.uint8, device="cuda"); d_tag = torch.zeros(16, dtype=torch.uint8, device="cuda")
for n in (16, 1500, 65536, 1 << 20, 16 << 20):
    d_in = torch.zeros(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty_like(d_in)
    for _ in range(5): eng.stream_crypt_device(0, iv, d_aad, d_in, d_out, d_tag)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): eng.stream_crypt_device(0, iv, d_aad, d_in, d_out, d_tag)
    e1.record(); torch.cuda.synchronize()
    dev_us = e0.elapsed_time(e1) / 50 * 1e3
    h = np.zeros(n, np.uint8); o = np.zeros(n, np.uint8)
    for _ in range(3): eng.encrypt(iv, aad, h, out=o)
    t0 = time.perf_counter()
    for _ in range(20): eng.encrypt(iv, aad, h, out=o)
    host_us = (time.perf_counter() - t0) / 20 * 1e6
    print(json.dumps({"bytes": n, "device_api_us": round(dev_us, 1), "host_api_us": round(host_us, 1)}), flush=True)

# where the small-message floor goes: the same 16 B through the partial entry points
def _t(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n * 1e3, 1)
d_in = torch.zeros(16, dtype=torch.uint8, device="cuda"); d_out = torch.empty_like(d_in)
d_p = torch.zeros(16, dtype=torch.uint8, device="cuda")
empty = torch.zeros(1, device="cuda")
print(json.dumps({"floor_breakdown_us": {
    "torch_elementwise_kernel": _t(lambda: empty.add_(1)),
    "gctr_only_16B": _t(lambda: eng.gctr_device(iv, 0, d_in, d_out)),
    "ghash_only_16B": _t(lambda: eng.ghash_device(d_in, d_tag)),
    "stream_part_16B": _t(lambda: eng.stream_part_device(0, iv, 0, d_in, d_out, 0, d_p)),
    "stream_part_16B_blocks_after_1000": _t(lambda: eng.stream_part_device(0, iv, 0, d_in, d_out, 1000, d_p)),
    "stream_crypt_16B_no_aad": _t(lambda: eng.stream_crypt_device(0, iv, None, d_in, d_out, d_tag)),
    "stream_crypt_16B_aad16": _t(lambda: eng.stream_crypt_device(0, iv, d_aad, d_in, d_out, d_tag)),
}}), flush=True)
