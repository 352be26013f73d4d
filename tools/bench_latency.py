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
