#!/usr/bin/env python
"""BASELINE config 3 through the HOST-buffer batch call (agcm_batch_crypt_uniform_host): records, IVs and tags in pinned host
memory, H2D and D2H inside the call.  Wall clock per call, GB/s of payload."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, aesgcm_b200
eng = aesgcm_b200.GcmEngine(0); eng.set_key(bytes(range(24)))
for nm, length, stride in ((1 << 20, 1500, 1504), (1 << 20, 1500, 1500), (1 << 18, 1500, 1504), (1 << 16, 16384, 16384)):
    data = torch.empty(nm * stride, dtype=torch.uint8, pin_memory=True); data.random_(0, 256)
    out = torch.empty(nm * stride, dtype=torch.uint8, pin_memory=True)
    ivs = torch.empty(12 * nm, dtype=torch.uint8, pin_memory=True); ivs.random_(0, 256)
    tags = torch.empty(16 * nm, dtype=torch.uint8, pin_memory=True)
    fn = lambda: eng.crypt_batch_uniform_host(0, ivs.numpy(), None, 0, 0, data.numpy(), out.numpy(), length, stride, tags.numpy())
    fn()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    dt = (time.perf_counter() - t0) / 5
    print(json.dumps({"n_msgs": nm, "len": length, "pitch": stride, "ms": round(dt * 1e3, 2), "payload_GBps": round(nm * length / dt / 1e9, 2),
                      "link_GBps_each_way": round(nm * (stride + 12 + 16) / dt / 1e9, 2)}), flush=True)
