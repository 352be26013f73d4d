#!/usr/bin/env python
"""Per-packet cost of the drop-in `gcm` model (aesgcm_b200.gcm_model) driven the way tb/gcm_test.py drives the reference
model: constructor, AAD blocks, one <=16 B block per load_plain_text call, get_tag.  Wall clock per packet."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aesgcm_b200.gcm_model import gcm
rng = np.random.default_rng(3)
def packet(n, alen, ed="enc", reps=50):
    key = {'data': rng.integers(0, 256, 32, dtype=np.uint8).tobytes().hex().upper(), 'n_bytes': 32}
    iv = {'data': rng.integers(0, 256, 12, dtype=np.uint8).tobytes().hex().upper(), 'n_bytes': 12}
    aad = rng.integers(0, 256, alen, dtype=np.uint8).tobytes(); pt = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    def once():
        m = gcm(key, iv, ed)
        for i in range(0, alen, 16): m.load_aad(aad[i:i + 16])
        for i in range(0, n, 16): m.load_plain_text(pt[i:i + 16])
        m.get_tag(bytes(16))   # (the model logs the mismatch with this dummy DUT tag, like the reference: run with 2>/dev/null)
        return m
    once()
    t0 = time.perf_counter()
    for _ in range(reps): once()
    return (time.perf_counter() - t0) / reps * 1e6
for n, alen in ((64, 16), (1500, 16), (4096, 64), (65536, 0)):
    print(json.dumps({"payload_bytes": n, "aad_bytes": alen, "us_per_packet": round(packet(n, alen), 1),
                      "us_per_16B_callback": round(packet(n, alen) / max(1, (n + 15) // 16), 2)}), flush=True)
