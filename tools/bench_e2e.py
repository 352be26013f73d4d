#!/usr/bin/env python
"""Host-buffer (e2e) throughput of GcmEngine.encrypt (agcm_stream_crypt_host, pinned buffers, copies inside the call)
over message sizes, for a few pipeline settings: "CHUNK_MB:RAMP_KB" pairs on the command line (RAMP_KB = 0: equal
granules, the round-1 pipeline).  Every setting is a fresh engine (the library reads the environment once per
context); the tag of every size is compared between the settings and, up to 64 MiB, with OpenSSL."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import aesgcm_b200

SIZES = [1 << 20, 4 << 20, 6 << 20, 8 << 20, 16 << 20, 32 << 20, (32 << 20) + 4097, 64 << 20, 256 << 20, 1 << 30]
n_max = max(SIZES)
h_in = torch.empty(n_max, dtype=torch.uint8, pin_memory=True); h_in.random_(0, 256)
h_out = torch.empty(n_max, dtype=torch.uint8, pin_memory=True)
key, iv, aad = bytes(range(32)), bytes(range(12)), b"a" * 16
try:
    from cryptography.hazmat.primitives.ciphers.aead import AESGCM
    ossl = AESGCM(key)
except Exception:
    ossl = None
tags = {}
for setting in sys.argv[1:] or ["32:0", "32:2048", "32:1024", "32:4096", "16:2048"]:
    mb, kb = setting.split(":")
    os.environ["AGCM_CHUNK_MB"], os.environ["AGCM_RAMP_KB"] = mb, kb
    eng = aesgcm_b200.GcmEngine(0); eng.set_key(key)
    row = {"chunk_MiB": int(mb), "ramp_KiB": int(kb), "GBps": {}}
    for n in SIZES:
        src, dst = h_in[:n], h_out[:n]
        _, tag = eng.encrypt(iv, aad, src, out=dst)
        if n not in tags:
            tags[n] = tag
            if ossl is not None and n <= (64 << 20):
                ref = ossl.encrypt(iv, src.numpy().tobytes(), aad)
                assert ref[-16:] == tag and ref[:64] == dst[:64].numpy().tobytes() and ref[n - 64:n] == dst[n - 64:n].numpy().tobytes(), n
        assert tags[n] == tag, (setting, n)
        reps = 20 if n <= (64 << 20) else 5
        t0 = time.perf_counter()
        for _ in range(reps):
            eng.encrypt(iv, aad, src, out=dst)
        dt = (time.perf_counter() - t0) / reps
        row["GBps"]["%.4g MiB" % (n / 2**20)] = round(n / dt / 1e9, 2)
    print(json.dumps(row), flush=True)
    eng.close()
