#!/usr/bin/env python
"""Per-launch fixed cost of the stream kernel at the full persistent grid: ONE row of blocks (a block per lane: 148 x 512 x 16 B)
and a 64-row message through the four modes, back to back on one stream (CUDA events; the GPU side is longer than the host side)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, aesgcm_b200
eng = aesgcm_b200.GcmEngine(0); eng.set_key(bytes(range(32)))
iv = bytes(12)
d_aad = torch.zeros(16, dtype=torch.uint8, device="cuda"); d_tag = torch.zeros(16, dtype=torch.uint8, device="cuda"); d_p = torch.zeros(16, dtype=torch.uint8, device="cuda")
def t(fn, n=300):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n * 1e3, 2)
for rows in (1, 64, 1024):
    nb = 148 * 512 * 16 * rows
    d_in = torch.zeros(nb, dtype=torch.uint8, device="cuda"); d_out = torch.empty_like(d_in)
    print(json.dumps({"rows": rows, "bytes": nb, "us": {
        "gctr_only": t(lambda: eng.gctr_device(iv, 0, d_in, d_out)),
        "ghash_only": t(lambda: eng.ghash_device(d_in, d_tag)),
        "part (enc, no finish)": t(lambda: eng.stream_part_device(0, iv, 0, d_in, d_out, 0, d_p)),
        "part, blocks_after=12345": t(lambda: eng.stream_part_device(0, iv, 0, d_in, d_out, 12345, d_p)),
        "crypt, no aad": t(lambda: eng.stream_crypt_device(0, iv, None, d_in, d_out, d_tag)),
        "crypt, 16 B aad": t(lambda: eng.stream_crypt_device(0, iv, d_aad, d_in, d_out, d_tag)),
    }}), flush=True)
