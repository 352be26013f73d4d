#!/usr/bin/env python
"""Stream and batch throughput when the caller's buffers are NOT 16-byte aligned (a payload behind a 5- or 13-byte record header)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, aesgcm_b200
eng = aesgcm_b200.GcmEngine(0); eng.set_key(bytes(range(32)))
iv = bytes(12); n = 1 << 28
buf_in = torch.randint(0, 256, (n + 64,), dtype=torch.uint8, device="cuda"); buf_out = torch.empty_like(buf_in)
d_tag = torch.zeros(16, dtype=torch.uint8, device="cuda")
def t(fn, it=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
for off_in, off_out in ((0, 0), (4, 4), (1, 1), (13, 13), (5, 0), (0, 5), (13, 4)):
    ms = t(lambda: eng.stream_crypt_device(0, iv, None, buf_in[off_in:off_in + n], buf_out[off_out:off_out + n], d_tag))
    print(json.dumps({"stream_bytes": n, "in_offset": off_in, "out_offset": off_out, "GBps": round(n / ms / 1e6, 1)}), flush=True)
# uniform batch of 1 MiB messages at an odd pitch (long messages: the warp-unit layout)
for length, stride in ((1 << 20, 1 << 20), (1 << 20, (1 << 20) + 4), ((1 << 20), (1 << 20) + 13), (16384, 16384), (16384, 16384 + 13)):
    nm = (1 << 28) // length
    d_in = torch.randint(0, 256, (nm * stride + 64,), dtype=torch.uint8, device="cuda"); d_out = torch.empty_like(d_in)
    d_iv = torch.randint(0, 256, (12 * nm,), dtype=torch.uint8, device="cuda"); d_tags = torch.zeros(16 * nm, dtype=torch.uint8, device="cuda")
    ms = t(lambda: eng.batch_crypt_uniform_device(0, d_iv, None, 0, 0, d_in, d_out, length, stride, d_tags, n_msgs=nm))
    print(json.dumps({"batch": "%d x %d B at a %d B pitch" % (nm, length, stride), "GBps": round(nm * length / ms / 1e6, 1)}), flush=True)
