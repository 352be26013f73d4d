#!/usr/bin/env python
"""Fixed cost per launch of the stream kernel: time enc+tag, GCTR-only and GHASH-only over 64 MiB ... 1 GiB
(device-resident, CUDA events, back-to-back launches) and fit ms = a + bytes / rate."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, aesgcm_b200
eng = aesgcm_b200.GcmEngine(0); eng.set_key(bytes(range(32)))
n_max = 1 << 30
d_in = torch.randint(0, 256, (n_max,), dtype=torch.uint8, device="cuda"); d_out = torch.empty_like(d_in)
d_tag = torch.zeros(16, dtype=torch.uint8, device="cuda"); d_aad = torch.zeros(16, dtype=torch.uint8, device="cuda")
iv = bytes(12)
def timeit(fn, iters):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
sizes = [1 << 26, 1 << 27, 1 << 28, 1 << 29, 1 << 30]
for name, fn in (("enc+tag", lambda n: eng.stream_crypt_device(0, iv, d_aad, d_in[:n], d_out[:n], d_tag)),
                 ("gctr only", lambda n: eng.gctr_device(iv, 0, d_in[:n], d_out[:n])),
                 ("ghash only", lambda n: eng.ghash_device(d_in[:n], d_tag)),
                 ("part (no finish)", lambda n: eng.stream_part_device(0, iv, 0, d_in[:n], d_out[:n], 0, d_tag))):
    ms = [timeit(lambda: fn(n), 40) for n in sizes]
    A = np.vstack([np.ones(len(sizes)), np.array(sizes, dtype=float)]).T
    a, b = np.linalg.lstsq(A, np.array(ms), rcond=None)[0]
    print("%-18s intercept %.1f us, asymptotic %.1f GB/s; ms: %s" % (name, a * 1e3, 1.0 / b / 1e6, " ".join("%.4f" % m for m in ms)), flush=True)
eng.close()
