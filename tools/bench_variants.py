#!/usr/bin/env python
"""Kernel-level throughput table on one GPU (device-resident inputs, CUDA events):
stream path for AES-128/192/256 (fused, CTR-only, GHASH-only) and the batched
packet path (BASELINE config 3) for each lanes-per-message setting.
  python tools/bench_variants.py [--only stream|packets] [--quick]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import aesgcm_b200


def timeit(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args()
    rng = np.random.default_rng(1)
    eng = aesgcm_b200.GcmEngine(0, threads=args.threads)
    iters = 3 if args.quick else 10
    res = []
    if args.only in ("", "stream"):
        n = 1 << 30
        d_in = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda")
        d_out = torch.empty_like(d_in)
        d_tag = torch.zeros(16, dtype=torch.uint8, device="cuda")
        d_ok = torch.zeros(1, dtype=torch.uint8, device="cuda")
        d_aad = torch.zeros(16, dtype=torch.uint8, device="cuda")
        iv = bytes(12)
        for kb in (16, 24, 32):
            eng.set_key(rng.integers(0, 256, kb, dtype=np.uint8).tobytes())
            for name, fn in (
                ("enc+tag", lambda: eng.stream_crypt_device(0, iv, d_aad, d_in, d_out, d_tag)),
                ("dec+verify", lambda: eng.stream_crypt_device(1, iv, d_aad, d_in, d_out, d_tag, d_ok)),
                ("gctr only", lambda: eng.gctr_device(iv, 0, d_in, d_out)),
                ("ghash only", lambda: eng.ghash_device(d_in, d_tag)),
            ):
                ms = timeit(fn, iters)
                r = {"path": "stream 2^30 B", "aes": kb * 8, "op": name, "ms": round(ms, 4), "GBps": round(n / ms / 1e6, 1)}
                print(json.dumps(r), flush=True)
                res.append(r)
        del d_in, d_out
    if args.only in ("", "packets"):
        n_msgs, length = 1 << 20, 1500
        for stride in (1504, 1500):
            d_buf = torch.randint(0, 256, (n_msgs * stride,), dtype=torch.uint8, device="cuda")
            d_out = torch.empty_like(d_buf)
            d_iv = torch.randint(0, 256, (n_msgs * 12,), dtype=torch.uint8, device="cuda")
            d_tags = torch.zeros(16 * n_msgs, dtype=torch.uint8, device="cuda")
            d_ok = torch.zeros(n_msgs, dtype=torch.uint8, device="cuda")
            key = rng.integers(0, 256, 24, dtype=np.uint8).tobytes()
            eng.set_key(key)
            eng.set_key(eng.round_keys())  # shared pre-expanded key
            for lanes in ((1, 2048) if args.quick else (0, 2048, 1, 2, 4, 8, 16, 32)):
                if lanes == 2048 and stride % 16:
                    continue   # the TMA-staged kernel needs a 16-byte pitch
                ms = timeit(lambda: eng.batch_crypt_uniform_device(0, d_iv, None, 0, 0, d_buf, d_out, length, stride, d_tags,
                                                                   n_msgs=n_msgs, lanes=lanes), iters)
                r = {"path": "packets 2^20 x 1500 B, stride %d" % stride, "aes": 192, "op": "enc+tag", "lanes": lanes,
                     "ms": round(ms, 4), "GBps": round(n_msgs * length / ms / 1e6, 1)}
                print(json.dumps(r), flush=True)
                res.append(r)
            del d_buf, d_out
    if args.only in ("", "keys"):
        # aes_kexp on the device: 2^20 distinct keys -> round keys (tb/key_exp.py:118 semantics), and set_key latency
        import time
        for kb in (16, 24, 32):
            n_keys = 1 << 20
            d_keys = torch.randint(0, 256, (n_keys * kb,), dtype=torch.uint8, device="cuda")
            d_rk = torch.empty((n_keys, 4 * kb + 112), dtype=torch.uint8, device="cuda")
            ms = timeit(lambda: eng.expand_keys_device(kb * 8, d_keys, d_rk), iters)
            t0 = time.perf_counter()
            for i in range(20):
                eng.set_key(bytes((i + j) & 255 for j in range(kb)))   # a NEW key every time: a reload is free
            sk_us = (time.perf_counter() - t0) / 20 * 1e6
            r = {"path": "k_key_expand, 2^20 keys", "aes": kb * 8, "op": "key schedule", "ms": round(ms, 4),
                 "GBps": round(n_keys * (4 * kb + 112) / ms / 1e6, 1), "Mkeys_per_s": round(n_keys / ms / 1e3, 1),
                 "set_key_us": round(sk_us, 1)}
            print(json.dumps(r), flush=True)
    if args.only in ("", "perkey"):
        # BASELINE config 4: AES-256 decrypt+verify, distinct key per message, 64 B AAD
        alen = 64
        for length, stride in (((1500, 1504),) if args.quick else ((64, 64), (256, 256), (1500, 1500), (1500, 1504), (4096, 4096))):
            n_msgs = (1 << 20) if length <= 1500 else (1 << 18)
            d_keys = torch.randint(0, 256, (n_msgs * 32,), dtype=torch.uint8, device="cuda")
            d_iv = torch.randint(0, 256, (n_msgs * 12,), dtype=torch.uint8, device="cuda")
            d_aad = torch.randint(0, 256, (n_msgs * alen,), dtype=torch.uint8, device="cuda")
            d_pt = torch.randint(0, 256, (n_msgs * stride,), dtype=torch.uint8, device="cuda")
            d_ct = torch.empty_like(d_pt)
            d_back = torch.empty_like(d_pt)
            d_tags = torch.zeros(16 * n_msgs, dtype=torch.uint8, device="cuda")
            d_ok = torch.zeros(n_msgs, dtype=torch.uint8, device="cuda")
            eng.batch_crypt_perkey_uniform_device(256, 0, d_keys, d_iv, d_aad, alen, alen, d_pt, d_ct, length, stride, d_tags,
                                                  n_msgs=n_msgs)
            ms = timeit(lambda: eng.batch_crypt_perkey_uniform_device(256, 1, d_keys, d_iv, d_aad, alen, alen, d_ct, d_back,
                                                                      length, stride, d_tags, d_ok, n_msgs=n_msgs), iters)
            assert int(d_ok.sum().item()) == n_msgs
            r = {"path": "per-message key, %d x %d B at a %d B pitch, 64 B AAD" % (n_msgs, length, stride), "aes": 256,
                 "op": "dec+verify", "ms": round(ms, 4), "GBps": round(n_msgs * length / ms / 1e6, 1),
                 "Mmsg_per_s": round(n_msgs / ms / 1e3, 2)}
            print(json.dumps(r), flush=True)
            del d_pt, d_ct, d_back
    eng.close()


if __name__ == "__main__":
    main()
