#!/usr/bin/env python
"""BASELINE config 5: AAD-heavy / GHASH-bound sweep on one GPU.
Message size in {64 B .. 64 MiB} x AAD/PT ratio in {0, 1/16, 1/4, 1, 4, 16}, AES-128 and
AES-256, about 1 GiB (PT+AAD) per point, device-resident inputs, CUDA events.
A single message goes through the stream API, everything else through ONE batch call
(lanes chosen by the library).  Prints one JSON object per point."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import aesgcm_b200


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--total", type=int, default=1 << 30)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rng = np.random.default_rng(4)
    eng = aesgcm_b200.GcmEngine(0)
    sizes = [64, 1 << 10, 16 << 10, 256 << 10, 4 << 20, 64 << 20]
    ratios = [(0, 1), (1, 16), (1, 4), (1, 1), (4, 1), (16, 1)]
    results = []
    for kb in (16, 32):
        eng.set_key(rng.integers(0, 256, kb, dtype=np.uint8).tobytes())
        for size in sizes:
            for num, den in ratios:
                alen = size * num // den
                per = size + alen
                n_msgs = max(1, args.total // per)
                stride = (size + 15) & ~15
                astride = (alen + 15) & ~15
                d_in = torch.randint(0, 256, (n_msgs * stride,), dtype=torch.uint8, device="cuda")
                d_out = torch.empty_like(d_in)
                d_aad = torch.randint(0, 256, (max(1, n_msgs * astride),), dtype=torch.uint8, device="cuda")
                d_iv = torch.randint(0, 256, (n_msgs * 12,), dtype=torch.uint8, device="cuda")
                d_tags = torch.zeros(16 * n_msgs, dtype=torch.uint8, device="cuda")
                # a single huge message: one stream call (the grid-wide kernel); everything else is ONE
                # batch call -- the library cuts few long messages (bulk AAD included) into per-CTA segments
                use_stream = n_msgs <= 1   # (round 1: <= 4; the balanced warp-unit layout now does better on 2-4 messages)

                def run():
                    if use_stream:
                        for m in range(n_msgs):
                            eng.stream_crypt_device(0, bytes(12), d_aad[m * astride:m * astride + alen] if alen else None,
                                                    d_in[m * stride:m * stride + size], d_out[m * stride:m * stride + size],
                                                    d_tags[16 * m:16 * m + 16])
                    else:
                        eng.batch_crypt_uniform_device(0, d_iv, d_aad if alen else None, alen, astride, d_in, d_out, size,
                                                       stride, d_tags, n_msgs=n_msgs)
                for _ in range(2):
                    run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.iters):
                    run()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.iters
                r = {"aes": kb * 8, "msg_bytes": size, "aad_bytes": alen, "n_msgs": n_msgs, "api": "stream" if use_stream else "batch",
                     "ms": round(ms, 4), "payload_GBps": round(n_msgs * size / ms / 1e6, 1),
                     "pt_plus_aad_GBps": round(n_msgs * per / ms / 1e6, 1)}
                print(json.dumps(r), flush=True)
                results.append(r)
                del d_in, d_out, d_aad
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)
    eng.close()


if __name__ == "__main__":
    main()
