#!/usr/bin/env python
"""Lanes-per-message sweep of the shared-key batch path (k_batch<G> / k_batch_cta) over message
sizes, about 1 GiB of payload per point, device-resident inputs, CUDA events.  The data behind
pick_lanes() in csrc/capi.cu.  Prints one JSON object per (aes, size, aad) row."""
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
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--sizes", default="64,256,1024,1500,4096,16384,65536,262144,4194304")
    ap.add_argument("--aes", default="256")
    ap.add_argument("--aad", default="0")
    args = ap.parse_args()
    rng = np.random.default_rng(5)
    eng = aesgcm_b200.GcmEngine(0)
    for bits in [int(x) for x in args.aes.split(",")]:
        eng.set_key(rng.integers(0, 256, bits // 8, dtype=np.uint8).tobytes())
        for alen in [int(x) for x in args.aad.split(",")]:
            for size in [int(x) for x in args.sizes.split(",")]:
                n_msgs = max(1, args.total // (size + alen))
                stride = (size + 15) & ~15
                astride = (alen + 15) & ~15
                d_in = torch.randint(0, 256, (n_msgs * stride,), dtype=torch.uint8, device="cuda")
                d_out = torch.empty_like(d_in)
                d_aad = torch.randint(0, 256, (max(1, n_msgs * astride),), dtype=torch.uint8, device="cuda")
                d_iv = torch.randint(0, 256, (n_msgs * 12,), dtype=torch.uint8, device="cuda")
                d_tags = torch.zeros(16 * n_msgs, dtype=torch.uint8, device="cuda")
                row = {"aes": bits, "msg_bytes": size, "aad_bytes": alen, "n_msgs": n_msgs}
                ref_tags = None
                for lanes in (0, 1, 2, 4, 8, 16, 32, 1024):
                    if lanes == 1024 and size < 8192:
                        continue
                    if lanes not in (0, 1024) and lanes > (size + 15) // 16 + 1:
                        continue

                    def run():
                        eng.batch_crypt_uniform_device(0, d_iv, d_aad if alen else None, alen, astride, d_in, d_out, size,
                                                       stride, d_tags, n_msgs=n_msgs, lanes=lanes)
                    run()
                    torch.cuda.synchronize()
                    t = d_tags.clone()
                    if ref_tags is None:
                        ref_tags = t
                    assert torch.equal(t, ref_tags), (size, lanes)   # every lane count gives the same tags
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(args.iters):
                        run()
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / args.iters
                    row["auto" if lanes == 0 else "g%d" % lanes] = round(n_msgs * size / ms / 1e6, 1)
                print(json.dumps(row), flush=True)
                del d_in, d_out, d_aad
    eng.close()


if __name__ == "__main__":
    main()
