#!/usr/bin/env python
"""Offset (ragged) batches: an IMIX-like mix of packet sizes and a heavy-tailed mix, 2^20 messages, AES-128 and AES-256,
through agcm_batch_crypt (lanes = 0) and agcm_batch_crypt_slots.  AGCM_NO_LEN_SORT=1 in the environment gives the arrival-order
assignment for comparison, AGCM_NO_LEN_CLASSES=1 the sorted order under one lane count, AGCM_RAGGED_LANES=G a fixed lane count (AGCM_SLOTS_LANES for the slots form; 2048 = the
row-gathering TMA kernel), RAGGED_SLOTS_ONLY=1 skips the packed form."""
import os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch, aesgcm_b200
eng = aesgcm_b200.GcmEngine(0)
rng = np.random.default_rng(11)
n = 1 << 20
mixes = {"IMIX (64 B 58 %, 576 B 33 %, 1500 B 9 %)": rng.choice([64, 576, 1500], n, p=[0.58, 0.33, 0.09]),
         "heavy tail (lognormal, median 400 B, max 64 KiB)": np.minimum(np.exp(rng.normal(6.0, 1.3, n)).astype(np.int64) + 1, 65536),
         "uniform 1 .. 3000 B": rng.integers(1, 3001, n)}
for kb in (16, 32):
    eng.set_key(bytes(range(kb)))
    for name, lens in mixes.items():
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        total = int(off[-1])
        d_in = torch.randint(0, 256, (total,), dtype=torch.uint8, device="cuda"); d_out = torch.empty_like(d_in)
        d_off = torch.from_numpy(off).cuda()
        d_iv = torch.randint(0, 256, (12 * n,), dtype=torch.uint8, device="cuda"); d_tags = torch.zeros(16 * n, dtype=torch.uint8, device="cuda")
        fn = lambda: eng.batch_crypt_device(0, d_iv, None, None, d_in, d_off, d_out, d_tags, avg_len_hint=int(lens.mean()),
                                            lanes=int(os.environ.get("AGCM_RAGGED_LANES", "0")))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if not os.environ.get("RAGGED_SLOTS_ONLY"):
            for _ in range(2): fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5): fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(json.dumps({"aes": kb * 8, "mix": name, "mean_len": round(float(lens.mean()), 1), "ms": round(ms, 4),
                              "payload_GBps": round(total / ms / 1e6, 1), "Mmsg_per_s": round(n / ms / 1e3, 1)}), flush=True)
        del d_in, d_out
        # the same lengths in fixed-pitch, 16-byte aligned slots (agcm_batch_crypt_slots)
        pitch = int((lens.max() + 15) // 16 * 16)
        if n * pitch <= (24 << 30):
            d_in = torch.randint(0, 256, (n * pitch,), dtype=torch.uint8, device="cuda"); d_out = torch.empty_like(d_in)
            d_len = torch.from_numpy(lens.astype(np.int32)).cuda()
            fn = lambda: eng.batch_crypt_slots_device(0, d_iv, None, None, 0, 0, d_in, d_out, d_len, pitch, d_tags, avg_len_hint=int(lens.mean()),
                                                      lanes=int(os.environ.get("AGCM_SLOTS_LANES", "0")))
            for _ in range(2): fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5): fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(json.dumps({"aes": kb * 8, "mix": name + " in %d B slots" % pitch, "mean_len": round(float(lens.mean()), 1), "ms": round(ms, 4),
                              "payload_GBps": round(total / ms / 1e6, 1), "Mmsg_per_s": round(n / ms / 1e3, 1)}), flush=True)
            del d_in, d_out
# a key per message as well (agcm_batch_crypt_perkey, offsets): the per-key kernel in length order
if not os.environ.get("RAGGED_SLOTS_ONLY"):
    for name, lens in mixes.items():
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        total = int(off[-1])
        d_in = torch.randint(0, 256, (total,), dtype=torch.uint8, device="cuda"); d_out = torch.empty_like(d_in)
        d_off = torch.from_numpy(off).cuda()
        d_keys = torch.randint(0, 256, (32 * n,), dtype=torch.uint8, device="cuda")
        d_iv = torch.randint(0, 256, (12 * n,), dtype=torch.uint8, device="cuda"); d_tags = torch.zeros(16 * n, dtype=torch.uint8, device="cuda")
        fn = lambda: eng.batch_crypt_perkey_device(256, 0, d_keys, d_iv, None, None, d_in, d_off, d_out, d_tags)
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(json.dumps({"aes": 256, "mix": name + ", a key per message", "mean_len": round(float(lens.mean()), 1), "ms": round(ms, 4),
                          "payload_GBps": round(total / ms / 1e6, 1), "Mmsg_per_s": round(n / ms / 1e3, 1)}), flush=True)
        del d_in, d_out
eng.close()
