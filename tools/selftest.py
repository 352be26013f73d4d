#!/usr/bin/env python
"""Bulk regression in the style of `tb/gcm_testbench.py -n N -r` (tb/gcm_testbench.py:25-39):
N random test configurations (mode, enc/dec, test size, raw or pre-expanded key), each resolved
like tb/gcm_gctr.py does and replayed against the CUDA-backed model; results are compared with
OpenSSL (`cryptography`) when it is installed.

  python tools/selftest.py -n 50 [-e SEED] [-t short|medium]
"""
import argparse
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-n", "--ntest", type=int, default=20)
    ap.add_argument("-e", "--seed", type=int, default=1)
    ap.add_argument("-t", "--tsize", default="short", choices=["short", "medium"])
    args = ap.parse_args()
    from aesgcm_b200 import gcm_model, key_exp, stimulus as st
    try:
        from cryptography.hazmat.primitives.ciphers.aead import AESGCM
    except Exception:
        AESGCM = None
    max_n = {"short": 2 ** 12 - 1, "medium": 2 ** 16 - 1}[args.tsize]
    top = random.Random(args.seed)
    failures = 0
    t0 = time.time()
    for i in range(args.ntest):
        cfg = {"seed": top.randrange(1 << 30), "aes_mode": top.choice(["128", "192", "256"]), "key": "RANDOM", "iv": "RANDOM",
               "aad": top.choice(["RANDOM", "RANDOM", "EMPTY"]), "data": top.choice(["RANDOM", "RANDOM", "RANDOM", "EMPTY"]),
               "enc_dec": top.choice(["enc", "dec"]), "max_n_byte": max_n}
        pre = top.random() < 0.3
        r = st.replay(cfg, gcm_model.gcm, pre_expanded=pre, expand_key=key_exp.aes_expand_key)
        ok = True
        if AESGCM is not None:
            key = bytes.fromhex(r["data"]["key"]["data"])
            iv = bytes.fromhex(r["data"]["iv"]["data"])
            ref = AESGCM(key).encrypt(iv, b"".join(r["pt_words"]), b"".join(r["aad_words"]))
            ok = ref[:-16] == b"".join(r["ct_words"]) and ref[-16:] == r["tag"]
        if cfg["enc_dec"] == "dec":
            ok = ok and b"".join(r["dec_words"]) == b"".join(r["pt_words"]) and r["dec_tag"] == r["tag"]
        failures += 0 if ok else 1
        print("test %3d  seed %-10d mode %s %s %s  aad %5d B  data %5d B  %s" % (
            i, cfg["seed"], r["config"]["aes_mode"], cfg["enc_dec"], "pre-exp" if pre else "raw    ",
            r["data"]["aad_n_bytes"], r["data"]["pt_n_bytes"], "PASS" if ok else "FAIL"), flush=True)
    print(json.dumps({"tests": args.ntest, "failures": failures, "seconds": round(time.time() - t0, 1),
                      "checked_against_openssl": AESGCM is not None}))
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
