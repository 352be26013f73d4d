#!/usr/bin/env python
"""Replay a reference test configuration (tb/tmp/<seed>.json, or the same flags as
tb/gcm_testbench.py -m/-k/-i/-a/-d/-b) against the B200 engine with no simulator:

  python tools/replay_config.py path/to/<seed>.json
  python tools/replay_config.py -m 128 -k AD7A2BD03EAC835A6F620FDCB506B345 -i 12153524C0895E81B2C28465 \\
      -a D609B1F056637A0D46DF998D88E52E00B2C2846512153524C0895E81 -d 08000F10...0002

Prints the ciphertext / plaintext words and the tag the scoreboard would have expected."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", nargs="?", help="tb/tmp/<seed>.json written by config/gcm_utils.py")
    ap.add_argument("-m", "--mode", default="128", choices=["128", "192", "256", "ALL"])
    ap.add_argument("-k", "--key", default="RANDOM")
    ap.add_argument("-i", "--iv", default="RANDOM")
    ap.add_argument("-a", "--aad", default="RANDOM")
    ap.add_argument("-d", "--data", default="RANDOM")
    ap.add_argument("-b", "--ed", default="enc", choices=["enc", "dec"])
    ap.add_argument("-e", "--seed", type=int, default=1)
    ap.add_argument("-x", "--rmexp", action="store_true", help="feed the model a PRE-EXPANDED key")
    ap.add_argument("--max-n-byte", type=int, default=2 ** 12 - 1)
    args = ap.parse_args()
    from aesgcm_b200 import gcm_model, key_exp, stimulus as st
    if args.config:
        cfg = st.load_config(args.config)
    else:
        cfg = {"seed": args.seed, "aes_mode": args.mode, "key": args.key.upper(), "iv": args.iv.upper(),
               "aad": args.aad.upper(), "data": args.data.upper(), "enc_dec": args.ed, "max_n_byte": args.max_n_byte}
    r = st.replay(cfg, gcm_model.gcm, pre_expanded=args.rmexp or bool(cfg.get("key_pre_exp")),
                  expand_key=key_exp.aes_expand_key)
    out = {"aes_mode": r["config"].get("aes_mode"), "enc_dec": r["config"].get("enc_dec", "enc"),
           "key": r["data"]["key"]["data"], "iv": r["data"]["iv"]["data"],
           "aad_n_bytes": r["data"]["aad_n_bytes"], "pt_n_bytes": r["data"]["pt_n_bytes"],
           "ct": b"".join(r["ct_words"]).hex().upper(), "tag": r["tag"].hex().upper()}
    if "dec_words" in r:
        out["dec_pt_matches"] = b"".join(r["dec_words"]) == b"".join(r["pt_words"])
        out["dec_tag"] = r["dec_tag"].hex().upper()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
