#!/usr/bin/env python
"""Turn the raw outputs of a gpurun round into the tracked tables under profiles/:
  gpurun_out/lane_sweep.jsonl  -> profiles/r1_lane_sweep.md      (tools/sweep_lanes.py)
  gpurun_out/<tag>_variants.log -> profiles/r1_kernel_table.md    (tools/bench_variants.py)
  gpurun_out/config5.json      -> profiles/r1_config5_sweep.json (tools/sweep_config5.py) + a summary table
Usage: python tools/make_profiles.py [tag]"""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "fin"
RND = sys.argv[2] if len(sys.argv) > 2 else "r1"   # file-name prefix of the round


def jl(path):
    return [json.loads(l) for l in open(path) if l.startswith("{")]


def lane_sweep():
    rows = jl(os.path.join(OUT, "lane_sweep.jsonl"))
    cols = ["auto", "g1", "g2", "g4", "g8", "g16", "g32", "g1024"]
    out = ["# Lanes-per-message sweep of the shared-key batch path (round 1, final build)", "",
           "`python tools/sweep_lanes.py --aes 128,256 --aad 0,64` on one B200: payload GB/s of `agcm_batch_crypt_uniform`,",
           "about 1 GiB of payload per point, device-resident inputs, CUDA events, 3 iterations. `auto` = `lanes=0`",
           "(`pick_lanes` in `csrc/capi.cu`; it may also cut few long messages into per-CTA segments, which no explicit",
           "column does); g1024 = one CTA per whole message (`k_batch_cta`). Every lane count is asserted to give identical",
           "tags. Bold = best explicit lane count of the row.", "",
           "| AES | message | AAD | messages | " + " | ".join(cols) + " |", "|---|---|---|---|" + "---|" * len(cols)]
    worst = 1.0
    for r in rows:
        best = max(r.get(c, 0) for c in cols[1:])
        worst = min(worst, r["auto"] / best)
        cells = []
        for c in cols:
            v = r.get(c)
            cells.append("" if v is None else ("**%.0f**" % v if (c != "auto" and v == best) else "%.0f" % v))
        out.append("| %d | %d B | %d B | %d | " % (r["aes"], r["msg_bytes"], r["aad_bytes"], r["n_msgs"]) + " | ".join(cells) + " |")
    out += ["", "Reading: the best lane count grows like sqrt(blocks)/4 (1 lane below 16 blocks, 2 at 1-1.5 KB, 4 at 4 KB, 8 at 16 KB,",
            "32 from 64 KB); for long messages the library compares the quantised step counts of the lane-group layout, the",
            "CTA-per-message layout and the CTA-per-segment layout (rounds x (rows per lane + per-unit overhead)) and takes the",
            "shortest. `auto` is never more than %.1f %% below the best explicit column, and above it where segments win" % (100 * (1 - worst)),
            "(4 MiB x 256 messages: 256 whole messages on 148 CTAs are 2 rounds at 86 % occupancy).",
            "History: with 2 lanes for every aligned record 16 KiB messages ran at 407 GB/s (AES-256). Wide groups (16 / 32 lanes)",
            "used to combine their lanes with G-1 serial table products on the lookup pipe; they now apply one generic product per",
            "lane on the integer pipe and a butterfly XOR (64 KiB: 473 -> 497 GB/s; 1500 B with 32 lanes: 137 -> 315). `k_batch_cta`",
            "uses the stream kernel's counter-byte cache (its lanes step the counter by 512): 1 MiB messages 479 -> 498 GB/s."]
    open(os.path.join(PROF, RND + "_lane_sweep.md"), "w").write("\n".join(out) + "\n")


def kernel_table():
    rows = jl(os.path.join(OUT, tag + "_variants.log"))
    out = ["# Kernel-level throughput, one B200, inputs resident in HBM (tools/bench_variants.py, CUDA events, 10 iterations, final build)", "",
           "Payload GB/s = message bytes / time. SM clock 1965 MHz, no throttle reasons.", "",
           "| path | AES | op | lanes/msg | ms | GB/s | Mmsg/s |", "|---|---|---|---|---|---|---|"]
    keys = []
    for r in rows:
        if r["op"] == "key schedule":
            keys.append(r)
            continue
        out.append("| %s | %d | %s | %s | %.4f | %.1f | %s |" % (r["path"], r["aes"], r["op"], r.get("lanes", ""), r["ms"], r["GBps"],
                                                             r.get("Mmsg_per_s", "")))
    out += ["", "Stream = `k_stream` (BASELINE config 2 shape). Packets = `k_batch` under one shared pre-expanded key (config 3; for `lanes = 0` the library picks the lane count from the message length, `profiles/r1_lane_sweep.md`: 2 at 1500 B). Per-message key = `k_batch_perkey` (config 4; `profiles/r1_ncu_perkey.md`).", "",
            "## Key schedule on the device (tools/bench_variants.py --only keys)", "",
            "| kernel | AES | 2^20 keys, ms | keys/s | output GB/s |", "|---|---|---|---|---|"]
    for i, r in enumerate(keys):
        out.append("| %s | %d | %.3f | %.2f G | %.0f |" % ("`k_key_expand` (one thread per key, tb/key_exp.py:118 semantics)" if i == 0 else "",
                                                      r["aes"], r["ms"], r["Mkeys_per_s"] / 1000, r["GBps"]))
    sk = sum(r["set_key_us"] for r in keys) / max(1, len(keys))
    out += ["", "`agcm_set_key` (one launch: key schedule + H + 63 linear squarings `gf_sqr` + power tables + Shoup tables, one 272 B readback, synchronous): ~%.0f µs per key (was ~215 µs with two launches, three copies and bit-serial squarings)." % sk]
    open(os.path.join(PROF, RND + "_kernel_table.md"), "w").write("\n".join(out) + "\n")


def config5():
    src = os.path.join(OUT, tag + "_config5.json") if os.path.exists(os.path.join(OUT, tag + "_config5.json")) else os.path.join(OUT, "config5.json")
    shutil.copy(src, os.path.join(PROF, RND + "_config5_sweep.json"))
    rows = json.load(open(src))
    ratios = ["0", "1/16", "1/4", "1", "4", "16"]
    out = ["# BASELINE config 5: message size x AAD/PT ratio (tools/sweep_config5.py, final build)", "",
           "(PT + AAD) GB/s, about 1 GiB per point, device-resident inputs, CUDA events; `s` marks points that ran as one stream",
           "call (a single message), every other point is ONE batch call. Raw rows with payload-only rates:",
           "`profiles/%s_config5_sweep.json`." % RND, ""]
    for aes in (128, 256):
        out += ["## AES-%d" % aes, "", "| message \\ AAD:PT | " + " | ".join(ratios) + " |", "|---|" + "---|" * len(ratios)]
        sizes = sorted({r["msg_bytes"] for r in rows})
        for sz in sizes:
            rs = sorted([r for r in rows if r["aes"] == aes and r["msg_bytes"] == sz], key=lambda r: r["aad_bytes"])
            out.append("| %d B | " % sz + " | ".join("%.0f%s" % (r["pt_plus_aad_GBps"], " s" if r["api"] == "stream" else "") for r in rs) + " |")
        out.append("")
    open(os.path.join(PROF, RND + "_config5_sweep.md"), "w").write("\n".join(out))


if __name__ == "__main__":
    for fn, need in ((lane_sweep, "lane_sweep.jsonl"), (kernel_table, tag + "_variants.log"), (config5, "config5.json")):
        if os.path.exists(os.path.join(OUT, need)) or (fn is config5 and os.path.exists(os.path.join(OUT, tag + "_config5.json"))):
            fn()
            print("wrote", fn.__name__)
