#!/usr/bin/env python
"""Static SASS statistics of one kernel of libaesgcm_b200.so: instruction mix of
the whole function and of its hottest loop (largest backward branch span).
usage: tools/sass_stats.py <substring of mangled name> [so]"""
import re
import subprocess
import sys

so = sys.argv[2] if len(sys.argv) > 2 else "aes-gcm-128-192-256-bits_b200/libaesgcm_b200.so"
pat = sys.argv[1]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s+Function : ", txt)
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", f)
    ins = [(int(a, 16), t.strip()) for a, t in ins]
    # hottest loop = the longest backward branch
    best = None
    for a, t in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and (best is None or a - tgt > best[1] - best[0]):
                best = (tgt, a)
    def mix(sel):
        d = {}
        for a, t in sel:
            t = re.sub(r"^@!?U?P\d\s+", "", t)
            op = t.split()[0]
            base = op.split(".")[0]
            key = base if base not in ("LDS", "LDG", "STG", "LDL", "STL", "STS") else op
            d[key] = d.get(key, 0) + 1
        return d
    print(name, "total", len(ins))
    if best:
        loop = [(a, t) for a, t in ins if best[0] <= a <= best[1]]
        m = mix(loop)
        print(" loop 0x%x..0x%x: %d instr" % (best[0], best[1], len(loop)))
        print("  ", ", ".join("%s=%d" % kv for kv in sorted(m.items(), key=lambda kv: -kv[1])))
