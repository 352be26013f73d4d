#!/usr/bin/env python
"""Timeline of the host-buffer schedule (H2D copy -> fused kernel -> D2H copy per granule, three granules in flight)
rebuilt from the device API with an event after every operation: where does the time between the link rate and the
measured end-to-end rate go?  Usage: pipeline_timeline.py TOTAL_MiB GRANULE_MiB [d2d]   (d2d: a device copy instead
of the cipher kernel).  Prints one line per granule (µs from the start) and the busy time of both copy directions."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import aesgcm_b200

total, gran = int(sys.argv[1]) << 20, int(sys.argv[2]) << 20
kind = sys.argv[3] if len(sys.argv) > 3 else "kstream"   # kstream | d2d | tiny | none | indep (k_stream on its own stream, no dependency)
d2d = kind == "d2d"
quiet = len(sys.argv) > 4 and "q" in sys.argv[4]
roles = len(sys.argv) > 4 and "r" in sys.argv[4]
ahead = len(sys.argv) > 4 and "a" in sys.argv[4]
slots = int(os.environ.get("TL_SLOTS", "3"))
h_in = torch.empty(total, dtype=torch.uint8, pin_memory=True); h_in.random_(0, 256)
h_out = torch.empty(total, dtype=torch.uint8, pin_memory=True)
engs = [aesgcm_b200.GcmEngine(0) for _ in range(slots)]
for e in engs: e.set_key(bytes(range(32)))
st = [torch.cuda.Stream() for _ in range(slots)]
stage = [torch.empty(gran, dtype=torch.uint8, device="cuda") for _ in range(slots)]
stage2 = [torch.empty(gran, dtype=torch.uint8, device="cuda") for _ in range(slots)]
part = [torch.zeros(16, dtype=torch.uint8, device="cuda") for _ in range(slots)]
iv = bytes(12)
side = torch.cuda.Stream()
n_gran = total // gran


def once(record):
    ev = []
    t0 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record(st[0])
    for s in range(1, slots): st[s].wait_event(t0)
    if ahead:
        # per-slot streams, but every H2D copy is SUBMITTED as early as its slot allows (right after the D2H copy that
        # frees the slot), so that a D2H copy waiting for its kernel is never queued in front of a copy that could run
        def h2d(k):
            with torch.cuda.stream(st[k % slots]):
                stage[k % slots].copy_(h_in[k * gran:(k + 1) * gran], non_blocking=True)
        for k in range(min(slots, n_gran)): h2d(k)
        for k in range(n_gran):
            s = k % slots
            with torch.cuda.stream(st[s]):
                if kind == "kstream":
                    engs[s].stream_part_device(0, iv, k * gran // 16, stage[s], stage[s], (n_gran - 1 - k) * gran // 16, part[s])
                elif kind == "tiny":
                    part[s].add_(1)
                h_out[k * gran:(k + 1) * gran].copy_(stage[s], non_blocking=True)
            if k + slots < n_gran: h2d(k + slots)
        torch.cuda.synchronize()
        return t0, ev
    if roles:
        # one stream per ROLE: all H2D copies in order on one, all kernels on the second, all D2H copies on the third
        done_d2h = [None] * n_gran
        for s in range(3): st[s].wait_event(t0)
        for k in range(n_gran):
            s = k % slots
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            with torch.cuda.stream(st[0]):
                if k >= slots: st[0].wait_event(done_d2h[k - slots])
                e[0].record()
                stage[s].copy_(h_in[k * gran:(k + 1) * gran], non_blocking=True)
                e[1].record()
            with torch.cuda.stream(st[1]):
                st[1].wait_event(e[1])
                if kind == "kstream":
                    engs[0].stream_part_device(0, iv, k * gran // 16, stage[s], stage[s], (n_gran - 1 - k) * gran // 16, part[s])
                elif kind == "tiny":
                    part[s].add_(1)
                e[2].record()
            with torch.cuda.stream(st[2]):
                st[2].wait_event(e[2])
                h_out[k * gran:(k + 1) * gran].copy_(stage[s], non_blocking=True)
                e[3].record()
            done_d2h[k] = e[3]
            ev.append(e)
        torch.cuda.synchronize()
        return t0, ev
    for k in range(n_gran):
        s = k % slots
        with torch.cuda.stream(st[s]):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if record else None
            if record: e[0].record()
            stage[s].copy_(h_in[k * gran:(k + 1) * gran], non_blocking=True)
            if record: e[1].record()
            src = stage[s]
            if d2d:
                stage2[s].copy_(stage[s]); src = stage2[s]
            elif kind == "tiny":
                part[s].add_(1)
            elif kind == "none":
                pass
            elif kind == "indep":
                with torch.cuda.stream(side):
                    engs[s].stream_part_device(0, iv, k * gran // 16, stage2[s], stage2[s], (n_gran - 1 - k) * gran // 16, part[s])
            else:
                engs[s].stream_part_device(0, iv, k * gran // 16, stage[s], stage[s], (n_gran - 1 - k) * gran // 16, part[s])
                src = stage[s]
            if record: e[2].record()
            h_out[k * gran:(k + 1) * gran].copy_(src, non_blocking=True)
            if record: e[3].record()
            ev.append(e)
    torch.cuda.synchronize()
    return t0, ev


once(False)
import time
t = time.perf_counter(); once(False); dt = time.perf_counter() - t
if ahead:
    ts = []
    for _ in range(3):
        t = time.perf_counter(); once(False); ts.append(time.perf_counter() - t)
    print(json.dumps({"total_MiB": total >> 20, "granule_MiB": gran >> 20, "kernel": kind, "streams": "per slot, H2D submitted ahead",
                      "slots": slots, "GBps": [round(total / x / 1e9, 2) for x in ts]}))
    sys.exit(0)
t0, ev = once(True)
rows = []
for k, e in enumerate(ev):
    rows.append([round(t0.elapsed_time(x) * 1e3, 1) for x in e])
print(json.dumps({"total_MiB": total >> 20, "granule_MiB": gran >> 20, "kernel": kind, "streams": "per slot, H2D submitted ahead" if ahead else "per role" if roles else "per slot", "slots": slots,
                  "GBps_untimed_events": round(total / dt / 1e9, 2)}))
steady = rows[3:-2]
if steady:
    print(json.dumps({"steady_period_us": round((steady[-1][1] - steady[0][1]) / (len(steady) - 1), 1),
                      "link_period_us_at_48.9": round(gran / 48.9e3, 1),
                      "kernel_us_mean": round(sum(r[2] - r[1] for r in steady) / len(steady), 1)}))
if quiet:
    sys.exit(0)
print("granule: h2d_ready  h2d_done  kernel_done  d2h_done   | h2d µs  kernel µs  d2h µs | gap before h2d (vs previous h2d_done)")
prev = 0.0
for k, r in enumerate(rows):
    # the copy starts when both its stream is ready (r[0]) and the engine is free (previous h2d done)
    start = max(r[0], prev)
    print("%3d: %9.1f %9.1f %9.1f %9.1f | %7.1f %7.1f %7.1f | %6.1f" % (k, r[0], r[1], r[2], r[3], r[1] - start, r[2] - r[1],
                                                                 r[3] - max(r[2], rows[k - 1][3] if k else 0), start - prev))
    prev = r[1]
print(json.dumps({"end_us": rows[-1][3], "link_only_us_at_48.9": round(total / 48.9e3, 1)}))
