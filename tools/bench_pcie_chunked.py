#!/usr/bin/env python
"""What a chunked H2D -> (kernel) -> D2H pipeline can reach on this box WITHOUT the cipher: the same
3-slot schedule as agcm_stream_crypt_host (csrc/capi.cu host_pipeline) with the kernel replaced by
nothing or by a device-to-device copy.  Puts the e2e figure of bench.py in context."""
import json, sys, time, torch
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
def run(chunk, slots, kernel):
    st = [torch.cuda.Stream() for _ in range(slots)]
    stage = [torch.empty(chunk, dtype=torch.uint8, device="cuda") for _ in range(slots)]
    stage2 = [torch.empty(chunk, dtype=torch.uint8, device="cuda") for _ in range(slots)]
    def once():
        for k in range(n // chunk):
            s = k % slots
            with torch.cuda.stream(st[s]):
                stage[s].copy_(h_in[k * chunk:(k + 1) * chunk], non_blocking=True)
                src = stage[s]
                if kernel:
                    stage2[s].copy_(stage[s]); src = stage2[s]
                h_out[k * chunk:(k + 1) * chunk].copy_(src, non_blocking=True)
        torch.cuda.synchronize()
    once()
    t0 = time.perf_counter()
    for _ in range(4): once()
    return n / ((time.perf_counter() - t0) / 4) / 1e9
for chunk_mb in (8, 16, 32, 64):
    for slots in (3, 4):
        for kernel in (False, True):
            print(json.dumps({"chunk_MiB": chunk_mb, "slots": slots, "d2d_copy_as_kernel": kernel,
                              "GBps_each_direction": round(run(chunk_mb << 20, slots, kernel), 1)}), flush=True)
