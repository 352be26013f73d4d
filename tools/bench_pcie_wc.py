#!/usr/bin/env python
"""Does write-combined pinned memory (cudaHostAllocWriteCombined) for the H2D source raise the host<->device
copy ceiling on this box?  1 GiB each way at once, two streams, default vs write-combined source."""
import ctypes, sys, os
import torch
rt = ctypes.CDLL("libcudart.so")
n = 1 << 30
def halloc(flags):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(flags)) == 0
    return p
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h_out = halloc(0)
for name, flags in (("default", 0), ("write-combined", 4)):
    h_in = halloc(flags)
    ctypes.memset(h_in, 7, n)
    def step(both=True):
        rt.cudaMemcpyAsync(ctypes.c_void_p(d_a.data_ptr()), h_in, ctypes.c_size_t(n), 1, ctypes.c_void_p(s1.cuda_stream))
        if both:
            rt.cudaMemcpyAsync(h_out, ctypes.c_void_p(d_b.data_ptr()), ctypes.c_size_t(n), 2, ctypes.c_void_p(s2.cuda_stream))
    for both in (False, True):
        step(both); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        import time
        t0 = time.perf_counter()
        for _ in range(5): step(both)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        print("%-15s H2D source, %s: %.1f GB/s per direction" % (name, "H2D + D2H at once" if both else "H2D alone", n / dt / 1e9), flush=True)
    rt.cudaFreeHost(h_in)
