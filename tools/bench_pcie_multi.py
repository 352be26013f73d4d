#!/usr/bin/env python
"""Host<->device copy ceiling with 1, 2, 4, ... N ranks copying AT THE SAME TIME (control measurement
for the end-to-end figure of bench.py at N > 1: is it the link / host memory, or the engine?).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_pcie_multi.py

For n in 1, 2, 4, ... world: ranks 0..n-1 each copy `--mib` MiB host->device and `--mib` MiB device->host
simultaneously (two streams, pinned buffers), the rest idle; wall clock between two barriers; aggregate GB/s
per direction.  Twice: pinned memory allocated with the process's default CPU affinity, and allocated after
binding the process to the CPUs of the GPU's own NUMA node (sysfs), when the box exposes that.  Then the
engine's host-buffer call (agcm_stream_crypt_host) on the same ranks."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def numa_cpus(dev):
    try:
        p = torch.cuda.get_device_properties(dev)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None, node
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        return (cpus & allowed) or None, node
    except Exception as ex:
        return None, "unknown (%s)" % type(ex).__name__


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=4)
    args = ap.parse_args()
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    dev = torch.device("cuda", lr)
    n = args.mib << 20
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    cpus, node = numa_cpus(lr)
    base_aff = os.sched_getaffinity(0)
    results = []

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sizes = [k for k in (1, 2, 4, 8, 16) if k <= world]
    for mode in ("default affinity", "bound to the GPU's NUMA node"):
        if mode != "default affinity":
            if not cpus:
                if rank == 0:
                    results.append({"mode": mode, "skipped": "NUMA node of the GPU not exposed (numa_node = %s)" % (node,)})
                continue
            os.sched_setaffinity(0, cpus)
        h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        h_in.fill_(7)
        for k in sizes:
            active = rank < k

            def step():
                if active:
                    with torch.cuda.stream(s1):
                        d_a.copy_(h_in, non_blocking=True)
                    with torch.cuda.stream(s2):
                        h_out.copy_(d_b, non_blocking=True)
            step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.iters):
                step()
            barrier()
            dt = allmax((time.perf_counter() - t0) / args.iters)
            if rank == 0:
                results.append({"mode": mode, "ranks_copying": k, "GBps_each_direction_aggregate": round(k * n / dt / 1e9, 2),
                                "GBps_each_direction_per_rank": round(n / dt / 1e9, 2)})
        del h_in, h_out
    os.sched_setaffinity(0, base_aff)
    # the engine's host-buffer call on the same ranks (independent messages: no exchange in the way)
    import aesgcm_b200
    eng = aesgcm_b200.GcmEngine(lr)
    eng.set_key(bytes(range(32)))
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_in.fill_(3)
    for k in sizes:
        active = rank < k
        if active:
            eng.encrypt(bytes(12), b"", h_in, out=h_out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.iters):
            if active:
                eng.encrypt(bytes(12), b"", h_in, out=h_out)
        barrier()
        dt = allmax((time.perf_counter() - t0) / args.iters)
        if rank == 0:
            results.append({"mode": "engine: agcm_stream_crypt_host, default affinity", "ranks_copying": k,
                            "GBps_each_direction_aggregate": round(k * n / dt / 1e9, 2),
                            "GBps_each_direction_per_rank": round(n / dt / 1e9, 2)})
    if rank == 0:
        print(json.dumps({"world": world, "mib_per_rank_per_direction": args.mib, "cpu_count": os.cpu_count(),
                          "affinity_cpus": len(base_aff), "gpu0_numa_node": node, "results": results}, indent=1), flush=True)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
