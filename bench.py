#!/usr/bin/env python
"""bench.py -- AES-GCM enc+tag throughput of the B200 engine (BASELINE.json metric).

  python bench.py [--gpus N --steps K --warmup W] [--impl reference]

One "step" = one fused GCTR+GHASH pass (encrypt + tag) over one synthetic stream
resident in HBM.  N=1: BASELINE config 2, AES-256, one 2^30-byte stream, 16 B AAD.
N>1 (torchrun, one rank per GPU): the stream is N x 2^30 bytes, split by counter
range (weak scaling: 2^30 bytes per GPU); every step ends with the 16-byte-per-rank
NCCL all_gather and the tag finish, timed inside the step.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "aes_gcm_enc_tag_throughput"
UNIT = "GB/s"
SHARD_BYTES = 1 << 30
AAD_BYTES = 16


def stream_inputs():
    """BASELINE config 2 constants: key / IV / AAD from default_rng(1)."""
    rng = np.random.default_rng(1)
    key = rng.integers(0, 256, 32, dtype=np.uint8).tobytes()
    iv = rng.integers(0, 256, 12, dtype=np.uint8).tobytes()
    aad = rng.integers(0, 256, AAD_BYTES, dtype=np.uint8).tobytes()
    return key, iv, aad


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        self.path = None
        exe = shutil.which("nvidia-smi")
        if not exe:
            return
        fd, self.path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        self.f = open(self.path, "w")
        self.proc = subprocess.Popen([exe, "-i", str(index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                      "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        self.t0 = self.t1 = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_stop(self):
        self.t1 = time.time()

    def finish(self):
        if not self.proc:
            return None
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        rows = []
        import datetime
        with open(self.path) as f:
            for line in f:
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    rows.append((ts, float(c[1]), float(c[2]), c[4], c[5], c[6], c[7], c[8]))
                except Exception:
                    continue
        os.unlink(self.path)
        if not rows:
            return None
        inside = [r for r in rows if self.t0 and self.t1 and self.t0 - 0.05 <= r[0] <= self.t1 + 0.05]
        scope = "timed region"
        if not inside:
            inside, scope = rows, "whole bench (timed region shorter than the sampling period)"
        sm = sorted(r[1] for r in inside)
        reasons = set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in inside:
            for nm, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": inside[0][2], "reasons": sorted(reasons), "samples": len(inside),
                "scope": scope}


def openssl_speed(cores, seconds=2):
    """`openssl speed -evp aes-256-gcm` on the host (AES-NI/VAES+PCLMUL): the strong CPU
    baseline that stands in for pycryptodome's C core.  Returns GB/s at 16 KiB blocks or None."""
    exe = shutil.which("openssl")
    if not exe:
        return None
    try:
        out = subprocess.run([exe, "speed", "-evp", "aes-256-gcm", "-bytes", "16384", "-seconds", str(seconds), "-multi",
                              str(cores)], capture_output=True, text=True, timeout=60).stdout
        for line in out.splitlines():
            if line.lower().startswith("aes-256-gcm") or line.startswith("evp"):
                tok = line.split()[-1]
                if tok.endswith("k"):
                    return float(tok[:-1]) * 1000.0 / 1e9
    except Exception:
        return None
    return None


def python_model_rates():
    """The reference model's calling patterns (tb/gcm_model.py) on one host core with the AES-GCM
    library that IS installed (`cryptography`/OpenSSL stands in for pycryptodome): whole message,
    and the testbench's one-<=16-byte-block-per-call streaming.  GB/s each, or None."""
    try:
        from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes
        from cryptography.hazmat.primitives.ciphers.aead import AESGCM
    except Exception:
        return None
    key, iv, aad = stream_inputs()
    res = {}
    pt = np.random.default_rng(1).integers(0, 256, 64 << 20, dtype=np.uint8).tobytes()
    t0 = time.perf_counter()
    AESGCM(key).encrypt(iv, pt, aad)
    res["whole_message_64MiB_1core_GBps"] = round(len(pt) / (time.perf_counter() - t0) / 1e9, 3)
    n = 1 << 20
    enc = Cipher(algorithms.AES(key), modes.GCM(iv)).encryptor()
    enc.authenticate_additional_data(aad)
    t0 = time.perf_counter()
    for i in range(0, n, 16):
        enc.update(pt[i:i + 16])
    enc.finalize()
    res["per_16B_call_streaming_1MiB_1core_GBps"] = round(n / (time.perf_counter() - t0) / 1e9, 5)
    return res


def cpu_reference_rate(sample_bytes, threads, steps, warmup):
    """Times the CPU restatement of the reference datapath (oracle/gcm_oracle.c, all host
    threads) on the first `sample_bytes` of the config-2 stream.  -> (GB/s, seconds per step)."""
    from oracle import cpu_oracle as o
    key, iv, aad = stream_inputs()
    pt = np.random.default_rng(1).integers(0, 256, sample_bytes, dtype=np.uint8)
    for _ in range(warmup):
        o.gcm_crypt(key, iv, aad, pt[: max(1 << 16, sample_bytes // 16)], threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.gcm_crypt(key, iv, aad, pt, threads=threads)
    dt = (time.perf_counter() - t0) / steps
    return sample_bytes / dt / 1e9, dt


def calibrate_sample(threads, target_s):
    from oracle import cpu_oracle as o
    key, iv, aad = stream_inputs()
    n = 1 << 20
    pt = np.random.default_rng(1).integers(0, 256, n, dtype=np.uint8)
    o.gcm_crypt(key, iv, aad, pt[:65536], threads=threads)
    t0 = time.perf_counter()
    o.gcm_crypt(key, iv, aad, pt, threads=threads)
    rate = n / (time.perf_counter() - t0)
    s = int(rate * target_s) >> 20 << 20
    return max(1 << 20, min(s, SHARD_BYTES))


def run_reference(args, rank, world):
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    budget = float(os.environ.get("AGCM_BENCH_REF_BUDGET_S", "150"))  # seconds for the whole arm
    per_step = max(0.2, min(5.0, budget / max(1, args.steps + 1)))
    sample = calibrate_sample(cores, per_step)
    val, dt = cpu_reference_rate(sample, cores, args.steps, min(args.warmup, 1))
    quick = budget < 30
    ossl = None if quick else openssl_speed(cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(val, 6), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": round(val, 6), "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "first %d MiB of the config-2 stream per step, oracle/gcm_oracle.c with %d threads "
                                   "(pycryptodome, the reference's own backend, is not installed)" % (sample >> 20, cores),
                         "openssl_speed_evp_aes256gcm_allcores_GBps": ossl,
                         "python_cryptography_aesgcm": None if quick else python_model_rates(),
                         "pycryptodome": "unavailable (not installed; no network)"},
        "e2e": {"value": round(val, 6), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(n_gpus):
    return {"workload": "BASELINE config 2: AES-256-GCM encrypt+tag, one stream of %d x 2^30 B, %d B AAD, counter-range "
                        "split (2^30 B per GPU)" % (n_gpus, AAD_BYTES),
            "bytes_per_gpu": SHARD_BYTES, "aad_bytes": AAD_BYTES, "key_bits": 256,
            "parallelism": ("counter-range x%d + 16 B/rank exchange (%s)" % (n_gpus, os.environ.get("AGCM_BENCH_EXCHANGE", "peer-memory stores fused in the kernel tail"))) if n_gpus > 1 else "single GPU",
            "l2": "inputs (1 GiB read + 1 GiB written per step) exceed the 126 MB L2; no flush needed"}


def other_workloads(eng, torch):
    """Kernel-level payload GB/s of BASELINE configs 3 and 4 (device-resident, CUDA events, 5 passes
    after 2 warm-ups); secondary evidence only -- the headline metric is config 2."""
    def timeit(fn, iters=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    res = []
    n_msgs, length, stride = 1 << 20, 1500, 1504
    rng = np.random.default_rng(2)
    d_buf = torch.randint(0, 256, (n_msgs * stride,), dtype=torch.uint8, device="cuda")
    d_out = torch.empty_like(d_buf)
    d_iv = torch.randint(0, 256, (n_msgs * 12,), dtype=torch.uint8, device="cuda")
    d_tags = torch.zeros(16 * n_msgs, dtype=torch.uint8, device="cuda")
    d_ok = torch.zeros(n_msgs, dtype=torch.uint8, device="cuda")
    eng.set_key(rng.integers(0, 256, 24, dtype=np.uint8).tobytes())
    eng.set_key(eng.round_keys())  # shared PRE-EXPANDED key
    ms = timeit(lambda: eng.batch_crypt_uniform_device(0, d_iv, None, 0, 0, d_buf, d_out, length, stride, d_tags, n_msgs=n_msgs))
    res.append({"workload": "config 3: AES-192 encrypt+tag, 2^20 x 1500 B at a 1504 B stride, per-message IV, shared "
                            "pre-expanded key", "ms": round(ms, 4), "payload_GBps": round(n_msgs * length / ms / 1e6, 1)})
    del d_buf, d_out
    alen = 64
    d_keys = torch.randint(0, 256, (n_msgs * 32,), dtype=torch.uint8, device="cuda")
    d_aad = torch.randint(0, 256, (n_msgs * alen,), dtype=torch.uint8, device="cuda")
    d_pt = torch.randint(0, 256, (n_msgs * length,), dtype=torch.uint8, device="cuda")
    d_ct = torch.empty_like(d_pt)
    d_back = torch.empty_like(d_pt)
    eng.batch_crypt_perkey_uniform_device(256, 0, d_keys, d_iv, d_aad, alen, alen, d_pt, d_ct, length, length, d_tags, n_msgs=n_msgs)
    ms = timeit(lambda: eng.batch_crypt_perkey_uniform_device(256, 1, d_keys, d_iv, d_aad, alen, alen, d_ct, d_back, length,
                                                              length, d_tags, d_ok, n_msgs=n_msgs))
    torch.cuda.synchronize()
    assert int(d_ok.sum().item()) == n_msgs and torch.equal(d_back, d_pt), "config 4 round trip failed"
    res.append({"workload": "config 4: AES-256 decrypt+verify, 2^20 x 1500 B, distinct key per message (schedule on "
                            "device), 64 B AAD", "ms": round(ms, 4), "payload_GBps": round(n_msgs * length / ms / 1e6, 1),
                "Mmsg_per_s": round(n_msgs / ms / 1e3, 1)})
    return res


def run_ours(args, rank, world, local_rank):
    import torch
    import aesgcm_b200
    from aesgcm_b200.parallel import shard_plan, gather_partials

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    key, iv, aad = stream_inputs()
    eng = aesgcm_b200.GcmEngine(local_rank)
    eng.set_key(key)
    total = SHARD_BYTES * world
    shard = shard_plan(total, world)[rank]
    gen = torch.Generator(device=dev)
    gen.manual_seed(1 + rank)
    d_in = torch.randint(0, 256, (shard.n_bytes,), dtype=torch.uint8, device=dev, generator=gen)
    d_out = torch.empty_like(d_in)
    d_aad = torch.from_numpy(np.frombuffer(aad, dtype=np.uint8).copy()).to(dev)
    d_tag = torch.zeros(16, dtype=torch.uint8, device=dev)
    d_part = torch.zeros(16, dtype=torch.uint8, device=dev)
    d_parts = torch.zeros((world, 16), dtype=torch.uint8, device=dev)

    # N > 1: the 16-byte partials cross NVLink as peer-memory stores from the kernel's own tail
    # (one launch per rank per step); AGCM_BENCH_EXCHANGE=nccl selects part + all_gather + finish
    exchange = os.environ.get("AGCM_BENCH_EXCHANGE", "peer") if world > 1 else "none"
    px = None
    if exchange == "peer":
        try:
            from aesgcm_b200.parallel import PeerExchange
            px = PeerExchange(eng)
        except Exception as ex:  # symmetric memory not available on this node: use the library exchange
            px = None
            exchange = "nccl (peer-memory rendezvous failed: %s)" % type(ex).__name__
            os.environ["AGCM_BENCH_EXCHANGE"] = exchange

    def step():
        if world == 1:
            eng.stream_crypt_device(0, iv, d_aad, d_in, d_out, d_tag)
        elif px is not None:
            px.crypt(0, iv, d_aad, shard, d_in, d_out, total, d_tag)
        else:
            eng.stream_part_device(0, iv, shard.first_block, d_in, d_out, shard.blocks_after, d_part)
            dist.all_gather_into_tensor(d_parts.view(-1), d_part)
            eng.stream_finish_device(0, iv, d_parts, world, d_aad, total, d_tag)

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    eng.timing_enable(True)
    l0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler:
        sampler.mark_start()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    if sampler:
        sampler.mark_stop()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - l0
    k_ms, k_n = eng.timing_read()
    eng.timing_enable(False)
    if dist:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = total * args.steps / (ms * 1e-3) / 1e9

    # the work was real: the tag verifies and the plaintext round-trips on this rank's shard
    d_chk = torch.empty_like(d_in)
    eng.gctr_device(iv, shard.first_block, d_out, d_chk)
    torch.cuda.synchronize()
    assert torch.equal(d_chk, d_in), "round trip failed"
    del d_chk

    # ---- e2e: the same step through the host-buffer API, copies inside the timed region
    e2e_steps = max(3, min(args.steps, 10))
    h_in = torch.empty(shard.n_bytes, dtype=torch.uint8, pin_memory=True)
    h_in.copy_(d_in)
    h_out = torch.empty(shard.n_bytes, dtype=torch.uint8, pin_memory=True)

    def e2e_step():
        if world == 1:
            _, tag = eng.encrypt(iv, aad, h_in, out=h_out)
            return tag
        part = eng.stream_part_host(0, iv, shard.first_block, h_in, h_out, shard.blocks_after)
        d_part.copy_(torch.frombuffer(bytearray(part), dtype=torch.uint8))
        dist.all_gather_into_tensor(d_parts.view(-1), d_part)
        return eng.stream_finish_host(0, iv, d_parts.cpu().numpy(), aad, total)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        tag_e2e = e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if dist:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    assert tag_e2e == d_tag.cpu().numpy().tobytes(), "host-buffer path and device path disagree on the tag"
    clocks = sampler.finish() if sampler else None

    if rank == 0:
        peak, peak_src = peaks()
        # algorithmic bytes of ONE launch of the dominant kernel k_stream<14,ENC> (SURVEY 8d /
        # BASELINE.md 3): PT read + CT written + IV + key; AAD and tag belong to the finish kernel
        alg = 2 * shard.n_bytes + 12 + 32
        k_avg_ms = k_ms / max(k_n, 1)
        achieved = alg / (k_avg_ms * 1e-3) / 1e9
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        # L1/shared-memory data-pipe bound of the formulation: per warp-row of 32
        # blocks 202 AES word lookups + 16 GHASH 128-bit row lookups (4 wavefronts each) + 8 global
        # load/store wavefronts, through a pipe that retires 1 wavefront (128 B) per clock per SM
        wavefronts_per_row = 202 + 64 + 8
        smem_bound = eng.sm_count * sm_mhz * 1e6 / wavefronts_per_row * 32 * 16 / 1e9  # payload GB/s
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                tj = json.load(f)
            if tj.get("n_bytes") == shard.n_bytes:
                traffic = int(tj["dram_bytes_read"] + tj["dram_bytes_write"])
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(world),
            "gpu_launches": int(launches),
            "e2e": {"value": round(total / e2e_s / 1e9, 3), "unit": UNIT,
                    "h2d_bytes_per_step": int(shard.n_bytes + AAD_BYTES + (16 * world if world > 1 else 0)),
                    "d2h_bytes_per_step": int(shard.n_bytes + 16), "steps": e2e_steps,
                    "api": "GcmEngine.encrypt (agcm_stream_crypt_host), pinned host buffers" if world == 1 else
                           "GcmEngine.stream_part_host + all_gather + stream_finish_host, pinned host buffers"},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": traffic, "kernel": "k_stream<14,ENC>",
                         "kernel_ms": round(k_avg_ms, 4), "kernel_launches": int(k_n), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(alg)},
            "bound_smem_lookup": {"payload_GBps_bound": round(smem_bound, 1), "sm_mhz_used": sm_mhz,
                                  "lsu_wavefronts_per_32_blocks": wavefronts_per_row,
                                  "frac": round(shard.n_bytes / (k_avg_ms * 1e-3) / 1e9 / smem_bound, 4),
                                  "note": "the binding roofline of this formulation is the 128 B/clk/SM L1/shared-memory "
                                          "data pipe (ncu: l1tex__data_pipe_lsu_wavefronts 97 % of peak), not HBM "
                                          "(DESIGN.md 4.1, profiles/r1_ncu_stream_final.md)"},
            "clocks": clocks,
        }
        if world == 1 and not args.no_extras:
            try:
                del d_in, d_out, h_in, h_out
                torch.cuda.empty_cache()
                line["other_workloads"] = other_workloads(eng, torch)
            except Exception as ex:  # never lose the headline line over a secondary measurement
                line["other_workloads"] = {"error": str(ex)}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample = calibrate_sample(cores, 4.0)
            v, dt = cpu_reference_rate(sample, cores, 3, 1)
            line["cpu_baseline"] = {"value": round(v, 6), "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "first %d MiB of the same stream, 3 passes, oracle/gcm_oracle.c with %d "
                                              "threads (%.1f s per pass)" % (sample >> 20, cores, dt),
                                    "openssl_speed_evp_aes256gcm_allcores_GBps": openssl_speed(cores),
                                    "python_cryptography_aesgcm": python_model_rates(),
                                    "pycryptodome": "unavailable (not installed; no network)"}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    eng.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
