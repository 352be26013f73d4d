"""Host-side multi-GPU logic on CPU: shard plan, message split, and the 16-byte
all_gather + combine over gloo with world_size 2 (the per-rank arithmetic is done
by the oracle here; on GPUs it is agcm_stream_part / agcm_stream_finish)."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT


def test_shard_plan_covers_message():
    from aesgcm_b200.parallel import shard_plan
    for n in (0, 1, 15, 16, 17, 160, 1000, 2 ** 30, 2 ** 30 + 5):
        for w in (1, 2, 3, 4, 8):
            plan = shard_plan(n, w)
            assert len(plan) == w
            assert sum(s.n_bytes for s in plan) == n
            blocks = (n + 15) // 16
            pos = 0
            for s in plan:
                assert s.byte_offset == s.first_block * 16
                if s.n_bytes:
                    assert s.byte_offset == pos
                    pos += s.n_bytes
                nb = (s.n_bytes + 15) // 16
                assert s.first_block + nb + s.blocks_after == blocks
                if s.blocks_after and s.n_bytes:
                    assert s.n_bytes % 16 == 0  # only the last shard may be ragged
    with pytest.raises(OverflowError):
        shard_plan(16 * 0xFFFFFFFF, 8)


def test_batch_split_balanced():
    from aesgcm_b200.parallel import batch_split
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 4, 8):
            spans = [batch_split(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_bytes, q):
    import sys
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from aesgcm_b200.parallel import gather_partials, shard_plan
    from oracle import cpu_oracle as o
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(1)
    key = rng.integers(0, 256, 32, dtype=np.uint8).tobytes()
    iv = rng.integers(0, 256, 12, dtype=np.uint8).tobytes()
    aad = rng.integers(0, 256, 16, dtype=np.uint8).tobytes()
    pt = rng.integers(0, 256, n_bytes, dtype=np.uint8).tobytes()
    rk = o.key_expand(key)
    h, ej0 = o.h_ej0(rk, iv)
    sh = shard_plan(n_bytes, world)[rank]
    mine = pt[sh.byte_offset:sh.byte_offset + sh.n_bytes]
    ct = o.gctr(rk, iv, 2 + sh.first_block, mine)                    # counter-range shard
    part = o.gfmul(o.gf_pow(h, sh.blocks_after), o.ghash_absorb(h, ct))  # scaled like agcm_stream_part
    parts = gather_partials(torch.frombuffer(bytearray(part), dtype=torch.uint8))
    y = bytes(16)
    for r in range(world):
        y = bytes(a ^ b for a, b in zip(y, parts[r].numpy().tobytes()))
    qa = o.gfmul(o.gf_pow(h, (n_bytes + 15) // 16), o.ghash_absorb(h, aad))
    y = bytes(a ^ b for a, b in zip(y, qa))
    lenblk = (len(aad) * 8).to_bytes(8, "big") + (n_bytes * 8).to_bytes(8, "big")
    y = o.ghash_absorb(h, lenblk, y)
    tag = bytes(a ^ b for a, b in zip(y, ej0))
    all_ct = [None] * world
    dist.all_gather_object(all_ct, ct)
    if rank == 0:
        want_ct, want_tag = o.gcm_crypt(key, iv, aad, pt)
        q.put((b"".join(all_ct) == want_ct, tag == want_tag))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_bytes", [0, 100, 4096 + 7])
def test_gloo_world2_counter_range_split(n_bytes):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_bytes, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == (True, True)
