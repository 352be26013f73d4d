import ctypes
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/gcm_oracle.c): the checker, never the product."""
    from oracle import cpu_oracle
    cpu_oracle.lib()
    return cpu_oracle


@pytest.fixture(scope="session")
def emul():
    """Host build of the engine's per-thread device code (tests/host_emul.cu)."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    bdir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(bdir, exist_ok=True)
    so = os.path.join(bdir, "libhost_emul.so")
    src = os.path.join(ROOT, "tests", "host_emul.cu")
    csrc = os.path.join(ROOT, "aes-gcm-128-192-256-bits_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("gcm_core.cuh", "aes_core.cuh", "gf128.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call([nvcc, "-w", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-o", so, src])
    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def engine_lib():
    """The product library; built in-tree if missing (nvcc cross-compiles on CPU)."""
    import aesgcm_b200
    if not os.path.exists(aesgcm_b200._lib.SO_PATH):
        aesgcm_b200._lib.build()
    return aesgcm_b200._lib.lib()


@pytest.fixture(scope="session")
def engine(engine_lib):
    import aesgcm_b200
    eng = aesgcm_b200.GcmEngine(0)
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def engine_small(engine_lib):
    """A context with a small, odd persistent grid (5 CTAs x 128 threads, stride 640 blocks): many
    rows at small sizes, and a grid stride that is NOT a multiple of 256, so the counter-byte
    cache of the stream kernel is refilled on every block."""
    import aesgcm_b200
    eng = aesgcm_b200.GcmEngine(0, n_cta=5, threads=128)
    yield eng
    eng.close()


U8P = ctypes.POINTER(ctypes.c_uint8)
U64P = ctypes.POINTER(ctypes.c_uint64)


def u8p(a):
    return a.ctypes.data_as(U8P)


def u64p(a):
    return a.ctypes.data_as(U64P)


def rand_bytes(rng, n):
    return rng.integers(0, 256, n, dtype=np.uint8)
