"""Stand-in for cocotb, used ONLY by tests/golden/make_*_traces.py to execute reference files
from /root/reference/tb: the logger (tb/gcm_model.py:2) and a pass-through `coroutine`
decorator (tb/gcm_gctr.py uses generator-style coroutines, driven here by plain iteration)."""
import logging

log = logging.getLogger("cocotb-shim")
RANDOM_SEED = 0


def coroutine(fn):
    return fn
