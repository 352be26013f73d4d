"""B200-native AES-GCM engine (sm_100a CUDA library + thin Python host).

Drop-in surface for the golden-model side of BLu85/AES-GCM-128-192-256-bits:
``gcm_model.gcm`` (tb/gcm_model.py) and ``key_exp.aes_expand_key`` (tb/key_exp.py),
plus whole-message, batched and sharded entry points over the C ABI in
include/aesgcm_b200.h.  Import as ``aesgcm_b200``.
"""
from . import _lib
from .engine import AuthenticationError, GcmEngine
from ._lib import AgcmError

__all__ = ["GcmEngine", "AuthenticationError", "AgcmError", "_lib"]
