"""Drop-in replacement for the reference key-expansion model tb/key_exp.py.

``aes_expand_key(key_hex, size)`` keeps the reference signature and return type
(tb/key_exp.py:118-121: hex string + '128'|'192'|'256' -> flat list of
(Nr+1)*16 ints, stage r at [16r, 16r+16)), but the schedule is computed by the
device kernel k_key_expand (the aes_kexp block, config/config_aes_kexp.py:128-159)
through agcm_key_expand_host.
"""
from .engine import GcmEngine, MODES

_engine = None


def _get_engine(device=0):
    global _engine
    if _engine is None:
        _engine = GcmEngine(device)
    return _engine


class exp_key(object):
    # valid key sizes (tb/key_exp.py:17-19)
    key_size = {'128': 16, '192': 24, '256': 32}

    def aes_expand_key(self, key, size):
        # tb/key_exp.py:79-114: `key` is a hex string, `size` the byte count
        if size not in MODES.values():
            raise ValueError("key size must be 16, 24 or 32 bytes")
        raw = bytes(int(key[2 * i:2 * i + 2], 16) for i in range(size))
        return list(_get_engine().expand_key_host(raw))


def aes_expand_key(key, size):
    # tb/key_exp.py:118-121
    aes = exp_key()
    return aes.aes_expand_key(key, aes.key_size[size])
