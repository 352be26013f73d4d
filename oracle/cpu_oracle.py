"""ctypes binding of the CPU oracle (oracle/gcm_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package never
imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "gcm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return _SO


_lib = None
_u8p = ctypes.POINTER(ctypes.c_uint8)
_u64p = ctypes.POINTER(ctypes.c_uint64)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_key_expand.restype = ctypes.c_int
        _lib.oracle_gcm_crypt.restype = ctypes.c_int
        _lib.oracle_gcm_batch.restype = ctypes.c_int
        _lib.oracle_gcm_stream_mt.restype = ctypes.c_int
    return _lib


def _buf(b):
    """bytes / bytearray / np.uint8 array -> (ctypes pointer, keepalive)."""
    if isinstance(b, np.ndarray):
        a = np.ascontiguousarray(b, dtype=np.uint8)
    else:
        a = np.frombuffer(bytes(b), dtype=np.uint8)
    if a.size == 0:
        a = np.zeros(1, dtype=np.uint8)
    return a.ctypes.data_as(_u8p), a


def sbox_table():
    out = np.zeros(256, dtype=np.uint8)
    lib().oracle_sbox_table(out.ctypes.data_as(_u8p))
    return out


def key_expand(key: bytes) -> bytes:
    """tb/key_exp.py:79-114 restated; returns 176/208/240 bytes."""
    out = np.zeros(240, dtype=np.uint8)
    kp, _k = _buf(key)
    nr = lib().oracle_key_expand(kp, len(key), out.ctypes.data_as(_u8p))
    if nr < 0:
        raise ValueError("bad key length")
    return out[: (nr + 1) * 16].tobytes()


def aes_encrypt_block(rk: bytes, block: bytes) -> bytes:
    out = np.zeros(16, dtype=np.uint8)
    rp, _r = _buf(rk)
    bp, _b = _buf(block)
    lib().oracle_aes_encrypt_block(rp, len(rk) // 16 - 1, bp, out.ctypes.data_as(_u8p))
    return out.tobytes()


def gfmul(h: bytes, x: bytes) -> bytes:
    out = np.zeros(16, dtype=np.uint8)
    hp, _h = _buf(h)
    xp, _x = _buf(x)
    lib().oracle_gfmul(hp, xp, out.ctypes.data_as(_u8p))
    return out.tobytes()


def gf_pow(h: bytes, e: int) -> bytes:
    out = np.zeros(16, dtype=np.uint8)
    hp, _h = _buf(h)
    lib().oracle_gf_pow(hp, ctypes.c_uint64(e), out.ctypes.data_as(_u8p))
    return out.tobytes()


def ghash_absorb(h: bytes, data: bytes, y: bytes = b"\0" * 16) -> bytes:
    yy = np.frombuffer(bytes(y), dtype=np.uint8).copy()
    hp, _h = _buf(h)
    dp, _d = _buf(data)
    lib().oracle_ghash_absorb(hp, dp, ctypes.c_uint64(len(data)), yy.ctypes.data_as(_u8p))
    return yy.tobytes()


def h_ej0(rk: bytes, iv: bytes):
    h = np.zeros(16, dtype=np.uint8)
    e = np.zeros(16, dtype=np.uint8)
    rp, _r = _buf(rk)
    ip, _i = _buf(iv)
    lib().oracle_h_ej0(rp, len(rk) // 16 - 1, ip, h.ctypes.data_as(_u8p), e.ctypes.data_as(_u8p))
    return h.tobytes(), e.tobytes()


def gctr(rk: bytes, iv: bytes, first_ctr: int, data: bytes) -> bytes:
    out = np.zeros(max(len(data), 1), dtype=np.uint8)
    rp, _r = _buf(rk)
    ip, _i = _buf(iv)
    dp, _d = _buf(data)
    lib().oracle_gctr(rp, len(rk) // 16 - 1, ip, ctypes.c_uint32(first_ctr & 0xFFFFFFFF), dp,
                      ctypes.c_uint64(len(data)), out.ctypes.data_as(_u8p))
    return out[: len(data)].tobytes()


def gcm_crypt(key: bytes, iv: bytes, aad: bytes, data, decrypt: bool = False, threads: int = 1):
    """Whole message; key is raw (16/24/32) or pre-expanded (176/208/240).
    Returns (out_bytes, computed_tag)."""
    n = len(data)
    out = np.zeros(max(n, 1), dtype=np.uint8)
    tag = np.zeros(16, dtype=np.uint8)
    kp, _k = _buf(key)
    ip, _i = _buf(iv)
    ap, _a = _buf(aad)
    dp, _d = _buf(data)
    if threads > 1:
        rc = lib().oracle_gcm_stream_mt(kp, len(key), ip, ap, ctypes.c_uint64(len(aad)), dp, ctypes.c_uint64(n),
                                        int(decrypt), out.ctypes.data_as(_u8p), tag.ctypes.data_as(_u8p), threads)
    else:
        rc = lib().oracle_gcm_crypt(kp, len(key), ip, ap, ctypes.c_uint64(len(aad)), dp, ctypes.c_uint64(n),
                                    int(decrypt), out.ctypes.data_as(_u8p), tag.ctypes.data_as(_u8p))
    if rc:
        raise ValueError("oracle_gcm_crypt rc=%d" % rc)
    return out[:n].tobytes(), tag.tobytes()


def _pad16(b: bytes) -> bytes:
    return bytes(b) + b"\0" * (-len(b) % 16)


def gcm_crypt_any_iv(key: bytes, iv: bytes, aad: bytes, data: bytes, decrypt: bool = False):
    """AES-GCM with an IV of any length, composed from the primitives above.  The reference fixes the
    IV at 96 bits (src/gcm_pkg.vhd:17; J0 = IV || 0^31 || 1, src/aes_icb.vhd:34); for any other
    length SP 800-38D 7.1 step 2 sets J0 = GHASH_H(IV || 0^(s+64) || [len(IV)]_64), the counter then
    runs inc32 from J0 (src/aes_icb.vhd:99-100 semantics on J0's last word) and the tag is
    GHASH_H(A, C) xor E_K(J0) (src/gcm_ghash.vhd:257,293).  Returns (out_bytes, computed_tag)."""
    key, iv, aad, data = bytes(key), bytes(iv), bytes(aad), bytes(data)
    if len(iv) == 12:
        return gcm_crypt(key, iv, aad, data, decrypt=decrypt)
    rk = key if len(key) in (176, 208, 240) else key_expand(key)
    h = aes_encrypt_block(rk, b"\0" * 16)
    j0 = ghash_absorb(h, _pad16(iv) + b"\0" * 8 + (8 * len(iv)).to_bytes(8, "big"))
    c0 = int.from_bytes(j0[12:], "big")
    out = gctr(rk, j0[:12], c0 + 1, data)
    ct = data if decrypt else out
    s = ghash_absorb(h, _pad16(aad) + _pad16(ct) + (8 * len(aad)).to_bytes(8, "big") + (8 * len(ct)).to_bytes(8, "big"))
    ej0 = aes_encrypt_block(rk, j0)
    return out, bytes(a ^ b for a, b in zip(s, ej0))


def gcm_batch(keys, key_len, shared_key, ivs, aad, aad_off, data, in_off, decrypt=False, threads=1):
    """numpy arrays in, (out, tags) numpy arrays back.  keys: uint8 (n*key_len or key_len)."""
    n = len(in_off) - 1
    keys = np.ascontiguousarray(keys, dtype=np.uint8)
    ivs = np.ascontiguousarray(ivs, dtype=np.uint8)
    data = np.ascontiguousarray(data, dtype=np.uint8)
    in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
    out = np.zeros(max(data.size, 1), dtype=np.uint8)
    tags = np.zeros(16 * max(n, 1), dtype=np.uint8)
    if aad is None or aad_off is None:
        ap, aop = None, None
    else:
        aad = np.ascontiguousarray(aad, dtype=np.uint8)
        if aad.size == 0:
            aad = np.zeros(1, dtype=np.uint8)
        aad_off = np.ascontiguousarray(aad_off, dtype=np.uint64)
        ap, aop = aad.ctypes.data_as(_u8p), aad_off.ctypes.data_as(_u64p)
    dptr = data.ctypes.data_as(_u8p) if data.size else out.ctypes.data_as(_u8p)
    rc = lib().oracle_gcm_batch(keys.ctypes.data_as(_u8p), int(key_len), ctypes.c_size_t(0 if shared_key else key_len),
                                ivs.ctypes.data_as(_u8p), ap, aop, dptr, in_off.ctypes.data_as(_u64p),
                                int(decrypt), out.ctypes.data_as(_u8p), tags.ctypes.data_as(_u8p),
                                ctypes.c_size_t(n), int(threads))
    if rc:
        raise ValueError("oracle_gcm_batch rc=%d" % rc)
    return out[: data.size], tags[: 16 * n]
