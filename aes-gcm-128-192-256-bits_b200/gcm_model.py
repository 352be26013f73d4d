"""Drop-in replacement for the reference golden model tb/gcm_model.py.

Same class name, constructor arguments, methods, attributes and error behaviour
(tb/gcm_model.py:5-51), so that tb/gcm_test.py:45,76-94 can bind its monitors and
scoreboard to it unchanged -- but AES and GHASH run on the B200 through
libaesgcm_b200.so instead of pycryptodome.

Call pattern of the testbench: synchronous callbacks, one <=16-byte
block per call, AAD blocks first, and the expected output must be in
``data_out`` as soon as ``load_plain_text`` returns.  A kernel launch per 16 bytes
would be all latency, so the adapter PREFETCHES keystream: the device runs GCTR
(agcm_gctr, src/gcm_gctr.vhd:150) over a zero window starting at the current
counter, and each callback XORs its bytes against that window.  The tag is
produced at ``get_tag`` by one fused GCTR+GHASH pass of the device over the
accumulated AAD and text (agcm_stream_crypt_host); that pass also re-derives the
output text, which is checked against what the callbacks handed out.
"""
import logging

try:  # the reference imports cocotb's logger (tb/gcm_model.py:2); optional here
    from cocotb import log  # type: ignore
except Exception:  # pragma: no cover - cocotb is not a dependency of this repo
    log = logging.getLogger("aesgcm_b200.gcm_model")

from .engine import GcmEngine

_KS_WINDOW = 1 << 16  # bytes of keystream fetched per device call

# One engine per device for all model instances (a testbench creates a model per packet; an engine
# owns a context, its streams and ~200 MiB of staging buffers).  Callers that drive several models
# concurrently pass their own `engine`.
_engines = {}


def shared_engine(device=0):
    eng = _engines.get(device)
    if eng is None:
        eng = _engines[device] = GcmEngine(device)
    return eng


def close_shared_engines():
    for eng in _engines.values():
        eng.close()
    _engines.clear()


class gcm:
    def __init__(self, key, icb, ed, device=0, engine=None, prefetch=True):
        # tb/gcm_model.py:8-18
        self.ed = ed
        self.data_out = []
        self.tag = []

        _key = int(key['data'], 16).to_bytes(key['n_bytes'], byteorder='big')
        _icb = int(icb['data'], 16).to_bytes(icb['n_bytes'], byteorder='big')
        if len(_icb) == 0:
            raise ValueError("empty IV")
        # The IP fixes the IV at 96 bits (src/gcm_pkg.vhd:17); pycryptodome's AES.new(nonce=...), which
        # the reference model hands icb['n_bytes'] bytes to (tb/gcm_model.py:14-18), takes any length:
        # so does the engine (J0 by GHASH on the device, SP 800-38D 7.1).
        self.model = engine if engine is not None else shared_engine(device)
        self._key = _key                  # raw 16/24/32 B, or 176/208/240 B pre-expanded stages
        self.model.set_key(_key)
        self._iv = _icb
        self._aad = bytearray()
        self._text = bytearray()          # everything fed to load_plain_text / load_cipher_text
        self._out = bytearray()           # everything handed out
        self._ks = b""                    # keystream window ...
        self._ks_base = 0                 # ... covering message bytes [_ks_base, _ks_base + len(_ks))
        self._zeros = None
        self._ksbuf = None
        # prefetch=False: no keystream window; every callback is its own device GCTR call
        # (H2D of the block, agcm_gctr, D2H): the XOR itself then also happens on the GPU
        self._prefetch = bool(prefetch)

    # ------------------------------------------------------------------
    def _bind(self):
        # the engine may be shared with other model instances: (re)load this model's key; loading the
        # key that is already loaded returns at once (H is kept, src/gcm_ghash.vhd:123)
        self.model.set_key(self._key)

    def _keystream(self, pos, n):
        """n bytes of keystream for message byte offset pos (device GCTR over zeros)."""
        out = bytearray()
        while n:
            off = pos - self._ks_base
            if off < 0 or off >= len(self._ks):
                import torch
                dev = "cuda:%d" % self.model.device
                if self._zeros is None:
                    self._zeros = torch.zeros(_KS_WINDOW, dtype=torch.uint8, device=dev)
                    self._ksbuf = torch.empty(_KS_WINDOW, dtype=torch.uint8, device=dev)
                first_block = pos // 16
                self._bind()
                with torch.cuda.device(self.model.device):
                    self.model.gctr_device(self._iv, first_block, self._zeros, self._ksbuf)
                    self._ks = self._ksbuf.cpu().numpy().tobytes()
                self._ks_base = first_block * 16
                off = pos - self._ks_base
            take = min(n, len(self._ks) - off)
            out += self._ks[off:off + take]
            pos += take
            n -= take
        return bytes(out)

    def _crypt_on_device(self, pos, data):
        import numpy as np
        import torch
        lead = pos % 16                       # keep the counter block-aligned for a mid-block call
        buf = np.zeros(lead + len(data), dtype=np.uint8)
        buf[lead:] = np.frombuffer(data, dtype=np.uint8)
        self._bind()
        with torch.cuda.device(self.model.device):
            d = torch.from_numpy(buf).cuda()
            o = torch.empty_like(d)
            self.model.gctr_device(self._iv, pos // 16, d, o)
            return o.cpu().numpy()[lead:].tobytes()

    def _crypt(self, data):
        data = bytes(data)
        if not data:
            res = b""
        elif self._prefetch:
            ks = self._keystream(len(self._text), len(data))
            res = (int.from_bytes(data, 'big') ^ int.from_bytes(ks, 'big')).to_bytes(len(data), 'big')
        else:
            res = self._crypt_on_device(len(self._text), data)
        self._text += data
        self._out += res
        return res

    # ------------------------------------------------------------------
    def load_aad(self, aad):
        # tb/gcm_model.py:21-22
        self._aad += bytes(aad)

    def load_plain_text(self, pt):
        # tb/gcm_model.py:25-26
        self.data_out.append(self._crypt(pt))

    def load_cipher_text(self, ct):
        # tb/gcm_model.py:29-30
        self.data_out.append(self._crypt(ct))

    # ------------------------------------------------------------------
    def get_tag(self, tag):
        """Close the message (tb/gcm_model.py:33-51 semantics).

        'enc': the engine's tag is appended to ``self.tag`` and compared with the DUT's `tag`
        (a mismatch is only logged -- the scoreboard does the failing).
        'dec': the engine verifies `tag`; on success the received tag is appended, on failure
        its bitwise complement, so that the scoreboard is guaranteed to see a mismatch."""
        aad, text = bytes(self._aad), bytes(self._text)
        self._bind()
        if self.ed == 'enc':
            out, computed = self.model.encrypt(self._iv, aad, text)
            authentic = True
        else:
            out, authentic = self.model.decrypt(self._iv, aad, text, bytes(tag), raise_on_fail=False)
            computed = None
        if out != bytes(self._out):
            raise RuntimeError("fused pass and prefetched keystream disagree")
        if self.ed == 'enc':
            self.tag.append(computed)
            log.info("model tag %s", computed.hex().upper())
            if tag is not None and tag != computed:
                log.error("tag mismatch: DUT %s", bytes(tag).hex().upper())
        elif authentic:
            self.tag.append(tag)
            log.info("tag verified: message is authentic")
        else:
            log.error("tag verification failed (wrong key/IV or corrupted message)")
            flipped = bytes(b ^ 0xFF for b in bytes(tag))
            self.tag.append(flipped)
