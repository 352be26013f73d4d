from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes
from cryptography.exceptions import InvalidTag

MODE_GCM = 11  # pycryptodome's constant


class _Gcm:
    """Lazy direction like pycryptodome: the first encrypt()/decrypt() call fixes it."""

    def __init__(self, key, nonce):
        self._key, self._nonce = bytes(key), bytes(nonce)
        self._ctx = None
        self._aad = []
        self._dec = None

    def _start(self, dec):
        if self._ctx is None:
            c = Cipher(algorithms.AES(self._key), modes.GCM(self._nonce, min_tag_length=16))
            self._ctx = c.decryptor() if dec else c.encryptor()
            self._dec = dec
            for a in self._aad:
                self._ctx.authenticate_additional_data(a)

    def update(self, aad):
        if self._ctx is None:
            self._aad.append(bytes(aad))
        else:
            self._ctx.authenticate_additional_data(bytes(aad))
        return self

    def encrypt(self, pt):
        self._start(False)
        return self._ctx.update(bytes(pt))

    def decrypt(self, ct):
        self._start(True)
        return self._ctx.update(bytes(ct))

    def digest(self):
        self._start(False)
        self._ctx.finalize()
        return self._ctx.tag

    def verify(self, tag):
        self._start(True)
        try:
            self._ctx.finalize_with_tag(bytes(tag))
        except InvalidTag:
            raise ValueError("MAC check failed")


def new(key, mode=None, nonce=None, **kw):
    assert mode == MODE_GCM and nonce is not None
    return _Gcm(key, nonce)
