#!/usr/bin/env python
"""Records what the REFERENCE golden model itself (tb/gcm_model.py, executed from
/root/reference) returns for the testbench's calling pattern -- one <=16-byte block per call,
AAD first, get_tag last (tb/gcm_test.py:76-94, tb/gcm_sequencer.py:137-231) -- for seeded
cases in both directions, including the forced-mismatch path (tb/gcm_model.py:47-51).

pycryptodome is not installable here, so `Crypto.Cipher.AES` resolves to the shim under
tests/golden/shims (same six calls, on OpenSSL); the control flow, the list handling and the
tag inversion are the reference's own code.  Build container only:

  python tests/golden/make_model_traces.py   ->  tests/golden/gcm_model_traces.json
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, "/root/reference/tb")

import gcm_model  # the reference file itself


def blocks(b):
    return [b[i:i + 16] for i in range(0, len(b), 16)]


def main():
    rnd = random.Random(42)
    cases = []
    for kb in (16, 24, 32):
        for n, alen in ((0, 20), (48, 28), (100, 0), (16, 16), (333, 7)):
            key = bytes(rnd.getrandbits(8) for _ in range(kb))
            iv = bytes(rnd.getrandbits(8) for _ in range(12))
            aad = bytes(rnd.getrandbits(8) for _ in range(alen))
            pt = bytes(rnd.getrandbits(8) for _ in range(n))
            kd = {'data': key.hex().upper(), 'n_bytes': kb}
            ivd = {'data': iv.hex().upper(), 'n_bytes': 12}
            enc = gcm_model.gcm(kd, ivd, 'enc')
            for a in blocks(aad):
                enc.load_aad(a)
            for p in blocks(pt):
                enc.load_plain_text(p)
            enc.get_tag(b"\0" * 16)           # a DUT tag that does not match: only logged
            ct = b"".join(enc.data_out)
            tag = enc.tag[0]
            rec = {"key": kd, "iv": ivd, "aad": aad.hex(), "pt": pt.hex(),
                   "enc_data_out": [x.hex() for x in enc.data_out], "enc_tag": [t.hex() for t in enc.tag]}
            for label, rx_tag in (("dec_good", tag), ("dec_bad", bytes([tag[0] ^ 1]) + tag[1:])):
                dec = gcm_model.gcm(kd, ivd, 'dec')
                for a in blocks(aad):
                    dec.load_aad(a)
                for c in blocks(ct):
                    dec.load_cipher_text(c)
                dec.get_tag(rx_tag)
                rec[label] = {"rx_tag": rx_tag.hex(), "data_out": [x.hex() for x in dec.data_out],
                              "tag": [t.hex() for t in dec.tag]}
            cases.append(rec)
    with open(os.path.join(HERE, "gcm_model_traces.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_model_traces.py running /root/reference/tb/gcm_model.py "
                                "(pycryptodome calls served by OpenSSL through tests/golden/shims)",
                   "cases": cases}, f, indent=1)
    print("wrote", len(cases), "model traces")


if __name__ == "__main__":
    main()
