#!/usr/bin/env python
"""Regenerates the committed golden fixtures.  Run in the build container only:
it imports the reference's own tb/key_exp.py from /root/reference (absent on the
GPU box; the tests read only the JSON files this script writes).

  python tests/golden/make_golden.py

Writes:
  tests/golden/key_exp_vectors.json -- outputs of the REFERENCE key schedule
      (tb/key_exp.py:118 aes_expand_key) for the FIPS-197 App. A keys and
      seeded random keys, plus the reference S-box list (tb/key_exp.py:23-54).
  tests/golden/kat_vectors.json -- AES-GCM known-answer vectors (96-bit IV):
      the McGrew-Viega / SP 800-38D test cases and the two IEEE 802.1AE vectors
      whose INPUTS the reference README quotes (README.md:249-258).  The
      reference stores no expected outputs; the CT/TAG here are the published
      ones (tags cross-checked below against OpenSSL via `cryptography`, which
      computes the same function as the pycryptodome call at
      tb/gcm_model.py:18).
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_TB = "/root/reference/tb"


def main():
    sys.path.insert(0, REF_TB)
    import key_exp  # the reference module itself (no third-party imports)

    rnd = random.Random(20261017)
    cases = []
    fips = {
        "128": "2b7e151628aed2a6abf7158809cf4f3c",
        "192": "8e73b0f7da0e6452c810f32b809079e562f8ead2522c6b7b",
        "256": "603deb1015ca71be2b73aef0857d77811f352c073b6108d72d9810a30914dff4",
    }
    for size, k in fips.items():
        cases.append({"size": size, "key": k.upper(), "src": "FIPS-197 App. A"})
    for size, nb in (("128", 16), ("192", 24), ("256", 32)):
        for _ in range(24):
            k = bytes(rnd.getrandbits(8) for _ in range(nb)).hex().upper()
            cases.append({"size": size, "key": k, "src": "random"})
        cases.append({"size": size, "key": "00" * nb, "src": "zeros"})
        cases.append({"size": size, "key": "FF" * nb, "src": "ones"})
    for c in cases:
        exp = key_exp.aes_expand_key(c["key"], c["size"])
        c["expanded"] = bytes(exp).hex()
    out = {
        "generator": "tests/golden/make_golden.py importing /root/reference/tb/key_exp.py",
        "sbox": bytes(key_exp.exp_key.sbox).hex(),
        "cases": cases,
    }
    with open(os.path.join(HERE, "key_exp_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)

    # ---- AES-GCM KATs ------------------------------------------------------
    K128 = "feffe9928665731c6d6a8f9467308308"
    K192 = K128 + "feffe9928665731c"
    K256 = K128 + K128
    IV = "cafebabefacedbaddecaf888"
    P64 = ("d9313225f88406e5a55909c5aff5269a86a7a9531534f7da2e4c303d8a318a72"
           "1c3c0c95956809532fcf0e2449a6b525b16aedf5aa0de657ba637b391aafd255")
    P60 = P64[:120]
    A20 = "feedfacedeadbeeffeedfacedeadbeefabaddad2"
    Z = "00" * 16
    kats = [
        ("MV-TC1", "00" * 16, "00" * 12, "", "", "58e2fccefa7e3061367f1d57a4e7455a"),
        ("MV-TC2", "00" * 16, "00" * 12, Z, "", "ab6e47d42cec13bdf53a67b21257bddf"),
        ("MV-TC3", K128, IV, P64, "", "4d5c2af327cd64a62cf35abd2ba6fab4"),
        ("MV-TC4", K128, IV, P60, A20, "5bc94fbc3221a5db94fae95ae7121a47"),
        ("MV-TC7", "00" * 24, "00" * 12, "", "", "cd33b28ac773f74ba00ed1f312572435"),
        ("MV-TC8", "00" * 24, "00" * 12, Z, "", "2ff58d80033927ab8ef4d4587514f0fb"),
        ("MV-TC9", K192, IV, P64, "", "9924a7c8587336bfb118024db8674a14"),
        ("MV-TC10", K192, IV, P60, A20, "2519498e80f1478f37ba55bd6d27618c"),
        ("MV-TC13", "00" * 32, "00" * 12, "", "", "530f8afbc74536b9a963b4f1c4cb738b"),
        ("MV-TC14", "00" * 32, "00" * 12, Z, "", "d0d1c8a799996bf0265b98b5d48ab919"),
        ("MV-TC15", K256, IV, P64, "", "b094dac5d93471bdec1a502270e3cc6c"),
        ("MV-TC16", K256, IV, P60, A20, "76fc6ece0f4e1768cddf8853bb2d551b"),
        ("802.1AE-GCM-AES-128-60B-enc (README.md:251)",
         "AD7A2BD03EAC835A6F620FDCB506B345", "12153524C0895E81B2C28465",
         "08000F101112131415161718191A1B1C1D1E1F202122232425262728292A2B2C2D2E2F303132333435363738393A0002",
         "D609B1F056637A0D46DF998D88E52E00B2C2846512153524C0895E81",
         "4F8D55E7D3F06FD5A13C0C29B9D5B880"),
        ("802.1AE-GCM-AES-256-65B-auth (README.md:257)",
         "691D3EE909D7F54167FD1CA0B5D769081F2BDE1AEE655FDBAB80BD5295AE6BE7", "F0761E8DCD3D000176D457ED",
         "",
         "E20106D7CD0DF0761E8DCD3D88E5400076D457ED08000F101112131415161718191A1B1C1D1E1F202122232425262728292A2B2C2D2E2F303132333435363738393A0003",
         "35217C774BBC31B63166BCF9D4ABED07"),
    ]
    from cryptography.hazmat.primitives.ciphers.aead import AESGCM
    recs = []
    for name, k, iv, pt, aad, tag in kats:
        kb, ivb, ptb, ab = bytes.fromhex(k), bytes.fromhex(iv), bytes.fromhex(pt), bytes.fromhex(aad)
        full = AESGCM(kb).encrypt(ivb, ptb, ab)
        ct, t = full[:-16], full[-16:]
        assert t.hex() == tag.lower(), name          # published tag == OpenSSL
        recs.append({"name": name, "key": k.lower(), "iv": iv.lower(), "pt": pt.lower(), "aad": aad.lower(),
                     "ct": ct.hex(), "tag": tag.lower()})
    # published CT spot checks
    assert recs[2]["ct"].startswith("42831ec2217774244b7221b784d0d49c")
    assert recs[12]["ct"] == ("701afa1cc039c0d765128a665dab69243899bf7318ccdc81c9931da17fbe8edd"
                              "7d17cb8b4c26fc81e3284f2b7fba713d")
    with open(os.path.join(HERE, "kat_vectors.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "vectors": recs}, f, indent=1)
    # ---- random cases computed by OpenSSL (stand-in for the pycryptodome call of tb/gcm_model.py:18)
    rv = []
    for kb in (16, 24, 32):
        for n, alen in ((0, 0), (1, 0), (15, 7), (16, 16), (17, 20), (33, 64), (100, 1), (255, 33), (256, 0), (511, 100)):
            key = bytes(rnd.getrandbits(8) for _ in range(kb))
            iv = bytes(rnd.getrandbits(8) for _ in range(12))
            aad = bytes(rnd.getrandbits(8) for _ in range(alen))
            pt = bytes(rnd.getrandbits(8) for _ in range(n))
            full = AESGCM(key).encrypt(iv, pt, aad)
            rv.append({"key": key.hex(), "iv": iv.hex(), "aad": aad.hex(), "pt": pt.hex(), "ct": full[:-16].hex(),
                       "tag": full[-16:].hex()})
    with open(os.path.join(HERE, "openssl_random_vectors.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py (cryptography/OpenSSL AESGCM)", "vectors": rv}, f, indent=1)
    print("wrote", len(cases), "key_exp cases,", len(recs), "KATs and", len(rv), "OpenSSL random vectors")


if __name__ == "__main__":
    main()
