"""Stimulus / config layer and wire-format packers: pure host logic,
checked on CPU; the replay is driven with an oracle-backed stand-in for the model here and
with the real CUDA-backed model in tests/test_gpu_parity.py."""
import json
import random

import pytest

from aesgcm_b200 import stimulus as st


class OracleModel:
    """gcm-compatible model (tb/gcm_model.py:5-51 surface) on the CPU oracle -- test only."""

    def __init__(self, key, icb, ed):
        from oracle import cpu_oracle
        self.o = cpu_oracle
        self.ed, self.data_out, self.tag = ed, [], []
        self.key = int(key['data'], 16).to_bytes(key['n_bytes'], 'big')
        self.iv = int(icb['data'], 16).to_bytes(icb['n_bytes'], 'big')
        self.aad, self.text = b"", b""

    def load_aad(self, a):
        self.aad += a

    def _crypt(self, d):
        rk = self.key if len(self.key) > 32 else self.o.key_expand(self.key)
        pos = len(self.text)
        ks = self.o.gctr(rk, self.iv, 2 + pos // 16, bytes(pos % 16 + len(d)))[pos % 16:]
        self.text += d
        self.data_out.append(bytes(a ^ b for a, b in zip(d, ks)))

    load_plain_text = _crypt
    load_cipher_text = _crypt

    def get_tag(self, tag):
        out, t = self.o.gcm_crypt(self.key, self.iv, self.aad, self.text, decrypt=(self.ed == 'dec'))
        assert out == b"".join(self.data_out)
        if self.ed == 'enc':
            self.tag.append(t)
        else:
            self.tag.append(tag if tag == t else bytes(x ^ 0xFF for x in tag))


def test_hex_normalisation_truncates_and_pads():
    cfg = {'aes_mode': '128', 'key': 'ABC', 'iv': 'F' * 30, 'aad': 'EMPTY', 'data': 'EMPTY', 'enc_dec': 'enc', 'seed': 1}
    _, d = st.resolve_config(cfg)
    assert d['key'] == {'n_bytes': 16, 'data': '0' * 29 + 'ABC'}           # right-aligned, zero-padded
    assert d['iv'] == {'n_bytes': 12, 'data': 'F' * 24}                      # truncated to 24 digits
    assert d['aad_n_bytes'] == 0 and d['pt_n_bytes'] == 0
    with pytest.raises(st.TestFailure):
        st.resolve_config(dict(cfg, key='xyz'))
    with pytest.raises(st.TestFailure):
        st.resolve_config(dict(cfg, iv='12g4'))


def test_user_hex_lengths_and_odd_nibble():
    cfg = {'aes_mode': '256', 'key': 'RANDOM', 'iv': 'RANDOM', 'aad': 'ABCDE', 'data': 'A' * 33, 'enc_dec': 'dec', 'seed': 7}
    cfg2, d = st.resolve_config(cfg)
    assert d['aad_n_bytes'] == 3 and d['pt_n_bytes'] == 17
    assert d['key']['n_bytes'] == 32 and len(d['key']['data']) == 64 and len(d['iv']['data']) == 24
    assert d['delays'] & 0x4 == 0                                           # no AAD/CT overlap when decrypting
    assert st.split_words('ABCDE', 3, None) == [bytes.fromhex('ABCDE0')]    # odd nibble padded on the right
    w = st.split_words('A' * 33, 17, None)
    assert w == [bytes.fromhex('A' * 32), bytes.fromhex('A0')]
    assert st.split_words('EMPTY', 0, None) == []
    rw = st.split_words('RANDOM', 37, random.Random(3))
    assert [len(x) for x in rw] == [16, 16, 5]
    # same seed, same stimulus (tb/tmp/<seed>.json replay)
    assert st.resolve_config(cfg) == st.resolve_config(cfg)


def test_key_pin_encodings():
    key = {'data': 'AD7A2BD03EAC835A6F620FDCB506B345', 'n_bytes': 16}
    w = st.pack_key_word(key)
    assert w >> 128 == int(key['data'], 16) and w & ((1 << 128) - 1) == 0   # left-aligned in 256 bits
    assert st.key_mode_val('128') == 0b0100 and st.key_mode_val('192') == 0b0110 and st.key_mode_val('256') == 0b0111
    assert st.unpack_key_word(w, '128') == bytes.fromhex(key['data'])
    k24 = {'data': '8E73B0F7DA0E6452C810F32B809079E562F8EAD2522C6B7B', 'n_bytes': 24}
    assert st.unpack_key_word(st.pack_key_word(k24), '192') == bytes.fromhex(k24['data'])
    exp = bytes(range(176))
    writes = st.pack_pre_expanded_key(exp)
    assert [v for v, _ in writes] == list(range(1, 12))
    assert all(wd & ((1 << 128) - 1) == 0 for _, wd in writes)               # stage in the top 128 bits
    assert st.unpack_pre_expanded_key(writes) == exp
    with pytest.raises(ValueError):
        st.pack_pre_expanded_key(bytes(100))


def test_data_word_packing():
    # the example in tb/gcm_driver.py:112-117
    word, bval = st.pack_data_word(bytes.fromhex('756A9E2C1904DF026D35'))
    assert bval == 0xFFC0 and word == 0x756A9E2C1904DF026D35000000000000
    assert st.unpack_data_word(word, bval) == bytes.fromhex('756A9E2C1904DF026D35')
    for n in range(1, 17):
        b = bytes(range(1, n + 1))
        w, v = st.pack_data_word(b)
        assert bin(v).count('1') == n and st.unpack_data_word(w, v) == b
    assert st.unpack_data_word(0, 0) == b""
    with pytest.raises(ValueError):
        st.pack_data_word(b"")


def test_replay_with_oracle_backed_model(tmp_path, oracle):
    # README.md:251 command line as a saved config file
    cfg = {'seed': 42, 'aes_mode': '128', 'key': 'AD7A2BD03EAC835A6F620FDCB506B345', 'iv': '12153524C0895E81B2C28465',
           'aad': 'D609B1F056637A0D46DF998D88E52E00B2C2846512153524C0895E81',
           'data': '08000F101112131415161718191A1B1C1D1E1F202122232425262728292A2B2C2D2E2F303132333435363738393A0002',
           'enc_dec': 'enc', 'max_n_byte': 4095, 'test_size': 'short'}
    p = tmp_path / "42.json"
    p.write_text(json.dumps(cfg))
    r = st.replay(st.load_config(str(p)), OracleModel)
    assert b"".join(r['ct_words']).hex().upper().startswith('701AFA1CC039C0D765128A665DAB6924')
    assert r['tag'].hex().upper() == '4F8D55E7D3F06FD5A13C0C29B9D5B880'
    # random stimulus, decrypt direction, pre-expanded key flow
    for seed in range(5):
        c = {'seed': seed, 'aes_mode': 'ALL', 'key': 'RANDOM', 'iv': 'RANDOM', 'aad': 'RANDOM', 'data': 'RANDOM',
             'enc_dec': 'dec', 'max_n_byte': 300}
        r = st.replay(c, OracleModel, pre_expanded=(seed % 2 == 0),
                      expand_key=lambda k, s: list(oracle.key_expand(bytes.fromhex(k))))
        assert b"".join(r['dec_words']) == b"".join(r['pt_words'])
        assert r['dec_tag'] == r['tag']


def _traces():
    import os
    from conftest import GOLDEN
    with open(os.path.join(GOLDEN, "stimulus_traces.json")) as f:
        return json.load(f)["cases"]


def test_resolve_config_matches_reference_gcm_gctr():
    """tests/golden/stimulus_traces.json was recorded from the reference's own tb/gcm_gctr.py
    (config_data + encrypt_data driven alone): same seed -> same key / IV / counts / delays and
    the same AAD / data word lists."""
    cases = _traces()
    assert len(cases) == 6
    for c in cases:
        rng = random.Random(c["config_in"]["seed"])
        cfg, d = st.resolve_config(c["config_in"], rng)
        assert d["key"] == c["data"]["key"] and d["iv"] == c["data"]["iv"], c["config_in"]["seed"]
        assert d["aad_n_bytes"] == c["data"]["aad_n_bytes"] and d["pt_n_bytes"] == c["data"]["pt_n_bytes"]
        assert d["delays"] == c["data"]["delays"]
        assert cfg["aes_mode"] == c["config_out"]["aes_mode"] and cfg["key"] == c["config_out"]["key"]
        aad_words = st.split_words(cfg["aad"], d["aad_n_bytes"], rng)
        pt_words = st.split_words(cfg["data"], d["pt_n_bytes"], rng)
        assert [w.hex() for w in aad_words] == c["aad_words"]
        assert [w.hex() for w in pt_words] == c["pt_words"]


def test_key_pin_encodings_match_reference_gcm_gctr(oracle):
    """load_key / load_pre_exp_key pin writes of the reference (tb/gcm_gctr.py:144-214)."""
    for c in _traces():
        key = c["data"]["key"]
        mode = c["config_out"]["aes_mode"]
        assert c["load_key"] == [[st.key_mode_val(mode), "%064X" % st.pack_key_word(key)]]
        exp = oracle.key_expand(bytes.fromhex(key["data"]))
        want = [[v, "%064X" % w] for v, w in st.pack_pre_expanded_key(exp)]
        assert c["load_pre_exp_key"] == want
