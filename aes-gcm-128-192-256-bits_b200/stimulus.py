"""Stimulus / config layer and wire-format packers of the reference testbench, as
pure functions, so that recorded reference test
configurations (`tb/tmp/<seed>.json`, config/gcm_utils.py:258) replay against the
engine with no cocotb, GHDL or DUT.

Nothing here touches AES or GHASH: it is data formatting only.

* `resolve_config`   -- tb/gcm_gctr.py:229-332 `config_data`: key / IV hex normalisation
  (right-aligned, zero-padded, TRUNCATED), `RANDOM` / `EMPTY` / user hex, byte counts.
* `split_words`      -- tb/gcm_gctr.py:337-437 `encrypt_data`: 32-nibble words, odd nibble
  padded with '0' on the right (:383-384,426-427), short last word.
* `pack_key_word` / `pack_pre_expanded_key` -- tb/gcm_gctr.py:144-214: the 256-bit
  left-aligned key word + 4-bit val, and the Nr+1 stage writes (val = index+1, stage in the
  top 128 bits; config/config_aes_kprexp.py:66-94).
* `pack_data_word` / `unpack_data_word` -- tb/gcm_driver.py:108-129 and
  src/gcm_ghash.vhd:228-244: 128-bit word + 16-bit byte-valid, left-aligned.
* `replay`           -- tb/gcm_test.py:36-140 minus the DUT: drive the model surface with a
  resolved configuration and return what the scoreboard would have compared.
"""
import json
import random
import re

RANDOM_PARAM = 'RANDOM'   # tb/gcm_gctr.py:19
EMPTY_PARAM = 'EMPTY'     # tb/gcm_gctr.py:20
AES_KEY_TYPES = ['128', '192', '256']
AES_KEY_256_WIDTH = 256
_HEX = re.compile(r"^[0-9A-F]+$")


class TestFailure(Exception):
    """Stand-in for cocotb.result.TestFailure (raised at tb/gcm_gctr.py:262,273,297,313)."""


def load_config(path):
    """A `tb/tmp/<seed>.json` file written by config/gcm_utils.py:248-263."""
    with open(path) as f:
        return dict(json.load(f))


def _norm_hex(value, n_bytes, what):
    # tb/gcm_gctr.py:258-262,269-273: must be upper-case hex; right-aligned, zero-padded to
    # 2*n_bytes digits and TRUNCATED to the first 2*n_bytes digits ('{:0>{w}.{w}}')
    if _HEX.fullmatch(value) is None:
        raise TestFailure("%s is not an hexadecimal number" % what)
    return '{:0>{width}.{max}}'.format(value, width=2 * n_bytes, max=2 * n_bytes)


def resolve_config(config, rng=None):
    """tb/gcm_gctr.py:229-332.  Returns (config', data) where data holds
    key{'data','n_bytes'}, iv{'data','n_bytes'}, aad_n_bytes, pt_n_bytes, delays.
    `rng`: a random.Random; defaults to one seeded with config['seed'] (the reference seeds
    Python's RNG with RANDOM_SEED through cocotb)."""
    cfg = dict(config)
    if rng is None:
        rng = random.Random(cfg.get('seed', 0))
    if cfg.get('aes_mode', '128') == "ALL":
        cfg['aes_mode'] = rng.choice(AES_KEY_TYPES)
    mode = str(cfg.get('aes_mode', '128'))
    if mode not in AES_KEY_TYPES:
        raise TestFailure("bad aes_mode %r" % mode)
    key = {'n_bytes': {'128': 16, '192': 24, '256': 32}[mode]}
    iv = {'n_bytes': 12}
    if cfg.get('iv', RANDOM_PARAM) == RANDOM_PARAM:
        # the reference draws randint(0, 16): 16 prints as "10" (SURVEY 4 "stimulus quirk")
        cfg['iv'] = ''.join(['{:X}'.format(rng.randint(0, 16)) for _ in range(24)])
    iv['data'] = _norm_hex(cfg['iv'], iv['n_bytes'], "IV")
    if cfg.get('key', RANDOM_PARAM) == RANDOM_PARAM:
        cfg['key'] = ''.join(['{:X}'.format(rng.randint(0, 16)) for _ in range(64)])
    key['data'] = _norm_hex(cfg['key'], key['n_bytes'], "Key")
    data = {'iv': iv, 'key': key}
    max_n = int(cfg.get('max_n_byte', 2 ** 12 - 1))
    counts = {}
    for name in ('aad_n_bytes', 'pt_n_bytes'):
        counts[name] = int(rng.betavariate(.1, .1) * max_n)     # tb/gcm_gctr.py:280
    for name, field, what in (('aad_n_bytes', 'aad', "AAD data"), ('pt_n_bytes', 'data', "Data")):
        v = cfg.get(field, RANDOM_PARAM)
        if v == EMPTY_PARAM:
            counts[name] = 0
        elif v != RANDOM_PARAM:
            if _HEX.fullmatch(v) is None:
                raise TestFailure("%s is not an hexadecimal number" % what)
            counts[name] = (len(v) + 1) >> 1                     # nibbles -> bytes, odd rounds up
    data.update(counts)
    data['delays'] = rng.randint(0, 31)
    if cfg.get('enc_dec', 'enc') == 'dec':
        data['delays'] &= ~(1 << 2)                              # tb/gcm_gctr.py:331-332
    return cfg, data


def split_words(value, n_bytes, rng):
    """The list of <=16-byte transactions of tb/gcm_gctr.py:337-437 for one of AAD / data.
    value: RANDOM (then n_bytes random bytes, last word short), EMPTY, or user hex."""
    if value == EMPTY_PARAM or (value == RANDOM_PARAM and n_bytes == 0):
        return []
    if value == RANDOM_PARAM:
        words = [bytes.fromhex('{:032X}'.format(rng.randint(0, (2 ** 128) - 1))) for _ in range(n_bytes >> 4)]
        rem = n_bytes & 0xF
        if rem:
            words.append(bytes.fromhex('{:0{width}X}'.format(rng.randint(0, (2 ** (8 * rem)) - 1), width=2 * rem)))
        return words
    words = []
    for i in range(0, len(value), 32):
        chunk = value[i:i + 32]
        if len(chunk) & 0x1:
            chunk = chunk + '0'                                  # tb/gcm_gctr.py:383-384
        words.append(bytes.fromhex(chunk))
    return words


# --------------------------------------------------------------------------- key pins
def key_mode_val(mode):
    """4-bit key_word_val for a raw key load: 0b0100 / 0b0110 / 0b0111 (tb/gcm_gctr.py:152-157;
    bits 2/1/0 gate words 7-4 / 3-2 / 1-0, config/config_aes_ecb.py:174-203)."""
    return {'128': 0b0100, '192': 0b0110, '256': 0b0111}[str(mode)]


def pack_key_word(key):
    """Raw key -> the 256-bit key word: LEFT-aligned, trailing zeros (tb/gcm_gctr.py:159-165).
    key: {'data': hex, 'n_bytes': n} -> int (256 bits)."""
    key_ext = key['data'] + (AES_KEY_256_WIDTH // 4 - int(key['n_bytes']) * 2) * '0'
    return int(key_ext, 16)


def unpack_key_word(word, mode):
    n = {'128': 16, '192': 24, '256': 32}[str(mode)]
    return (word >> (8 * (32 - n))).to_bytes(n, 'big')


def pack_pre_expanded_key(expanded):
    """Expanded key bytes (176/208/240) -> list of (val, 256-bit word): stage i in the TOP 128
    bits, val = i+1 (tb/gcm_gctr.py:199-207; config/config_aes_kprexp.py:78,91-93)."""
    expanded = bytes(expanded)
    if len(expanded) not in (176, 208, 240):
        raise ValueError("expanded key must be 176, 208 or 240 bytes")
    out = []
    for i in range(len(expanded) // 16):
        stage = expanded[16 * i:16 * i + 16]
        out.append((i + 1, int.from_bytes(stage, 'big') << 128))
    return out


def unpack_pre_expanded_key(writes):
    """Inverse of pack_pre_expanded_key: slot val-1 <- top 128 bits (config_aes_kprexp.py:85-95)."""
    stages = {}
    for val, word in writes:
        stages[val - 1] = (word >> 128).to_bytes(16, 'big')
    return b"".join(stages[i] for i in range(len(stages)))


# --------------------------------------------------------------------------- data pins
def pack_data_word(block):
    """<=16 bytes -> (128-bit word, 16-bit byte-valid), both left-aligned
    (src/gcm_ghash.vhd:228-244: 0x8000 = 1 byte ... 0xFFFF = 16 bytes)."""
    n = len(block)
    if not 1 <= n <= 16:
        raise ValueError("a bus word carries 1..16 bytes")
    word = int.from_bytes(block, 'big') << (8 * (16 - n))
    bval = (0xFFFF << (16 - n)) & 0xFFFF
    return word, bval


def unpack_data_word(word, bval):
    """tb/gcm_driver.py:108-129: count leading ones of bval, take that many leftmost bytes."""
    n = 0
    v = bval
    for _ in range(16):
        if v & 0x8000:
            n += 1
            v <<= 1
        else:
            break
    return (word >> ((16 - n) * 8)).to_bytes(n, 'big') if n else b""


# --------------------------------------------------------------------------- replay
def replay(config, model_factory, rng=None, pre_expanded=False, expand_key=None):
    """Run one reference test configuration against a `gcm`-compatible model with no DUT
    (tb/gcm_test.py:36-140 with the DUT's outputs replaced by the model's own).

    model_factory(key_dict, iv_dict, ed) -> object with load_aad / load_plain_text /
    load_cipher_text / get_tag / data_out / tag (tb/gcm_model.py:5-51).
    For 'dec' the ciphertext and tag fed to the model are produced by a first 'enc' pass of the
    same model class, as the DUT would have produced them.
    pre_expanded: hand the model the Nr+1 stages instead of the raw key (the -x / rmexp flow,
    tb/gcm_test.py:103-106); expand_key(key_hex, size_str) -> list[int] is then required.
    Returns a dict with the resolved data, the word lists and the model outputs."""
    # one RNG stream, consumed in the reference's order: config_data, then the AAD words, then the
    # data words (tb/gcm_gctr.py:229-332, 337-437)
    if rng is None:
        rng = random.Random(config.get('seed', 0))
    cfg, data = resolve_config(config, rng)
    key = dict(data['key'])
    if pre_expanded:
        exp = bytes(expand_key(key['data'], str(cfg.get('aes_mode', '128'))))
        # pin encoding round trip, then the model takes the stages as one long key
        exp = unpack_pre_expanded_key(pack_pre_expanded_key(exp))
        key = {'data': exp.hex().upper(), 'n_bytes': len(exp)}
    else:
        assert unpack_key_word(pack_key_word(key), cfg.get('aes_mode', '128')) == bytes.fromhex(key['data'])
    aad_words = split_words(cfg.get('aad', RANDOM_PARAM), data['aad_n_bytes'], rng)
    txt_words = split_words(cfg.get('data', RANDOM_PARAM), data['pt_n_bytes'], rng)
    # the words cross the bus as (word, bval) pairs and come back through the monitors
    aad_words = [unpack_data_word(*pack_data_word(w)) for w in aad_words]
    txt_words = [unpack_data_word(*pack_data_word(w)) for w in txt_words]
    ed = cfg.get('enc_dec', 'enc')

    enc = model_factory(key, data['iv'], 'enc')
    for w in aad_words:
        enc.load_aad(w)
    for w in txt_words:
        enc.load_plain_text(w)
    enc.get_tag(None)
    result = {'config': cfg, 'data': data, 'aad_words': aad_words, 'pt_words': txt_words,
              'ct_words': list(enc.data_out), 'tag': enc.tag[-1]}
    if ed == 'dec':
        dec = model_factory(key, data['iv'], 'dec')
        for w in aad_words:
            dec.load_aad(w)
        for w in enc.data_out:
            dec.load_cipher_text(w)
        dec.get_tag(enc.tag[-1])
        result['dec_words'] = list(dec.data_out)
        result['dec_tag'] = dec.tag[-1]
    return result
