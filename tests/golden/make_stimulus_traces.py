#!/usr/bin/env python
"""Records what the REFERENCE stimulus code itself (tb/gcm_gctr.py, executed from
/root/reference with cocotb / progress shims and a fake DUT) produces:

* `config_data` (tb/gcm_gctr.py:229-332) for seeded configurations: normalised key / IV,
  byte counts, delay mask -- the RNG is Python's global `random`, seeded like cocotb does;
* `encrypt_data` (tb/gcm_gctr.py:337-437): the AAD / data word lists (driven alone, so the RNG
  stream is not interleaved with the sequencer's);
* `load_key` / `load_pre_exp_key` (tb/gcm_gctr.py:144-214): every (key_word_val, key_word)
  pair written to the DUT pins.

  python tests/golden/make_stimulus_traces.py  ->  tests/golden/stimulus_traces.json
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, "/root/reference/tb")

import gcm_gctr  # the reference file itself


class Sig:
    def __init__(self, log, name):
        object.__setattr__(self, "_log", log)
        object.__setattr__(self, "_name", name)
        object.__setattr__(self, "_v", 0)

    def __setattr__(self, k, v):
        if k == "value":
            object.__setattr__(self, "_v", v)
            self._log.append((self._name, int(v)))
        else:
            object.__setattr__(self, k, v)

    @property
    def value(self):
        return self._v


class Log:
    def info(self, *a, **k):
        pass
    debug = error = warning = info


class Dut:
    def __init__(self):
        self.writes = []
        self._log = Log()
        self.clk_i = object()
        self.aes_gcm_key_word_val_i = Sig(self.writes, "val")
        self.aes_gcm_key_word_i = Sig(self.writes, "word")


def run(gen, queues=()):
    for _ in gen:              # each yield is a clock edge: drain the bounded driver queues
        for q in queues:
            del q[:]


def key_writes(dut):
    """Pair up (val, word) as sampled at each clock edge, dropping the idle (0, x) states."""
    out, val, word = [], 0, 0
    for name, v in dut.writes:
        if name == "val":
            val = v
        else:
            word = v
            if val:
                out.append([val, "%064X" % word])
    return out


def main():
    cases = []
    cfgs = [
        {"seed": 11, "aes_mode": "128", "key": "RANDOM", "iv": "RANDOM", "aad": "RANDOM", "data": "RANDOM", "enc_dec": "enc", "max_n_byte": 4095},
        {"seed": 12, "aes_mode": "192", "key": "RANDOM", "iv": "RANDOM", "aad": "RANDOM", "data": "RANDOM", "enc_dec": "dec", "max_n_byte": 300},
        {"seed": 13, "aes_mode": "256", "key": "RANDOM", "iv": "ABCDEF", "aad": "EMPTY", "data": "RANDOM", "enc_dec": "enc", "max_n_byte": 100},
        {"seed": 14, "aes_mode": "ALL", "key": "RANDOM", "iv": "RANDOM", "aad": "RANDOM", "data": "EMPTY", "enc_dec": "dec", "max_n_byte": 2000},
        {"seed": 15, "aes_mode": "128", "key": "AD7A2BD03EAC835A6F620FDCB506B345", "iv": "12153524C0895E81B2C28465",
         "aad": "D609B1F056637A0D46DF998D88E52E00B2C2846512153524C0895E81",
         "data": "08000F101112131415161718191A1B1C1D1E1F202122232425262728292A2B2C2D2E2F303132333435363738393A0002",
         "enc_dec": "enc", "max_n_byte": 4095},
        {"seed": 16, "aes_mode": "256", "key": "1" * 70, "iv": "F" * 30, "aad": "ABCDE", "data": "A" * 33, "enc_dec": "enc", "max_n_byte": 4095},
    ]
    for cfg in cfgs:
        random.seed(cfg["seed"])
        dut = Dut()
        tb = gcm_gctr.gcm_gctr(dut)
        tb.config = dict(cfg)
        tb.config_data()
        aad_model, pt_model, aad_q, pt_q = [], [], [], []
        run(tb.encrypt_data(tb.data["aad_n_bytes"], tb.data["pt_n_bytes"], aad_q, pt_q, aad_model, pt_model), (aad_q, pt_q))
        del dut.writes[:]
        run(tb.load_key(tb.data["key"]))
        raw = key_writes(dut)
        del dut.writes[:]
        run(tb.load_pre_exp_key(tb.data["key"]))
        pre = key_writes(dut)
        cases.append({"config_in": cfg, "config_out": tb.config, "data": tb.data,
                      "aad_words": [w.hex() for w in aad_model], "pt_words": [w.hex() for w in pt_model],
                      "load_key": raw, "load_pre_exp_key": pre})
    with open(os.path.join(HERE, "stimulus_traces.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_stimulus_traces.py running /root/reference/tb/gcm_gctr.py",
                   "cases": cases}, f, indent=1)
    print("wrote", len(cases), "stimulus traces")


if __name__ == "__main__":
    main()
