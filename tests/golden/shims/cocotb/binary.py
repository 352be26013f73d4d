class BinaryValue:
    """Just enough of cocotb.binary.BinaryValue for tb/gcm_gctr.py:163-165,201-202."""

    def __init__(self, n_bits=0):
        self.n_bits = n_bits
        self._v = 0

    def assign(self, s):
        self._v = int(s, 2)

    def get_value(self):
        return self._v
