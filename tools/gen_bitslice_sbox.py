#!/usr/bin/env python
"""Generates a bitsliced AES S-box circuit (tower field GF(((2^2)^2)^2), Canright-style) as
straight-line CUDA code, verifies it for all 256 inputs, and writes
tools/bitslice_sbox_generated.cuh for tools/microbench_bitslice.cu.

This is MEASUREMENT SCAFFOLDING for DESIGN.md 4.1 (T-table vs bitslice): it lets the B200 tell us
how many S-box evaluations per second the integer pipe can deliver when the S-box is computed
with logic instead of looked up.  It is not part of the product.

The circuit is derived, not copied: the field isomorphism is found by brute force, the
inversion uses the usual norm trick at each tower level, and gates are emitted through
operator overloading with common-subexpression elimination.
"""
import itertools
import os

HERE = os.path.dirname(os.path.abspath(__file__))

# ---------------------------------------------------------------- reference S-box (by definition)
def gf256_mul(a, b):
    r = 0
    for _ in range(8):
        if b & 1:
            r ^= a
        a = ((a << 1) ^ (0x11B if a & 0x80 else 0)) & 0x1FF
        a &= 0xFF if not (a & 0x100) else 0x1FF
        b >>= 1
    return r & 0xFF


def xtime(a):
    a <<= 1
    return (a ^ 0x11B) & 0xFF if a & 0x100 else a


def mul(a, b):
    r = 0
    while b:
        if b & 1:
            r ^= a
        a = xtime(a)
        b >>= 1
    return r


def sbox_ref(x):
    inv = 0
    if x:
        for y in range(1, 256):
            if mul(x, y) == 1:
                inv = y
                break
    s = inv
    for i in range(1, 5):
        s ^= ((inv << i) | (inv >> (8 - i))) & 0xFF
    return s ^ 0x63


SBOX = [sbox_ref(x) for x in range(256)]

# ---------------------------------------------------------------- tower field on plain integers
# GF(2^2) = {0,1,w,w+1}, w^2 = w + 1; element = 2 bits (b1 w + b0)
def m2(a, b):
    a1, a0, b1, b0 = a >> 1, a & 1, b >> 1, b & 1
    hi = (a1 & b1) ^ (a1 & b0) ^ (a0 & b1)
    lo = (a1 & b1) ^ (a0 & b0)
    return (hi << 1) | lo


N2 = 2  # w: z^2 + z + w is irreducible over GF(2^2)


def m4(a, b):  # GF(2^4) = GF(2^2)[z]/(z^2 + z + N2); element = (a1 z + a0), 2 bits each
    a1, a0, b1, b0 = a >> 2, a & 3, b >> 2, b & 3
    p = m2(a1, b1)
    hi = m2(a1, b0) ^ m2(a0, b1) ^ p
    lo = m2(a0, b0) ^ m2(p, N2)
    return (hi << 2) | lo


def find_nu():
    # y^2 + y + nu irreducible over GF(2^4): no root
    for nu in range(1, 16):
        if all((m4(y, y) ^ y ^ nu) != 0 for y in range(16)):
            return nu
    raise RuntimeError


NU = find_nu()


def m8(a, b):  # GF(2^8) = GF(2^4)[y]/(y^2 + y + NU)
    a1, a0, b1, b0 = a >> 4, a & 15, b >> 4, b & 15
    p = m4(a1, b1)
    hi = m4(a1, b0) ^ m4(a0, b1) ^ p
    lo = m4(a0, b0) ^ m4(p, NU)
    return (hi << 4) | lo


def find_iso():
    """8x8 bit matrix (list of 8 tower values: image of x^k) mapping AES GF(2^8) -> tower."""
    for beta in range(2, 256):
        # beta must be a root of x^8 + x^4 + x^3 + x + 1 in the tower field
        pw = [1]
        for _ in range(8):
            pw.append(m8(pw[-1], beta))
        if pw[8] ^ pw[4] ^ pw[3] ^ pw[1] ^ pw[0] == 0:
            cols = pw[:8]
            # check it is a ring isomorphism on a sample
            def to_t(x):
                r = 0
                for k in range(8):
                    if (x >> k) & 1:
                        r ^= cols[k]
                return r
            if len({to_t(x) for x in range(256)}) == 256 and all(
                    to_t(mul(a, b)) == m8(to_t(a), to_t(b)) for a, b in ((3, 7), (0x53, 0xCA), (200, 31))):
                return cols
    raise RuntimeError


ISO = find_iso()


def invert_matrix(cols):
    # cols[k] = image of basis vector k; build inverse by brute force table
    fwd = {}
    for x in range(256):
        r = 0
        for k in range(8):
            if (x >> k) & 1:
                r ^= cols[k]
        fwd[r] = x
    return [fwd[1 << k] for k in range(8)]


ISO_INV = invert_matrix(ISO)

# ---------------------------------------------------------------- circuit builder
class Circuit:
    def __init__(self):
        self.gates = []      # (op, a, b) with wire ids; inputs are ids 0..7
        self.cse = {}
        self.n = 8

    def gate(self, op, a, b):
        if a > b:
            a, b = b, a
        key = (op, a, b)
        if key in self.cse:
            return self.cse[key]
        self.gates.append((op, a, b))
        self.cse[key] = self.n
        self.n += 1
        return self.n - 1


C = Circuit()
ZERO, ONE = -1, -2


class W:
    """A wire: id >= 0, or constant ZERO / ONE; `inv` marks a pending NOT (folded into XORs)."""
    __slots__ = ("i", "inv")

    def __init__(self, i, inv=False):
        self.i, self.inv = i, inv

    def __xor__(self, o):
        if self.i == ZERO:
            return W(o.i, o.inv ^ self.inv)
        if o.i == ZERO:
            return W(self.i, self.inv ^ o.inv)
        if self.i == o.i:
            return W(ZERO, self.inv ^ o.inv)
        return W(C.gate("^", self.i, o.i), self.inv ^ o.inv)

    def __and__(self, o):
        assert not self.inv and not o.inv   # ANDs only see plain wires in this construction
        if self.i == ZERO or o.i == ZERO:
            return W(ZERO)
        if self.i == o.i:
            return W(self.i)
        return W(C.gate("&", self.i, o.i))


Z = W(ZERO)


def lin(bits, cols, nout=8):
    """Apply a GF(2) matrix given by the images `cols` of each input bit."""
    out = []
    for j in range(nout):
        acc = Z
        for k, w in enumerate(bits):
            if (cols[k] >> j) & 1:
                acc = acc ^ w
        out.append(acc)
    return out


# GF(2^2) on wires: element = [b0, b1]
def g2_mul(a, b):
    t = (a[0] ^ a[1]) & (b[0] ^ b[1])
    p0 = a[0] & b[0]
    p1 = a[1] & b[1]
    return [p0 ^ p1, t ^ p0]          # lo = a0b0 ^ a1b1 ; hi = a1b1 ^ a1b0 ^ a0b1 = t ^ a0b0


def g2_sq(a):      # (b1 w + b0)^2 = b1 (w+1) + b0 = b1 w + (b0 ^ b1)
    return [a[0] ^ a[1], a[1]]


def g2_scl_N(a):   # multiply by w:  (b1 w + b0) w = b1 (w+1) + b0 w = (b0^b1) w + b1
    return [a[1], a[0] ^ a[1]]


def g2_add(a, b):
    return [a[0] ^ b[0], a[1] ^ b[1]]


# GF(2^4): element = [lo(2 wires), hi(2 wires)]
def g4_mul(a, b):
    p = g2_mul(a[1], b[1])
    t = g2_mul(g2_add(a[0], a[1]), g2_add(b[0], b[1]))
    q = g2_mul(a[0], b[0])
    hi = g2_add(t, q)                                  # a1b0 + a0b1 + a1b1
    lo = g2_add(q, g2_scl_N(p))
    return [lo, hi]


def g4_add(a, b):
    return [g2_add(a[0], b[0]), g2_add(a[1], b[1])]


def g4_inv(a):
    # norm trick: d = a1^2 N + a1 a0 + a0^2 in GF(2^2); d^-1 = d^2; a^-1 = (a1 d^-1) z + (a1 + a0) d^-1
    d = g2_add(g2_add(g2_scl_N(g2_sq(a[1])), g2_mul(a[1], a[0])), g2_sq(a[0]))
    di = g2_sq(d)
    return [g2_mul(g2_add(a[1], a[0]), di), g2_mul(a[1], di)]


def const4(v):
    return [[W(ONE if (v >> 0) & 1 else ZERO), W(ONE if (v >> 1) & 1 else ZERO)],
            [W(ONE if (v >> 2) & 1 else ZERO), W(ONE if (v >> 3) & 1 else ZERO)]]


def g4_mul_const(a, cst):
    """a * cst for a fixed cst in GF(2^4): linear, expand as a matrix on the 4 wires."""
    bits = [a[0][0], a[0][1], a[1][0], a[1][1]]
    cols = [m4(1 << k, cst) for k in range(4)]
    o = lin(bits, cols, 4)
    return [[o[0], o[1]], [o[2], o[3]]]


def g4_sq(a):
    bits = [a[0][0], a[0][1], a[1][0], a[1][1]]
    cols = [m4(1 << k, 1 << k) for k in range(4)]   # squaring is linear
    o = lin(bits, cols, 4)
    return [[o[0], o[1]], [o[2], o[3]]]


def g8_inv(a):   # a = [lo(GF16), hi(GF16)]
    d = g4_add(g4_add(g4_mul_const(g4_sq(a[1]), NU), g4_mul(a[1], a[0])), g4_sq(a[0]))
    di = g4_inv(d)
    return [g4_mul(g4_add(a[1], a[0]), di), g4_mul(a[1], di)]


def build():
    x = [W(k) for k in range(8)]
    t = lin(x, ISO)                                            # AES basis -> tower basis
    a = [[[t[0], t[1]], [t[2], t[3]]], [[t[4], t[5]], [t[6], t[7]]]]
    inv = g8_inv(a)
    flat = [inv[0][0][0], inv[0][0][1], inv[0][1][0], inv[0][1][1], inv[1][0][0], inv[1][0][1], inv[1][1][0], inv[1][1][1]]
    # tower -> AES basis fused with the affine map: s_j = sum_i A[j][i] inv_i, A = circulant 0x1F
    aff_cols = []
    for k in range(8):
        v = ISO_INV[k]
        s = v
        for i in range(1, 5):
            s ^= ((v << i) | (v >> (8 - i))) & 0xFF
        aff_cols.append(s)
    out = lin(flat, aff_cols)
    return [W(o.i, o.inv ^ bool((0x63 >> j) & 1)) for j, o in enumerate(out)]


OUT = build()


def simulate():
    """Evaluate the netlist on all 256 inputs at once (bit x of every wire value = input x)."""
    vals = {}
    full = (1 << 256) - 1
    for k in range(8):
        v = 0
        for x in range(256):
            if (x >> k) & 1:
                v |= 1 << x
        vals[k] = v
    for idx, (op, a, b) in enumerate(C.gates):
        vals[8 + idx] = (vals[a] ^ vals[b]) if op == "^" else (vals[a] & vals[b])
    res = []
    for o in OUT:
        v = vals[o.i] if o.i >= 0 else (full if o.i == ONE else 0)
        res.append(v ^ (full if o.inv else 0))
    for x in range(256):
        got = sum(((res[j] >> x) & 1) << j for j in range(8))
        assert got == SBOX[x], (x, got, SBOX[x])


def emit(path):
    n_and = sum(1 for g in C.gates if g[0] == "&")
    with open(path, "w") as f:
        f.write("// GENERATED by tools/gen_bitslice_sbox.py -- do not edit.  Bitsliced AES S-box: %d gates (%d AND, %d XOR)\n"
                "// + output inversions; x[k] / s[k] hold bit k of 32 bytes.  Verified for all 256 inputs by the generator.\n"
                % (len(C.gates), n_and, len(C.gates) - n_and))
        f.write("#define BITSLICE_SBOX_GATES %d\n" % len(C.gates))
        f.write("__device__ __forceinline__ void bitslice_sbox(const uint32_t x[8], uint32_t s[8])\n{\n")
        for idx, (op, a, b) in enumerate(C.gates):
            na = "x[%d]" % a if a < 8 else "t%d" % a
            nb = "x[%d]" % b if b < 8 else "t%d" % b
            f.write("    const uint32_t t%d = %s %s %s;\n" % (8 + idx, na, op, nb))
        for j, o in enumerate(OUT):
            name = "x[%d]" % o.i if 0 <= o.i < 8 else "t%d" % o.i
            f.write("    s[%d] = %s%s;\n" % (j, "~" if o.inv else "", name))
        f.write("}\n")


if __name__ == "__main__":
    simulate()
    out = os.path.join(HERE, "bitslice_sbox_generated.cuh")
    emit(out)
    n_and = sum(1 for g in C.gates if g[0] == "&")
    print("S-box circuit verified for all 256 inputs: %d gates (%d AND, %d XOR); nu=%d; wrote %s"
          % (len(C.gates), n_and, len(C.gates) - n_and, NU, out))
