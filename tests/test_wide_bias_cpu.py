"""The bias column of the swapped-operand scan kernel (scan_mma.cu, hamming_scan_mma_wide_kernel), restated in Python.

The kernel adds bias(q) = -(popc(q) - tau(q)) to every accumulator column with one extra MMA whose A operand is the constant
[4.0] * 62 + [1.0] * 2 (e2m1) and whose B operand encodes the bias as e2m1 codes.  This checks the encoding rule the CUDA
code uses: every integer in [-1440, 1440] is hit exactly, with codes that are valid e2m1 values at valid positions."""
import numpy as np

E2M1 = {0: 0.0, 1: 0.5, 2: 1.0, 3: 1.5, 4: 2.0, 5: 3.0, 6: 4.0, 7: 6.0}
A_CONST = np.array([4.0] * 62 + [1.0, 1.0])
BIAS_MAX = 1440  # WIDE_BIAS_MAX
REM_CODES = {0: (0, 0), 2: (1, 0), 4: (2, 0), 6: (3, 0), 8: (4, 0), 10: (4, 1), 12: (5, 0), 14: (5, 1), 16: (6, 0), 18: (6, 1),
             20: (6, 2), 22: (6, 3)}


def encode(bias):
    """Mirror of the per-query loop in the kernel's prologue: 64 nibbles (sign in bit 3)."""
    sgn = 0x8 if bias < 0 else 0x0
    mag = -bias if bias < 0 else bias
    odd = mag & 1
    n24, rem = divmod(mag - odd, 24)
    c1, c2 = REM_CODES[rem]
    nib = []
    for pos in range(64):
        code = 0
        if pos < n24:
            code = 7
        elif pos == n24:
            code = c1
        elif pos == n24 + 1:
            code = c2
        if pos == 62:
            code = 2 if odd else 0
        if pos == 63:
            code = 0
        if code:
            code |= sgn
        nib.append(code)
    return nib


def decode(nib):
    vals = np.array([(-1.0 if n & 8 else 1.0) * E2M1[n & 7] for n in nib])
    return float(np.dot(A_CONST, vals))


def test_every_bias_is_represented_exactly():
    for bias in range(-BIAS_MAX, BIAS_MAX + 1):
        nib = encode(bias)
        assert decode(nib) == bias
        # the products with the 4.0 elements occupy positions < 62 only; position 63 is never used
        mag = abs(bias)
        assert (mag - (mag & 1)) // 24 + 2 <= 62
        assert nib[63] == 0


def test_clamps_mean_always_and_never():
    # |dot| <= 1024: +BIAS_MAX makes every column positive ("always a candidate": thresholds at infinity, the fallback pass),
    # -BIAS_MAX makes every column negative ("never": padding columns)
    assert -1024 + BIAS_MAX > 0 and 1024 - BIAS_MAX < 0


def test_constant_operand_words():
    # the eight 32-bit words the kernel stores to tensor memory: nibble j of word w = K position 8 w + j
    words = [0x66666666] * 7 + [0x22666666]
    vals = []
    for w in words:
        for j in range(8):
            vals.append(E2M1[(w >> (4 * j)) & 7])
    assert np.array_equal(np.array(vals), A_CONST)
