"""The integer arithmetic of rescore_int8cos_dp2a_kernel (csrc/rescore.cu, VRQ_RESCORE_DP2A=1) restated in NumPy:
the float32 query as 64-bit fixed point against its own largest exponent, four 16-bit limbs (three unsigned, the top one
signed), limb x int8 products summed in int32 per lane, recombined in float64.  Must agree with the float64 dot product
to ~1e-16 of sum|q_i d_i| and never overflow an int32 lane sum."""
import numpy as np
import pytest


def fixed_point_dot(q: np.ndarray, d: np.ndarray) -> float:
    amax = np.float32(np.max(np.abs(q)))
    bexp = (int(amax.view(np.uint32)) >> 23) & 0xFF
    if bexp == 0:
        return 0.0
    e = bexp - 127
    big = np.rint(q.astype(np.float64) * 2.0 ** (61 - e)).astype(np.int64)   # |Q| < 2^62
    limbs = [(big >> (16 * l)) & 0xFFFF for l in range(3)] + [big >> 48]     # arithmetic shift: the top limb is signed
    total = 0.0
    lane_sums = []
    for lane in range(32):  # the kernel's lane owns elements 16 lane .. +15 and 512 + 16 lane .. +15
        idx = np.r_[16 * lane:16 * lane + 16, 512 + 16 * lane:512 + 16 * lane + 16]
        a = [int(np.sum(limb[idx] * d[idx].astype(np.int64))) for limb in limbs]
        assert all(abs(v) < 2 ** 31 for v in a)
        s = float(a[3])
        for l in (2, 1, 0):
            s = s * 65536.0 + float(a[l])
        lane_sums.append(s)
    total = float(np.sum(np.array(lane_sums)))
    return total * 2.0 ** (e - 61)


@pytest.mark.parametrize("scale", [1.0, 1e20, 1e-25, 3e38 / 8, 1e-37])
def test_fixed_point_matches_float64(scale):
    rng = np.random.default_rng(5)
    for trial in range(6):
        q = (rng.normal(0, 0.05, 1024) * scale).astype(np.float32)
        if trial == 1:
            q[::2] *= np.float32(1e-12)   # wide exponent range inside one query
        if trial == 2:
            q[:] = 0
        d = rng.integers(-128, 128, 1024).astype(np.int8)
        if trial == 3:
            d[:] = -128                   # the largest products
            q = np.abs(q)
        ref = float(np.dot(q.astype(np.float64), d.astype(np.float64)))
        mag = float(np.dot(np.abs(q.astype(np.float64)), np.abs(d.astype(np.float64))))
        got = fixed_point_dot(q, d)
        assert abs(got - ref) <= 4e-16 * mag, (scale, trial, got, ref)
