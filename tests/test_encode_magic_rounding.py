"""The arithmetic shortcuts of encode1024_ring_kernel (csrc/encode.cu), restated in NumPy float32 and held to the oracle
bit for bit: symmetric clip as min(|x|, lim) with the sign of x, np.round + cast as one add of 1.5 * 2^23 whose low
mantissa bits are the integer, no clip at all for int4 on an ordinary row, the +8 nibble bias folded into the constant,
and mean = sum * 2^-10.  (The kernel itself is checked against the same oracle on the GPU in test_gpu_kernels.py.)"""
import numpy as np
import pytest

from oracle import vrq_oracle as o

MAGIC = np.float32(12582912.0)  # 1.5 * 2^23


def rows(n=600, d=1024, seed=7):
    rng = np.random.default_rng(seed)
    sig = np.array([1e-4, 0.036, 0.2, 1.0, 5.0])[rng.integers(0, 5, n)]
    x = (rng.normal(0, 1, (n, d)) * sig[:, None]).astype(np.float32)
    for i, (lim, qmax) in enumerate([(0.18, 127.0), (0.3, 127.0), (1.0, 32767.0), (0.3, 7.0), (1.0, 127.0)]):
        x[i] = ((np.arange(d) - d // 2 + 0.5) * np.float32(lim / qmax)).astype(np.float32)  # exact .5 steps: ties
    x[5] = np.where(np.arange(d) % 2 == 0, 0.3, -0.3)
    x[6] = np.nextafter(np.float32(0.3), np.float32(9))
    x[7, :] = 0
    x[7, 3] = -0.0
    return x


@pytest.mark.parametrize("lim", [0.18, 0.3, 1.0])
@pytest.mark.parametrize("qmax,fn,dtype", [(127.0, o.quantize_int8_global, np.int8), (32767.0, o.quantize_int16_global, np.int16)])
def test_global_codecs(lim, qmax, fn, dtype):
    x = rows()
    limf, scale = np.float32(lim), np.float32(qmax / lim)
    assert limf * scale < qmax + 0.49  # the host-side guard of the kernel (ring_params_ok)
    c = np.copysign(np.minimum(np.abs(x), limf), x).astype(np.float32)      # min.xorsign.abs.f32
    bits = ((c * scale).astype(np.float32) + MAGIC).astype(np.float32).view(np.int32)
    got = bits.astype(np.uint32).astype(np.uint8 if dtype is np.int8 else np.uint16).view(dtype)
    assert np.array_equal(got, fn(x, lim))


def test_int4_without_clip():
    x = rows()
    ref, lo, hi = o.quantize_int4(x)
    m = np.maximum(np.abs(lo), np.abs(hi)).astype(np.float32)
    ordinary = (m >= np.float32(1e-36)) & (m <= np.float32(1e36))
    assert ordinary.sum() >= len(x) - 2
    with np.errstate(all="ignore"):
        scale = (7.0 / m.astype(np.float64)).astype(np.float32)
        t = (x * scale[:, None]).astype(np.float32)
    assert np.all(np.abs(t[ordinary]) <= 7.000001)                           # why np.clip(., -8, 7) never acts
    bits = (t + np.float32(12582920.0)).astype(np.float32).view(np.int32)    # MAGIC + 8: the low 4 bits are the nibble
    by = ((bits[:, 0::2] * 16 + bits[:, 1::2]) & 0xFF).astype(np.uint8).view(np.int8)
    assert np.array_equal(by[ordinary & (lo != hi)], ref[ordinary & (lo != hi)])


def test_mean_scaling_is_a_division():
    rng = np.random.default_rng(3)
    t = np.concatenate([rng.normal(0, 50, 4000), rng.normal(0, 1e-38, 2000), [0.0, -0.0, 1e-45, -1e-45, 3e38]]).astype(np.float32)
    with np.errstate(all="ignore"):
        assert np.array_equal((t / np.float32(1024)).view(np.uint32), (t * np.float32(2.0 ** -10)).view(np.uint32))
