"""GPU parity tests, kernel level: every libvrq entry point against the CPU oracle on the same inputs, through the
C ABI (host-buffer path unless stated).  Bit-exact for codes / bits / distances / ids; floats to 1e-5 relative
(+ the float32-accumulation floor of the reference's own sdot)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle_c as oc  # noqa: E402
from oracle import vrq_oracle as o  # noqa: E402


@pytest.fixture(scope="module")
def K():
    from vectorragquantization_b200 import kernels
    return kernels


def adversarial_rows(d=1024, limit=0.3):
    rng = np.random.default_rng(1234)
    rows = []
    rows.append(np.zeros(d, np.float32))
    rows.append(np.full(d, 0.25, np.float32))
    rows.append(np.where(np.arange(d) % 2 == 0, limit, -limit).astype(np.float32))
    for qmax in (127.0, 32767.0, 7.0):
        rows.append(((np.arange(d) - d // 2 + 0.5) * np.float32(limit / qmax)).astype(np.float32))  # exact .5 steps
    r = np.full(d, np.nextafter(np.float32(limit), np.float32(9)), np.float32)
    r[::3] = -r[::3]
    rows.append(r)
    rows.append(np.where(np.arange(d) < d // 2, 1.0, -1.0).astype(np.float32))
    r = rng.normal(0, 0.05, d).astype(np.float32)
    r[7] = r.mean()  # an element (nearly) equal to the mean
    rows.append(r)
    r = np.zeros(d, np.float32)
    r[17] = 1e-30
    rows.append(r)
    r = np.zeros(d, np.float32)
    r[3] = -0.0
    r[5] = 1e-42  # denormal
    rows.append(r)
    rows.append((rng.normal(0, 1, d) * 1e-4).astype(np.float32))
    rows.append((rng.normal(0, 5, d)).astype(np.float32))
    return np.stack(rows)


def make_x(n, d=1024, seed=0):
    rng = np.random.default_rng(seed)
    sig = np.array([1e-4, 0.036, 0.2, 1.0, 5.0])[rng.integers(0, 5, n)]
    x = (rng.normal(0, 1, (n, d)) * sig[:, None]).astype(np.float32)
    adv = adversarial_rows(d)
    x[: len(adv)] = adv
    return x


@pytest.mark.parametrize("d", [1024, 256, 384, 1000, 4096])
def test_encoders_bit_exact(K, d):
    n = 3000 if d == 1024 else 300
    x = make_x(n, d, seed=d)
    q, lo, hi, ub = K.quantize_int8_perdoc(x, want_binary=True)
    rq, rlo, rhi = o.quantize_int8_perdoc(x)
    rub = o.to_binary_f32(x)
    assert np.array_equal(lo, rlo) and np.array_equal(hi, rhi)
    bad = np.nonzero((q != rq).any(axis=1))[0]
    assert bad.size == 0, (bad[:10], [(int(c), int(q[r, c]), int(rq[r, c])) for r in bad[:3] for c in np.nonzero(q[r] != rq[r])[0][:4]])
    bad = np.nonzero((ub != rub).any(axis=1))[0]
    assert bad.size == 0, bad[:10]
    for lim in (0.18, 0.3, 1.0):
        q8, ub8 = K.quantize_int8_global(x, lim, want_binary=True)
        assert np.array_equal(q8, o.quantize_int8_global(x, lim)) and np.array_equal(ub8, rub)
        q16 = K.quantize_int16_global(x, lim)
        assert np.array_equal(q16, o.quantize_int16_global(x, lim))
    p4, l4, h4, ub4 = K.quantize_int4(x, want_binary=True)
    r4, rl4, rh4 = o.quantize_int4(x)
    bad = np.nonzero((p4 != r4).any(axis=1))[0]
    assert bad.size == 0, (bad[:10], [(int(c), int(p4[r, c]), int(r4[r, c])) for r in bad[:3] for c in np.nonzero(p4[r] != r4[r])[0][:4]])
    assert np.array_equal(l4, rl4) and np.array_equal(h4, rh4) and np.array_equal(ub4, rub)
    assert np.array_equal(K.to_binary(x), rub)
    assert np.array_equal(K.to_binary(x, ge=True), o.to_binary_f32(x, ge=True))
    # single-vector call shape, like the reference's static methods
    q1, a1, b1 = K.quantize_int8_perdoc(x[20])
    assert q1.shape == (d,) and np.array_equal(q1, rq[20]) and a1 == rlo[20] and b1 == rhi[20]


@pytest.mark.parametrize("shape", ["register", "ring_8x2", "ring_8x4", "ring_4x6", "ring_3x3", "ring_1x2"])
@pytest.mark.parametrize("n", [1, 7, 1201])
def test_encoder_launch_shapes_bit_exact(K, monkeypatch, shape, n):
    """d = 1024 has two kernels (encode1024_kernel: registers; encode1024_ring_kernel: cp.async ring, magic-number rounding)
    and the ring kernel several launch shapes; every one must give the oracle's bytes, ragged row counts included."""
    if shape == "register":
        monkeypatch.setenv("VRQ_ENCODE_RING", "0")
    else:
        w, s = shape.split("_")[1].split("x")
        monkeypatch.setenv("VRQ_ENCODE_RING", "1")
        monkeypatch.setenv("VRQ_ENCODE_WARPS", w)
        monkeypatch.setenv("VRQ_ENCODE_STAGES", s)
    x = np.ascontiguousarray(make_x(max(n, 16), 1024, seed=n)[:n])  # the adversarial rows come first
    rub = o.to_binary_f32(x)
    q, lo, hi, ub = K.quantize_int8_perdoc(x, want_binary=True)
    rq, rlo, rhi = o.quantize_int8_perdoc(x)
    assert np.array_equal(q, rq) and np.array_equal(lo, rlo) and np.array_equal(hi, rhi) and np.array_equal(ub, rub)
    for lim in (0.18, 0.3, 1.0):
        q8, ub8 = K.quantize_int8_global(x, lim, want_binary=True)
        assert np.array_equal(q8, o.quantize_int8_global(x, lim)) and np.array_equal(ub8, rub)
        q16, ub16 = K.quantize_int16_global(x, lim, want_binary=True)
        assert np.array_equal(q16, o.quantize_int16_global(x, lim)) and np.array_equal(ub16, rub)
        assert np.array_equal(K.quantize_int8_global(x, lim), q8)  # without the fused code
    p4, l4, h4, ub4 = K.quantize_int4(x, want_binary=True)
    r4, rl4, rh4 = o.quantize_int4(x)
    assert np.array_equal(p4, r4) and np.array_equal(l4, rl4) and np.array_equal(h4, rh4) and np.array_equal(ub4, rub)
    assert np.array_equal(K.quantize_int4(x)[0], r4) and np.array_equal(K.quantize_int8_perdoc(x)[0], rq)
    assert np.array_equal(K.to_binary(x), rub)
    assert np.array_equal(K.to_binary(x, ge=True), o.to_binary_f32(x, ge=True))


def test_encoders_golden(K, golden_static):
    g = golden_static
    x = g["x"]
    q, lo, hi = K.quantize_int8_perdoc(x)
    assert np.array_equal(q, g["int8_perdoc.q"]) and np.array_equal(np.stack([lo, hi], 1), g["int8_perdoc.min_max"])
    for lim in (0.18, 0.3, 1.0):
        assert np.array_equal(K.quantize_int8_global(x, lim), g[f"int8_global.q.{lim}"])
        assert np.array_equal(K.quantize_int16_global(x, lim), g[f"int16_global.q.{lim}"])
    p, l4, h4 = K.quantize_int4(x)
    assert np.array_equal(p, g["int4.q"]) and np.array_equal(np.stack([l4, h4], 1), g["int4.min_max"])
    assert np.array_equal(K.to_binary(x), g["ubinary_f32"])
    assert np.array_equal(K.to_binary(x, ge=True), g["ubinary_f32_ge"])
    assert np.array_equal(K.to_binary(g["i8"]), g["ubinary_i8"])
    assert np.array_equal(K.to_binary(g["i16"]), g["ubinary_i16"])
    assert np.array_equal(K.dequantize_int8_perdoc(g["int8_perdoc.q"], lo, hi), g["int8_perdoc.deq"])
    assert np.array_equal(K.dequantize_int8_global(g["int8_global.q.0.3"], 0.3), g["int8_global.deq.0.3"])
    assert np.array_equal(K.dequantize_int16_global(g["int16_global.q.1.0"], 1.0), g["int16_global.deq.1.0"])


def test_kat1_on_gpu(K, golden_dbs):
    pay = golden_dbs["db_cohere_int8.payload"]
    assert np.array_equal(K.to_binary(pay), golden_dbs["db_cohere_int8.codes"][: pay.shape[0]])


@pytest.mark.parametrize("variant", ["dp2a", "register_ring", "batchred", "cuda_core", "imma", "imma62", "imma143", "imma43", "imma123", "imma121", "imma81", "imma34"])
def test_rescore_int8cos_kernel_variants(K, monkeypatch, variant):
    """Phase III has several d = 1024 kernels: the cp.async ring with float64 FMAs, the same ring with integer dot
    products on 16-bit limbs of the fixed-point query (VRQ_RESCORE_DP2A=1), with batched reductions
    (VRQ_RESCORE_BATCHRED=1), the register ring (VRQ_RESCORE_ASYNC=0) and the tensor-core kernel (mma.sync s8 over eight
    base-256 digits of the fixed-point query, VRQ_RESCORE_IMMA=1, several launch shapes).  All must agree with the
    float64 evaluation far inside the parity tolerance."""
    rng = np.random.default_rng(11)
    n, nq, m = 4000, 9, 257
    x = o.synth_f32(31, 0, n)
    i8 = o.synth_int8_from_f32(x)
    i8[17] = 0
    qf = o.synth_f32(32, 0, nq)
    qf[1] = 0                    # all-zero query
    qf[2, ::2] *= 1e-12          # wide exponent range inside one query
    qf[3] *= 1e20                # large magnitudes
    qf[4] *= 1e-25               # small magnitudes
    pos = rng.integers(0, n, (nq, m))
    pos[0, 0] = 17
    pos[5, 3] = -1
    pos[6, -40:] = -1            # a whole 16-row group without a valid row
    base = K.rescore_int8cos(i8, pos, qf)
    monkeypatch.setenv("VRQ_RESCORE_IMMA", "1" if variant.startswith("imma") else "0")
    if variant == "dp2a":
        monkeypatch.setenv("VRQ_RESCORE_DP2A", "1")
    elif variant == "batchred":
        monkeypatch.setenv("VRQ_RESCORE_BATCHRED", "1")
    elif variant == "register_ring":
        monkeypatch.setenv("VRQ_RESCORE_ASYNC", "0")
    elif variant.startswith("imma") and len(variant) > 4:
        monkeypatch.setenv("VRQ_RESCORE_IMMA_SHAPE", variant[4:])
    sc = K.rescore_int8cos(i8, pos, qf)
    for i in range(nq):
        ok = pos[i] >= 0
        rc = np.full(m, -np.inf)
        rc[ok] = o.rescore_int8cos(qf[i], i8[pos[i][ok]], literal=False)
        fin = np.isfinite(rc)
        assert np.array_equal(np.isfinite(sc[i]), fin) and np.array_equal(np.isfinite(base[i]), fin)
        mag = np.zeros(m)
        mag[ok] = o.rescore_int8cos_absfloor(qf[i], i8[pos[i][ok]]) / (4.0 * 2.0 ** -24)  # sum|q x| / |x|
        assert np.all(np.abs(sc[i][fin] - rc[fin]) <= 1e-13 * mag[fin]), i
        assert np.all(np.abs(base[i][fin] - rc[fin]) <= 1e-13 * mag[fin]), i
    assert sc[0, 0] == -np.inf and sc[5, 3] == -np.inf


def test_rescore_binary_kernel_variants(K, monkeypatch):
    """Phase II for d = 1024: the tensor-core kernel (default: mma.sync s8 over the code bits x digit planes of the fixed-point
    query), the nibble-table kernel and the register kernel agree with the float64 evaluation to rounding, incl. invalid
    positions, all-zero and wide-range queries, m not a multiple of the group / block size."""
    rng = np.random.default_rng(12)
    n, nq, m = 3000, 6, 1000
    codes = rng.integers(0, 256, (n, 128)).astype(np.uint8)
    codes[5] = 0
    codes[6] = 255
    qf = o.synth_f32(33, 0, nq)
    qf[1] = 0
    qf[2, ::2] *= 1e-12
    qf[3] *= 1e20
    pos = rng.integers(0, n, (nq, m))
    pos[0, :2] = (5, 6)
    pos[4, 7] = -1
    pos[5, -20:] = -1
    got = {}
    for lut in ("2", "1", "0"):  # tensor cores (default) / nibble table / register kernel
        monkeypatch.setenv("VRQ_RESCORE_BIN", lut)
        got[lut] = K.rescore_binary(codes, pos, qf)
    for i in range(nq):
        ok = pos[i] >= 0
        ref = o.rescore_binary(qf[i], codes[pos[i][ok]], literal=False)
        mag = np.abs(qf[i].astype(np.float64)).sum()
        for lut in ("2", "1", "0"):
            assert np.all(np.abs(got[lut][i][ok] - ref) <= 1e-13 * mag + 1e-300), (lut, i)
            assert np.all(got[lut][i][~ok] == -np.inf)


def test_dequant_int4(K):
    x = make_x(500, 1024, seed=5)
    p, lo, hi = o.quantize_int4(x)
    assert np.array_equal(K.dequantize_int4_perdoc(p, 1024, lo, hi), o.dequantize_int4_perdoc(p, 1024, lo, hi))
    assert np.array_equal(K.dequantize_int4_global(p, 1024, 0.18), o.dequantize_int4_global(p, 1024, 0.18))


def test_to_binary_int(K):
    rng = np.random.default_rng(8)
    for dt, lo, hi in ((np.int8, -128, 128), (np.int16, -32768, 32768)):
        for d in (1024, 384):
            x = rng.integers(lo, hi, (700, d)).astype(dt)
            x[0] = 7
            x[1] = np.where(np.arange(d) % 2 == 0, 3, 4)
            assert np.array_equal(K.to_binary(x), o.to_binary_int(x))
            assert np.array_equal(K.to_binary(x, ge=True), o.to_binary_int(x, ge=True))


def test_large_host_pipeline_chunks(K):
    """> 1 staging chunk (96 MiB) so the double-buffered H2D/kernel/D2H pipeline is exercised."""
    n = 40000
    x = oc.synth_f32(3, 0, n, 1024, row_scale=True)
    q, ub = K.quantize_int8_global(x, 0.3, want_binary=True)
    assert np.array_equal(q, oc.quantize_int8_global(x, 0.3))
    assert np.array_equal(ub, oc.to_binary_f32(x))


def test_synth_matches_oracle(K):
    a = K.synth_f32(5, 1000, 300, 1024, row_scale=True)
    assert np.array_equal(a, oc.synth_f32(5, 1000, 300, 1024, row_scale=True))
    c, i8 = K.synth_codes_int8(9, 123456789012, 300)
    rc, ri8 = oc.synth_codes_int8(9, 123456789012, 300)
    assert np.array_equal(c, rc) and np.array_equal(i8, ri8)


def test_rescore_kernels(K):
    rng = np.random.default_rng(11)
    n, nq, m = 5000, 7, 300
    x = o.synth_f32(21, 0, n)
    codes, i8 = o.synth_ubinary_from_f32(x), o.synth_int8_from_f32(x)
    i8[17] = 0
    qf = o.synth_f32(22, 0, nq)
    pos = rng.integers(0, n, (nq, m))
    pos[0, 0] = 17
    sb = K.rescore_binary(codes, pos, qf)
    sc = K.rescore_int8cos(i8, pos, qf)
    for i in range(nq):
        rb = o.rescore_binary(qf[i], codes[pos[i]])
        assert np.all(np.abs(sb[i] - rb) <= 1e-5 * np.abs(rb) + 1e-12)
        rc = o.rescore_int8cos(qf[i], i8[pos[i]])
        fl = o.rescore_int8cos_absfloor(qf[i], i8[pos[i]])
        fin = np.isfinite(rc)
        assert np.array_equal(np.isfinite(sc[i]), fin)
        assert np.all(np.abs(sc[i][fin] - rc[fin]) <= 1e-5 * np.abs(rc[fin]) + fl[fin])
    assert sc[0, 0] == -np.inf
