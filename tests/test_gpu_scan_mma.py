"""GPU parity tests for the tensor-core Phase-I scan (csrc/scan_mma.cu) through the C ABI.

The contraction runs on tcgen05.mma.kind::i8; every test here checks it bit for bit against the oracle (or NumPy's
bitwise_count), and against the integer-pipe kernel of csrc/scan.cu, on the paths a query batch can take: thresholds from
a strided sample, the gated exact fallback, no sampling (small databases), ragged tails, ties."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle_c as oc  # noqa: E402
from oracle import vrq_oracle as o  # noqa: E402

ENVS = {
    "default": {},
    "int8_kind": {"VRQ_MMA_KIND": "8"},
    "single_ctas": {"VRQ_MMA_PAIR": "0"},
    "forced_fallback": {"VRQ_MMA_SAMPLE_K": "1", "VRQ_MMA_SAFETY": "1"},  # a threshold far too tight -> verification fails
    "no_sampling": {"VRQ_MMA_SAFETY": "0"},
    "force_mma_small_batch": {"VRQ_SCAN_MMA": "2"},
    "integer_pipes": {"VRQ_SCAN_MMA": "0"},
}


@pytest.fixture(scope="module")
def V():
    import vectorragquantization_b200 as v
    return v


def set_env(monkeypatch, name):
    for k in ("VRQ_MMA_SAMPLE_K", "VRQ_MMA_SAFETY", "VRQ_SCAN_MMA", "VRQ_MMA_RAW_STAGES", "VRQ_MMA_KIND", "VRQ_MMA_PAIR",
              "VRQ_MMA_GROUP_TILES", "VRQ_MMA_FEW", "VRQ_MMA_MID", "VRQ_MMA_TAIL", "VRQ_MMA_LOCKSTEP", "VRQ_MMA_W128", "VRQ_MMA_VAR"):
        monkeypatch.delenv(k, raising=False)
    for k, v in ENVS[name].items():
        monkeypatch.setenv(k, v)


def ref_distances(q, codes):
    return np.stack([np.bitwise_count(qq[None, :] ^ codes).sum(-1).astype(np.int32) for qq in q])


@pytest.mark.parametrize("n,nq", [(1, 33), (127, 40), (128, 128), (129, 64), (5000, 200), (40000, 5), (300001, 130)])
def test_mma_distance_matrix_exact(V, monkeypatch, n, nq):
    """Every accumulator element: popc(q) - dot(+-1 query, {0,1} code) == Hamming distance."""
    set_env(monkeypatch, "force_mma_small_batch")
    rng = np.random.default_rng(n * 1000 + nq)
    codes = rng.integers(0, 256, (n, 128), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, 128), dtype=np.uint8)
    if n > 2:
        codes[0] = 0
        codes[1] = 255
        q[0] = 0
        q[-1] = 255
    ix = V.BinaryIndex(1024)
    ix.add_with_ids(codes, np.arange(n))
    d = ix.distances(q)
    assert d.shape == (nq, n)
    assert np.array_equal(d, ref_distances(q, codes))


@pytest.mark.parametrize("cfg", [("1", "4", "1"), ("2", "4", "0"), ("4", "8", "0"), ("3", "4", "1"), ("1", "8", "0")])
def test_mma_pipeline_variants(V, monkeypatch, cfg):
    """Raw-ring depth, operand kind (e2m1 / int8) and CTA pairs on / off must not change a single distance: the hazards
    are in the hand-offs between TMA, expanders, MMA issuers and epilogue."""
    set_env(monkeypatch, "force_mma_small_batch")
    monkeypatch.setenv("VRQ_MMA_RAW_STAGES", cfg[0])
    monkeypatch.setenv("VRQ_MMA_KIND", cfg[1])
    monkeypatch.setenv("VRQ_MMA_PAIR", cfg[2])
    rng = np.random.default_rng(7)
    n, nq = 200000, 256
    codes = rng.integers(0, 256, (n, 128), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, 128), dtype=np.uint8)
    ix = V.BinaryIndex(1024)
    ix.add_with_ids(codes, np.arange(n))
    d = ix.distances(q)
    sel = [0, 1, 63, 64, 127, 128, 200, 255]
    assert np.array_equal(d[sel], ref_distances(q[sel], codes))


@pytest.mark.parametrize("env", ["default", "int8_kind", "single_ctas", "forced_fallback", "no_sampling", "integer_pipes"])
def test_mma_topk_matches_oracle_all_paths(V, monkeypatch, env):
    set_env(monkeypatch, env)
    n, nq = 2_000_000, 256  # two 128-query tiles: the default path runs them as one CTA pair
    codes, _ = oc.synth_codes_int8(61, 0, n, want_int8=False)
    qx = oc.synth_f32(62, 0, nq)
    q = o.synth_ubinary_from_f32(qx)
    q[0] = codes[n - 5]  # an exact hit near the end of the last strip
    ix = V.BinaryIndex(1024)
    ix.add_with_ids(codes, np.arange(n))
    for k in (10, 100, 1000):
        dist, labels = ix.search(q, k)
        rd, rp = oc.hamming_topk(codes, q, k)
        assert np.array_equal(dist, rd)
        assert np.array_equal(labels, rp)


@pytest.mark.parametrize("env", ["default", "no_sampling"])
def test_mma_massive_ties(V, monkeypatch, env):
    """Thousands of codes at the k-th distance: ties resolve by ascending position across strips and query tiles."""
    set_env(monkeypatch, env)
    rng = np.random.default_rng(42)
    n = 700000
    base = rng.integers(0, 256, (16, 128), dtype=np.uint8)
    codes = base[rng.integers(0, 16, n)]
    q = np.concatenate([base[:3], rng.integers(0, 256, (37, 128), dtype=np.uint8)])
    ix = V.BinaryIndex(1024)
    ix.add_with_ids(codes, np.arange(n))
    for k in (1, 100, 1000, 4096):
        dist, labels = ix.search(q, k)
        rd, rp = oc.hamming_topk(codes, q, k)
        assert np.array_equal(dist, rd) and np.array_equal(labels, rp)
    codes[:] = base[0]  # all-equal database: every candidate ties
    ix2 = V.BinaryIndex(1024)
    ix2.add_with_ids(codes, np.arange(n))
    dist, labels = ix2.search(q[:40], 1000)
    rd, rp = oc.hamming_topk(codes, q[:40], 1000)
    assert np.array_equal(dist, rd) and np.array_equal(labels, rp)


def test_mma_sorted_database_worst_case(V, monkeypatch):
    """Database sorted by DEcreasing distance to the queries' common ancestor: a strided sample is still representative,
    but every strip ends with its best rows (stress for list compaction under the sampled threshold)."""
    set_env(monkeypatch, "default")
    rng = np.random.default_rng(4)
    n = 600000
    q0 = rng.integers(0, 256, (1, 128), dtype=np.uint8)
    flips = np.sort(rng.integers(0, 1024, n))[::-1]
    bits = np.unpackbits(np.repeat(q0, n, 0), axis=1)
    mask = np.arange(1024)[None, :] < flips[:, None]
    codes = np.packbits(bits ^ mask, axis=1)
    q = np.repeat(q0, 48, 0)
    q[1:] ^= rng.integers(0, 256, (47, 128), dtype=np.uint8) & rng.integers(0, 256, (47, 128), dtype=np.uint8) & 1
    ix = V.BinaryIndex(1024)
    ix.add_with_ids(codes, np.arange(n))
    for k in (77, 1000):
        dist, labels = ix.search(q, k)
        rd, rp = oc.hamming_topk(codes, q, k)
        assert np.array_equal(dist, rd) and np.array_equal(labels, rp)


def test_mma_and_integer_kernels_agree_with_ids(V, monkeypatch):
    """Same index, same queries, both kernels: identical (distance, label) lists, explicit (non-contiguous) ids."""
    rng = np.random.default_rng(99)
    n, nq, k = 1_000_000, 256, 500
    codes = rng.integers(0, 256, (n, 128), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, 128), dtype=np.uint8)
    ids = rng.permutation(n).astype(np.int64) * 3 + 11
    ix = V.BinaryIndex(1024)
    ix.add_with_ids(codes, ids)
    set_env(monkeypatch, "default")
    a = ix.search(q, k)
    set_env(monkeypatch, "integer_pipes")
    b = ix.search(q, k)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    rd, rp = oc.hamming_topk(codes, q[:8], k)
    assert np.array_equal(a[0][:8], rd) and np.array_equal(a[1][:8], ids[rp])


# ---- <= 32 queries per pass: the swapped-operand kernel (database rows = MMA M, expanded straight into tensor memory, bias column) ----
@pytest.mark.parametrize("n,nq", [(1, 6), (127, 8), (129, 7), (5000, 17), (40000, 32), (300001, 24), (70000, 3)])
def test_few_queries_distance_matrix_exact(V, monkeypatch, n, nq):
    set_env(monkeypatch, "default")
    rng = np.random.default_rng(n * 77 + nq)
    codes = rng.integers(0, 256, (n, 128), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, 128), dtype=np.uint8)
    q[0] = 0
    q[-1] = 255
    ix = V.BinaryIndex(1024)
    ix.add_with_ids(codes, np.arange(n))
    assert np.array_equal(ix.distances(q), ref_distances(q, codes))


@pytest.mark.parametrize("env", ["default", "forced_fallback", "no_sampling"])
@pytest.mark.parametrize("nq", [3, 6, 20, 32])
def test_few_queries_topk_matches_oracle(V, monkeypatch, env, nq):
    set_env(monkeypatch, env)
    n = 2_000_000
    codes, _ = oc.synth_codes_int8(61, 0, n, want_int8=False)
    q = o.synth_ubinary_from_f32(oc.synth_f32(62, 0, nq))
    q[0] = codes[n - 5]
    ix = V.BinaryIndex(1024)
    ix.add_with_ids(codes, np.arange(n))
    for k in (10, 1000):
        dist, labels = ix.search(q, k)
        rd, rp = oc.hamming_topk(codes, q, k)
        assert np.array_equal(dist, rd) and np.array_equal(labels, rp)
    monkeypatch.setenv("VRQ_MMA_FEW", "0")  # the 128-query-tile kernel on the same batch
    d2, l2 = ix.search(q, 1000)
    assert np.array_equal(d2, rd) and np.array_equal(l2, rp)


def test_few_queries_ties_and_sorted_database(V, monkeypatch):
    set_env(monkeypatch, "default")
    rng = np.random.default_rng(43)
    n = 500000
    base = rng.integers(0, 256, (16, 128), dtype=np.uint8)
    codes = base[rng.integers(0, 16, n)]
    q = np.concatenate([base[:3], rng.integers(0, 256, (9, 128), dtype=np.uint8)])
    ix = V.BinaryIndex(1024)
    ix.add_with_ids(codes, np.arange(n))
    for k in (1, 100, 4096):
        dist, labels = ix.search(q, k)
        rd, rp = oc.hamming_topk(codes, q, k)
        assert np.array_equal(dist, rd) and np.array_equal(labels, rp)
    codes[:] = base[0]
    ix2 = V.BinaryIndex(1024)
    ix2.add_with_ids(codes, np.arange(n))
    dist, labels = ix2.search(q, 1000)
    rd, rp = oc.hamming_topk(codes, q, 1000)
    assert np.array_equal(dist, rd) and np.array_equal(labels, rp)


# ---- 33 .. 96 queries per pass: the same swapped-operand kernel with four accumulator column groups (33 .. 64) or, 65 .. 96,
# with one A buffer handed over per K half and eight epilogue warps; VRQ_MMA_W128=128 extends the latter to 128 queries ----
@pytest.mark.parametrize("n,nq", [(1, 33), (127, 48), (129, 50), (5000, 64), (40000, 96), (300001, 81)])
def test_mid_queries_distance_matrix_exact(V, monkeypatch, n, nq):
    set_env(monkeypatch, "default")
    monkeypatch.setenv("VRQ_MMA_MID", "1")
    monkeypatch.setenv("VRQ_MMA_W128", "128")
    rng = np.random.default_rng(n * 79 + nq)
    codes = rng.integers(0, 256, (n, 128), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, 128), dtype=np.uint8)
    q[0] = 0
    q[-1] = 255
    ix = V.BinaryIndex(1024)
    ix.add_with_ids(codes, np.arange(n))
    assert np.array_equal(ix.distances(q), ref_distances(q, codes))


@pytest.mark.parametrize("env", ["default", "forced_fallback", "no_sampling"])
@pytest.mark.parametrize("nq", [33, 64, 96, 128])
def test_mid_queries_topk_matches_oracle(V, monkeypatch, env, nq):
    set_env(monkeypatch, env)
    monkeypatch.setenv("VRQ_MMA_MID", "1")
    monkeypatch.setenv("VRQ_MMA_W128", "128")
    n = 2_000_000
    codes, _ = oc.synth_codes_int8(61, 0, n, want_int8=False)
    q = o.synth_ubinary_from_f32(oc.synth_f32(62, 0, nq))
    q[0] = codes[n - 5]
    ix = V.BinaryIndex(1024)
    ix.add_with_ids(codes, np.arange(n))
    for k in (10, 1000):
        dist, labels = ix.search(q, k)
        rd, rp = oc.hamming_topk(codes, q, k)
        assert np.array_equal(dist, rd) and np.array_equal(labels, rp)
    monkeypatch.setenv("VRQ_MMA_MID", "0")  # the 128-query-tile kernel on the same batch
    monkeypatch.setenv("VRQ_MMA_W128", "0")
    d2, l2 = ix.search(q, 1000)
    assert np.array_equal(d2, rd) and np.array_equal(l2, rp)


def test_mid_queries_ties_and_sorted_database(V, monkeypatch):
    set_env(monkeypatch, "default")
    monkeypatch.setenv("VRQ_MMA_MID", "1")
    rng = np.random.default_rng(44)
    n = 500000
    base = rng.integers(0, 256, (16, 128), dtype=np.uint8)
    codes = base[rng.integers(0, 16, n)]
    q = np.concatenate([base[:3], rng.integers(0, 256, (67, 128), dtype=np.uint8)])
    ix = V.BinaryIndex(1024)
    ix.add_with_ids(codes, np.arange(n))
    for k in (1, 100, 4096):
        dist, labels = ix.search(q, k)
        rd, rp = oc.hamming_topk(codes, q, k)
        assert np.array_equal(dist, rd) and np.array_equal(labels, rp)


@pytest.mark.parametrize("tail,lock", [("0", "0"), ("1", "0"), ("1", "64"), ("0", "64")])
def test_pair_scheduler_variants(V, monkeypatch, tail, lock):
    """1024 queries = 4 CTA-pair columns: the pair scheduler (full strips + the tail strip shared by the left-over pairs,
    VRQ_MMA_TAIL) and the lockstep throttle of the TMA producers (VRQ_MMA_LOCKSTEP) must not change a single key."""
    set_env(monkeypatch, "default")
    monkeypatch.setenv("VRQ_MMA_TAIL", tail)
    monkeypatch.setenv("VRQ_MMA_LOCKSTEP", lock)
    n, nq, k = 3_000_017, 1024, 1000
    codes, _ = oc.synth_codes_int8(63, 0, n, want_int8=False)
    q = o.synth_ubinary_from_f32(oc.synth_f32(64, 0, nq))
    q[5] = codes[n - 3]
    q[700] = codes[0]
    ix = V.BinaryIndex(1024)
    ix.add_with_ids(codes, np.arange(n))
    dist, labels = ix.search(q, k)
    sel = [0, 5, 127, 128, 300, 511, 512, 700, 1023]
    rd, rp = oc.hamming_topk(codes, q[sel], k)
    assert np.array_equal(dist[sel], rd) and np.array_equal(labels[sel], rp)
    key = dist.astype(np.int64) << 40 | labels
    assert np.all(np.diff(key, axis=1) > 0)
    d768, l768 = ix.search(q[:768], k)  # 3 pair columns: the left-over pairs do not divide them -> fewer tail pairs
    assert np.array_equal(d768, dist[:768]) and np.array_equal(l768, labels[:768])


@pytest.mark.parametrize("nq", [12, 40, 256])
def test_mma_clustered_rows_partial_fallback(V, monkeypatch, nq):
    """A tight cluster that the strided sample happens to cover (the first 2048 rows are near-copies of query 0) makes the
    sampled threshold of THAT query far too small: the verification must catch it and the fallback pass must redo it (with
    an infinite threshold for that query only) while every other query keeps its result."""
    set_env(monkeypatch, "default")
    rng = np.random.default_rng(17)
    n = 3_000_000
    codes = rng.integers(0, 256, (n, 128), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, 128), dtype=np.uint8)
    bits = np.unpackbits(np.repeat(q[:1], 2048, 0), axis=1)
    flips = rng.integers(0, 1024, (2048, 40))
    nflip = np.arange(2048) % 40
    for i in range(2048):
        bits[i, flips[i, :nflip[i]]] ^= 1
    codes[:2048] = np.packbits(bits, axis=1)
    ix = V.BinaryIndex(1024)
    ix.add_with_ids(codes, np.arange(n))
    for k in (100, 1000):
        dist, labels = ix.search(q, k)
        rd, rp = oc.hamming_topk(codes, q, k)
        assert np.array_equal(dist, rd) and np.array_equal(labels, rp)


# ---- the two epilogue loops of the dense pass: lean (128 registers, VRQ_MMA_VAR=4, default) and generic (VRQ_MMA_VAR=0) ----
@pytest.mark.parametrize("n,nq", [(300001, 256), (1_000_003, 1024), (70000, 128), (250000, 384)])
def test_lean_and_generic_epilogue_agree(V, monkeypatch, n, nq):
    """CTA pairs (even numbers of 128-query tiles) and single CTAs (odd), a last tile that is partly past the end of the
    database, a tail strip shared by two clusters: both loops return the oracle's top-k."""
    set_env(monkeypatch, "default")
    codes, _ = oc.synth_codes_int8(71, 0, n, want_int8=False)
    q = o.synth_ubinary_from_f32(oc.synth_f32(72, 0, nq))
    q[0] = codes[n - 1]  # an exact hit in the last, partial tile
    ix = V.BinaryIndex(1024)
    ix.add_with_ids(codes, np.arange(n))
    k = 200
    a = ix.search(q, k)
    monkeypatch.setenv("VRQ_MMA_VAR", "0")
    b = ix.search(q, k)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    rd, rp = oc.hamming_topk(codes, q[:12], k)
    assert np.array_equal(a[0][:12], rd) and np.array_equal(a[1][:12], rp)
    assert a[1][0, 0] == n - 1 and a[0][0, 0] == 0
