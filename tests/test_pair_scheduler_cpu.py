"""The CTA-pair scheduler of the tensor-core scan (scan_mma.cu: plan_scan_mma + the kernel's segment()), restated in Python.

P = sms / 2 clusters serve cols = qtiles / 2 query-tile pairs: F full strips per column, the E clusters left over share one tail
strip (each walks cols / E columns one after the other).  Checked here: every (column, tile) is scanned exactly once, no
cluster is idle, and the work per cluster is balanced to within one strip's rounding."""
import itertools


def plan(sms, qtiles, tiles):
    P, cols = sms // 2, qtiles // 2
    F, E = P // cols, P % cols
    while E > 0 and cols % E != 0:
        E -= 1
    if F < 1:
        F, E = 1, 0
    T1 = max(1, -(-tiles * cols // (E + F * cols)))
    if T1 * F >= tiles:  # nothing left for a tail strip (small inputs)
        E = 0
        F = -(-tiles // T1)
    return cols, F, E, T1


def segments(cluster, cols, F, E, T1, tiles):
    """(column, first tile, number of tiles) of every segment of a cluster - the kernel's segment() lambda."""
    nfull = cols * F
    if cluster < nfull:
        col, strip = cluster % cols, cluster // cols
        t0 = strip * T1
        return [(col, t0, max(0, min(T1, tiles - t0)))]
    per = cols // E
    t0 = F * T1
    return [((cluster - nfull) * per + s, t0, max(0, tiles - t0)) for s in range(per)]


def test_every_tile_of_every_column_exactly_once():
    for sms, qtiles, tiles in itertools.product((148, 132, 16), (2, 4, 8, 16), (1, 7, 100, 781250, 7812500, 999983)):
        cols, F, E, T1 = plan(sms, qtiles, tiles)
        if cols > sms // 2:
            continue  # more columns than clusters: the classic grid is used instead
        nclusters = cols * F + E
        assert nclusters <= sms // 2
        seen = {}
        load = []
        for c in range(nclusters):
            segs = segments(c, cols, F, E, T1, tiles)
            load.append(sum(n for _, _, n in segs))
            for col, t0, n in segs:
                assert 0 <= col < cols
                if n:
                    seen.setdefault(col, []).append((t0, t0 + n))
        for col in range(cols):
            spans = sorted(seen.get(col, []))
            assert spans and spans[0][0] == 0 and spans[-1][1] == tiles, (sms, qtiles, tiles, col, spans)
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]  # contiguous, no overlap
        if tiles >= 10 * nclusters:
            assert min(load) > 0 and max(load) <= min(load) + 2 * T1 // max(1, F) + cols, (sms, qtiles, tiles, min(load), max(load))


def test_the_headline_shape_keeps_all_148_sms_busy():
    cols, F, E, T1 = plan(148, 8, 781250)  # 100 M rows, 1024 queries
    assert (cols, F, E) == (4, 18, 2) and cols * F + E == 74
    # 72 clusters own a full strip of one column, 2 clusters walk the tail strip for 2 columns each
    tail = 781250 - F * T1
    assert abs(T1 - 2 * tail) <= 16  # 42230 tiles per full strip, 2 x 21110 per tail cluster
