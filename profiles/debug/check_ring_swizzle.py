"""Checks the shared-memory swizzle of encode1024_ring_kernel (encode.cu) on the CPU: the XOR is a permutation, the three
access patterns are bank-conflict free, and the split address formulas the kernel uses equal chunk ^ ring_swz(chunk >> 3)."""


def sw(line):
    return (line & 7) ^ (((line >> 3) & 1) << 1)


def phys(chunk):
    return chunk ^ sw(chunk >> 3)


def worst_conflict(addr_lists, width):
    worst = 0
    for addrs in addr_lists:
        banks = {}
        for a in addrs:
            for w in range(0, width, 4):
                bk = ((a + w) // 4) % 32
                banks[bk] = banks.get(bk, 0) + 1
        worst = max(worst, max(banks.values()))
    return worst


assert sorted(phys(c) for c in range(256)) == list(range(256))
# mean tree: LDS.64, half-warp phases
p1 = [[phys(32 * (l >> 2) + 2 * s + ((l & 3) >> 1)) * 16 + (l & 1) * 8 for l in range(16 * h, 16 * h + 16)] for s in range(16) for h in range(2)]
assert worst_conflict(p1, 8) == 1
# encode loads (LDS.128, quarter-warp phases) for K = 1, 2, 4 and the cp.async fill pattern
for K in (1, 2, 4):
    cpr = 8 // K
    p2 = [[phys((i // cpr) * (256 // K) + l * cpr + i % cpr) * 16 for l in range(8 * q, 8 * q + 8)] for i in range(8) for q in range(4)]
    assert worst_conflict(p2, 16) == 1, K
fill = [[phys(32 * j + l) * 16 for l in range(8 * q, 8 * q + 8)] for j in range(8) for q in range(4)]
assert worst_conflict(fill, 16) == 1

# split formulas: (lane part) ^ (compile-time part) + (compile-time add)
for lane in range(32):
    fl = (lane ^ (lane >> 3)) * 16
    for j in range(8):
        assert 512 * j + (fl ^ ((4 * (j & 1) ^ 2 * ((j >> 1) & 1)) * 16)) == phys(32 * j + lane) * 16
    b, jp = lane >> 2, lane & 3
    h = jp >> 1
    l1 = 512 * b + ((h ^ 4 * (b & 1) ^ 2 * ((b >> 1) & 1)) * 16) + (jp & 1) * 8
    for s in range(16):
        assert (l1 ^ (((2 * (s & 3)) ^ (s >> 2)) * 16)) + (s >> 2) * 128 == phys(32 * b + 2 * s + h) * 16 + (jp & 1) * 8
    for K in (1, 2, 4):
        cpr = 8 // K
        if K == 1:
            l2 = ((8 * lane) ^ ((lane & 7) ^ (((lane >> 3) & 1) << 1))) * 16
        elif K == 2:
            l2 = ((4 * lane) ^ (((lane >> 1) & 7) ^ (((lane >> 4) & 1) << 1))) * 16
        else:
            l2 = ((2 * lane) ^ ((lane >> 2) & 7)) * 16
        for i in range(8):
            k, u = i // cpr, i % cpr
            a = 1024 * k + (l2 ^ ((u ^ 2 * (k & 1)) * 16)) if K == 4 else (4096 // K) * k + (l2 ^ (u * 16))
            assert a == phys(k * (256 // K) + lane * cpr + u) * 16
print("ring swizzle: permutation, conflict-free, split formulas exact")
