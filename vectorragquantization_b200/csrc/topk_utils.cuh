// Block-level exact selection and sorting primitives on 64-bit keys (shared + global memory).
// Keys are unique by construction ((hamming << 40) | position, or (score, rank)), so "the k smallest keys" is a
// well-defined set and every selection here is exact and deterministic.
#pragma once
#include <stdint.h>

namespace vrq {

template <int NT>
__device__ __forceinline__ void group_sync(int bar_id) {
    // named barrier over NT threads (bar 0 == __syncthreads when NT == blockDim)
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NT) : "memory");
}

struct SelectScratch {
    uint32_t hist[256];
    unsigned long long prefix;
    unsigned long long mask;
    unsigned long long kmin, kmax;
    int kk;
    int counter;
};

// k-th smallest (1-based, 1 <= k <= n) of the n keys load(0) .. load(n-1).  All NT threads of the group must call.
// MSB-first radix select, 8 bits per pass, with two refinements that matter for (distance << 40 | position) keys:
//   * the passes start at the first byte in which the keys actually differ (min ^ max), so the all-equal top bytes
//     cost nothing;
//   * lanes that hold the same digit are merged with match.any before the shared-memory atomic, so a digit on which
//     most keys agree (the distance byte near the threshold) does not serialise the whole block on one address.
// min_shift > 0 stops the refinement above that bit: the result is the k-th smallest value of (key >> min_shift) with all
// lower bits SET, i.e. "key <= result" keeps every key that ties with the k-th one in its upper bits (>= k keys).
template <int NT, class Load>
__device__ unsigned long long radix_select_kth(Load load, int n, int k, int tid, SelectScratch* sc, int bar_id, int min_shift = 0) {
    const int lane = tid & 31;
    group_sync<NT>(bar_id);  // nobody is still reading the scratch of a previous call
    if (tid == 0) {
        sc->kmin = ~0ull;
        sc->kmax = 0ull;
        sc->kk = k;
    }
    group_sync<NT>(bar_id);
    // (loads are issued four at a time: when load() reads global memory a dependent one-key-per-trip loop is pure latency)
    unsigned long long lo = ~0ull, hi = 0ull;
    for (int i = tid; i < n; i += 4 * NT) {
        unsigned long long key[4];
#pragma unroll
        for (int u = 0; u < 4; u++) key[u] = (i + u * NT < n) ? load(i + u * NT) : load(i);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            lo = key[u] < lo ? key[u] : lo;
            hi = key[u] > hi ? key[u] : hi;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    if (lane == 0) {
        atomicMin(&sc->kmin, lo);
        atomicMax(&sc->kmax, hi);
    }
    group_sync<NT>(bar_id);
    const unsigned long long diff = sc->kmin ^ sc->kmax;
    if (diff == 0ull) return sc->kmin;  // every key identical
    if (min_shift > 0 && (diff >> min_shift) == 0ull) return sc->kmin | ((1ull << min_shift) - 1ull);  // identical upper bits
    const int top_shift = ((63 - __clzll((long long)diff)) >> 3) << 3;
    if (tid == 0) {
        sc->mask = (top_shift >= 56) ? 0ull : (~0ull << (top_shift + 8));
        sc->prefix = sc->kmin & sc->mask;
    }
    const int n_round = ((n + 31) >> 5) << 5;  // whole warps stay converged for match.any
    for (int shift = top_shift; shift >= min_shift; shift -= 8) {
        for (int i = tid; i < 256; i += NT) sc->hist[i] = 0;
        group_sync<NT>(bar_id);
        const unsigned long long prefix = sc->prefix, mask = sc->mask;
#ifdef VRQ_SELECT_MATCH_ANY
        for (int i = tid; i < n_round; i += NT) {
            unsigned long long key = 0ull;
            bool ok = false;
            if (i < n) {
                key = load(i);
                ok = (key & mask) == prefix;
            }
            const unsigned digit = ok ? ((unsigned)(key >> shift) & 255u) : (256u + (unsigned)lane);
            const unsigned peers = __match_any_sync(0xffffffffu, digit);
            if (ok && lane == (__ffs(peers) - 1)) atomicAdd(&sc->hist[digit], (uint32_t)__popc(peers));
        }
#else
        for (int i = tid; i < n; i += 4 * NT) {
            unsigned long long key[4];
#pragma unroll
            for (int u = 0; u < 4; u++) key[u] = (i + u * NT < n) ? load(i + u * NT) : ~prefix;  // ~prefix never matches
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (i + u * NT < n && (key[u] & mask) == prefix) atomicAdd(&sc->hist[(unsigned)(key[u] >> shift) & 255u], 1u);
        }
#endif
        group_sync<NT>(bar_id);
        if (tid < 32) {
            // 8 bins per lane, inclusive scan over lanes, locate the bin holding rank kk
            uint32_t c[8], s = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                c[i] = sc->hist[tid * 8 + i];
                s += c[i];
            }
            uint32_t inc = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
                if (tid >= o) inc += v;
            }
            const uint32_t kk = (uint32_t)sc->kk;
            const uint32_t before = inc - s;
            const bool mine = (before < kk) && (kk <= inc);
            __syncwarp();
            if (mine) {
                uint32_t acc = before;
                int bin = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (acc + c[i] < kk) {
                        acc += c[i];
                        bin = i + 1;
                    } else {
                        break;
                    }
                }
                sc->kk = (int)(kk - acc);
                sc->prefix = prefix | ((unsigned long long)(tid * 8 + bin) << shift);
                sc->mask = mask | (0xFFull << shift);
            }
        }
        group_sync<NT>(bar_id);
    }
    return min_shift > 0 ? (sc->prefix | ((1ull << min_shift) - 1ull)) : sc->prefix;
}

// In-place ascending bitonic sort of n2 (power of two) keys in shared memory, optional 32-bit payload.
template <int NT, bool HAS_VAL>
__device__ void bitonic_sort(unsigned long long* keys, uint32_t* vals, int n2, int tid, int bar_id) {
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            group_sync<NT>(bar_id);
            for (int i = tid; i < (n2 >> 1); i += NT) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool asc = ((lo & size) == 0);
                unsigned long long a = keys[lo], b = keys[hi];
                bool sw;
                if (HAS_VAL) {
                    uint32_t va = vals[lo], vb = vals[hi];
                    const bool gt = (a > b) || (a == b && va > vb);
                    sw = (gt == asc);
                    if (sw && (a != b || va != vb)) {
                        keys[lo] = b;
                        keys[hi] = a;
                        vals[lo] = vb;
                        vals[hi] = va;
                    }
                } else {
                    sw = ((a > b) == asc);
                    if (sw && a != b) {
                        keys[lo] = b;
                        keys[hi] = a;
                    }
                }
            }
        }
    }
    group_sync<NT>(bar_id);
}

__device__ __forceinline__ int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// Monotone map double -> u64 (ascending); -0.0 is folded onto +0.0 first so that it ties with +0.0 the way
// Python's float comparison does in list.sort.
__device__ __forceinline__ unsigned long long ordered_from_double(double v) {
    v = v + 0.0;
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ unsigned long long ordered_from_float(float v) {
    v = v + 0.0f;
    uint32_t b = __float_as_uint(v);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return (unsigned long long)b;
}

}  // namespace vrq
