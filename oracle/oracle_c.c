/* CPU oracle, C half.  TEST INFRASTRUCTURE ONLY - never linked into or called by the product.
 *
 * Plain-C restatement of the integer / bit-exact pieces of the reference path, written independently of
 * oracle/vrq_oracle.py so the two check each other, and fast enough (OpenMP over queries, the way faiss
 * parallelises IndexBinaryFlat::search) to serve as the timed CPU baseline in bench.py.
 *
 * Reference lines restated (all in /root/reference):
 *   vrqo_pairwise_sum_f32      np.mean's float32 pairwise sum, used by every _to_binary (VectorDBInt8.py:146)
 *   vrqo_to_binary_f32         VectorDBInt8.py:140-146 (+ siblings), ge=1: CohereVectorDBBinary.py:133-151
 *   vrqo_to_binary_i8/_i16     CohereVectorDBInt8.py:130-135, VectorDBInt16.py:148-157
 *   vrqo_quantize_int8_perdoc  VectorDBInt8.py:114-126
 *   vrqo_quantize_int8_global  VectorDBInt8Global.py:130-142
 *   vrqo_quantize_int16_global VectorDBInt16Global.py:130-142
 *   vrqo_quantize_int4         VectorDBInt4.py:116-154 (== VectorDBInt4Global.py:129-164)
 *   vrqo_hamming_topk          faiss IndexBinaryFlat::search as called at CohereEnhancedVectorDB.py:268;
 *                              faiss-cpu is unpinned and absent: published algorithm (max-heap of k, strict
 *                              '<' replace, heap ordered by (dist, pos), final ascending reorder).
 *   vrqo_rescore_binary        CohereEnhancedVectorDB.py:288-290 (float64 accumulation)
 *   vrqo_int8_sumsq            CohereEnhancedVectorDB.py:308 (exact integer part of the norm)
 *
 * Build: see oracle/Makefile (-ffp-contract=off, no fast-math: every float op is a single IEEE rounding).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

float vrqo_pairwise_sum_f32(const float* a, int64_t n) {
    if (n < 8) {
        float r = (n == 0) ? 0.0f : -0.0f;
        for (int64_t i = 0; i < n; i++) r += a[i];
        return r;
    }
    if (n <= 128) {
        float r[8];
        for (int j = 0; j < 8; j++) r[j] = a[j];
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    return vrqo_pairwise_sum_f32(a, n2) + vrqo_pairwise_sum_f32(a + n2, n - n2);
}

static void packbits_row(const uint8_t* bits, int d, uint8_t* out) {
    int nb = (d + 7) / 8;
    for (int b = 0; b < nb; b++) {
        uint8_t v = 0;
        for (int j = 0; j < 8; j++) {
            int i = b * 8 + j;
            v = (uint8_t)(v << 1) | (uint8_t)((i < d) ? bits[i] : 0);
        }
        out[b] = v;
    }
}

void vrqo_to_binary_f32(const float* x, int64_t n, int d, int ge, uint8_t* out) {
    int nb = (d + 7) / 8;
#pragma omp parallel
    {
        uint8_t* bits = (uint8_t*)malloc((size_t)d);
#pragma omp for schedule(static)
        for (int64_t r = 0; r < n; r++) {
            const float* row = x + r * (int64_t)d;
            float mean = vrqo_pairwise_sum_f32(row, d) / (float)d;
            for (int i = 0; i < d; i++) bits[i] = ge ? (row[i] >= mean) : (row[i] > mean);
            packbits_row(bits, d, out + r * (int64_t)nb);
        }
        free(bits);
    }
}

#define TO_BINARY_INT(NAME, T)                                                              \
    void NAME(const T* x, int64_t n, int d, int ge, uint8_t* out) {                         \
        int nb = (d + 7) / 8;                                                               \
        _Pragma("omp parallel") {                                                           \
            uint8_t* bits = (uint8_t*)malloc((size_t)d);                                    \
            _Pragma("omp for schedule(static)") for (int64_t r = 0; r < n; r++) {           \
                const T* row = x + r * (int64_t)d;                                          \
                int64_t s = 0;                                                              \
                for (int i = 0; i < d; i++) s += row[i];                                    \
                for (int i = 0; i < d; i++) {                                               \
                    int64_t v = (int64_t)d * row[i];                                        \
                    bits[i] = ge ? (v >= s) : (v > s);                                      \
                }                                                                           \
                packbits_row(bits, d, out + r * (int64_t)nb);                               \
            }                                                                               \
            free(bits);                                                                     \
        }                                                                                   \
    }
TO_BINARY_INT(vrqo_to_binary_i8, int8_t)
TO_BINARY_INT(vrqo_to_binary_i16, int16_t)

static void row_minmax(const float* row, int d, float* lo, float* hi) {
    float a = row[0], b = row[0];
    for (int i = 1; i < d; i++) {
        if (row[i] < a) a = row[i];
        if (row[i] > b) b = row[i];
    }
    *lo = a;
    *hi = b;
}

void vrqo_quantize_int8_perdoc(const float* x, int64_t n, int d, int8_t* q, float* lo, float* hi) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; r++) {
        const float* row = x + r * (int64_t)d;
        int8_t* o = q + r * (int64_t)d;
        float a, b;
        row_minmax(row, d, &a, &b);
        lo[r] = a;
        hi[r] = b;
        if (a == b) {
            memset(o, 0, (size_t)d);
            continue;
        }
        float m = fmaxf(fabsf(a), fabsf(b));
        float scale = 127.0f / m; /* float32 division */
        for (int i = 0; i < d; i++) o[i] = (int8_t)(int32_t)(row[i] * scale); /* truncation */
    }
}

static void global_quant(const float* x, int64_t n, int d, double limit, double qmax, int8_t* q8, int16_t* q16) {
    const float lim = (float)limit;            /* np.clip bounds: float32(limit) */
    const float scale = (float)(qmax / limit); /* float64 divide, one rounding  */
    const float qm = (float)qmax;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < n * (int64_t)d; e++) {
        float v = x[e];
        v = v < -lim ? -lim : (v > lim ? lim : v);
        float s = rintf(v * scale); /* round-half-to-even (default rounding mode) */
        s = s < -qm ? -qm : (s > qm ? qm : s);
        if (q8) q8[e] = (int8_t)(int32_t)s;
        else q16[e] = (int16_t)(int32_t)s;
    }
}
void vrqo_quantize_int8_global(const float* x, int64_t n, int d, double limit, int8_t* q) {
    global_quant(x, n, d, limit, 127.0, q, 0);
}
void vrqo_quantize_int16_global(const float* x, int64_t n, int d, double limit, int16_t* q) {
    global_quant(x, n, d, limit, 32767.0, 0, q);
}

void vrqo_quantize_int4(const float* x, int64_t n, int d, int8_t* packed, double* lo, double* hi) {
    int np_ = (d + 1) / 2;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; r++) {
        const float* row = x + r * (int64_t)d;
        int8_t* o = packed + r * (int64_t)np_;
        float a, b;
        row_minmax(row, d, &a, &b);
        lo[r] = (double)a;
        hi[r] = (double)b;
        if (a == b) {
            memset(o, 0, (size_t)np_);
            continue;
        }
        double m = fmax(fabs((double)a), fabs((double)b));
        float scale = (float)(7.0 / m); /* float64 divide then cast */
        for (int i = 0; i < np_; i++) {
            float sa = rintf(row[2 * i] * scale);
            sa = sa < -8.f ? -8.f : (sa > 7.f ? 7.f : sa);
            int va = (int)sa + 8, vb = 0;
            if (2 * i + 1 < d) {
                float sb = rintf(row[2 * i + 1] * scale);
                sb = sb < -8.f ? -8.f : (sb > 7.f ? 7.f : sb);
                vb = (int)sb + 8;
            }
            o[i] = (int8_t)(uint8_t)(((va & 0xF) << 4) | (vb & 0xF));
        }
    }
}

/* ---- Hamming top-k: faiss-style binary max-heap keyed on (dist, pos) -------------------------------- */
typedef struct {
    int32_t d;
    int64_t p;
} ent_t;
static inline int ent_gt(ent_t a, ent_t b) { return a.d > b.d || (a.d == b.d && a.p > b.p); }

static void heap_sift_down(ent_t* h, int k, int i) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < k && ent_gt(h[l], h[m])) m = l;
        if (r < k && ent_gt(h[r], h[m])) m = r;
        if (m == i) return;
        ent_t t = h[i];
        h[i] = h[m];
        h[m] = t;
        i = m;
    }
}
static int ent_cmp(const void* a, const void* b) {
    const ent_t *x = (const ent_t*)a, *y = (const ent_t*)b;
    if (x->d != y->d) return x->d < y->d ? -1 : 1;
    return x->p < y->p ? -1 : (x->p > y->p ? 1 : 0);
}

static inline int32_t hamming(const uint8_t* a, const uint8_t* b, int nbytes) {
    int32_t acc = 0;
    int i = 0;
    for (; i + 8 <= nbytes; i += 8) {
        uint64_t x, y;
        memcpy(&x, a + i, 8);
        memcpy(&y, b + i, 8);
        acc += __builtin_popcountll(x ^ y);
    }
    for (; i < nbytes; i++) acc += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return acc;
}

int vrqo_hamming_topk(const uint8_t* codes, int64_t n, int code_bytes, const uint8_t* q, int nq, int k,
                      int64_t pos_base, int32_t* dist, int64_t* pos, int nthreads) {
    if (k <= 0) return 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int qi = 0; qi < nq; qi++) {
        ent_t* h = (ent_t*)malloc(sizeof(ent_t) * (size_t)k);
        for (int i = 0; i < k; i++) {
            h[i].d = INT32_MAX;
            h[i].p = INT64_MAX; /* sentinel: any real entry replaces it */
        }
        const uint8_t* qc = q + (int64_t)qi * code_bytes;
        for (int64_t j = 0; j < n; j++) {
            int32_t d = hamming(qc, codes + j * (int64_t)code_bytes, code_bytes);
            if (d < h[0].d) { /* strict, like faiss: a tie with the current worst never enters */
                h[0].d = d;
                h[0].p = pos_base + j;
                heap_sift_down(h, k, 0);
            }
        }
        qsort(h, (size_t)k, sizeof(ent_t), ent_cmp);
        for (int i = 0; i < k; i++) {
            dist[(int64_t)qi * k + i] = h[i].d;
            pos[(int64_t)qi * k + i] = (h[i].d == INT32_MAX) ? -1 : h[i].p;
        }
        free(h);
    }
    return 0;
}

/* Phase II: sum_i q[i] * (2*bit_i - 1), float64 accumulation, bit i = byte i/8, bit 7 - i%8 */
void vrqo_rescore_binary(const float* qf, int d, const uint8_t* cand_codes, int64_t m, double* score) {
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < m; c++) {
        const uint8_t* code = cand_codes + c * (int64_t)(d / 8);
        double acc = 0.0;
        for (int i = 0; i < d; i++) {
            int bit = (code[i >> 3] >> (7 - (i & 7))) & 1;
            acc += bit ? (double)qf[i] : -(double)qf[i];
        }
        score[c] = acc;
    }
}

void vrqo_int8_sumsq(const int8_t* rows, int64_t m, int d, int64_t* out) {
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < m; c++) {
        int64_t s = 0;
        for (int i = 0; i < d; i++) s += (int64_t)rows[c * (int64_t)d + i] * rows[c * (int64_t)d + i];
        out[c] = s;
    }
}

/* ---- counter-based synthetic generator, C twin of vrq_oracle.synth_* ------------------ */
static inline uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline float synth_elem(uint64_t base, uint64_t r, uint64_t c, int d) {
    uint64_t h = splitmix64(base + r * (uint64_t)d + c);
    int64_t s = (int64_t)((h & 0xFFFF) + ((h >> 16) & 0xFFFF) + ((h >> 32) & 0xFFFF) + (h >> 48));
    int64_t mc = (int64_t)(splitmix64(c ^ 0xC01DBEEFCAFEF00DULL) & 0x7FFF) - 16384;
    return (float)(s - 131070 + mc) * 0x1p-20f;
}
void vrqo_synth_f32(uint64_t seed, int64_t row0, int64_t nrows, int d, int row_scale, float* out) {
    uint64_t base = seed * 0xD1342543DE82EF95ULL;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < nrows; r++) {
        uint64_t rr = (uint64_t)(row0 + r);
        float sc = 1.0f;
        if (row_scale) {
            uint64_t hr = splitmix64(base ^ (rr + 0x5851F42D4C957F2DULL));
            sc = ldexpf(1.0f, (int)(hr & 3) - 1);
        }
        for (int c = 0; c < d; c++) out[r * (int64_t)d + c] = synth_elem(base, rr, (uint64_t)c, d) * sc;
    }
}
/* Cohere-like code + int8 rows for the search benchmark, generated without materialising the floats */
void vrqo_synth_codes_int8(uint64_t seed, int64_t row0, int64_t nrows, int d, uint8_t* codes, int8_t* i8) {
    uint64_t base = seed * 0xD1342543DE82EF95ULL;
    int nb = d / 8;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < nrows; r++) {
        uint64_t rr = (uint64_t)(row0 + r);
        for (int b = 0; b < nb; b++) {
            uint8_t v = 0;
            for (int j = 0; j < 8; j++) {
                int c = b * 8 + j;
                float x = synth_elem(base, rr, (uint64_t)c, d);
                v = (uint8_t)(v << 1) | (uint8_t)(x > 0.0f);
                if (i8) {
                    float t = x * 1259.0f;
                    t = t - 0.69f;
                    t = rintf(t);
                    t = t < -128.f ? -128.f : (t > 127.f ? 127.f : t);
                    i8[r * (int64_t)d + c] = (int8_t)(int32_t)t;
                }
            }
            if (codes) codes[r * (int64_t)nb + b] = v;
        }
    }
}

void vrqo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int vrqo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
