/*
 * oracle/dpk_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * A plain-C CPU restatement of the reference's shuffle hot path, used only as
 * the checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  Nothing under dpark_b200/ may import, link or call it.
 *
 * Parity is PINNED: tests/test_oracle_golden.py checks every function here
 * against tests/golden/*.json, which tests/golden/make_golden.py captured from
 * the real reference (compiled portable_hash.pyx + DparkContext('local')).
 *
 * Each function cites the reference lines it restates (paths relative to the
 * reference root).  CPython builtins the reference leans on (hash(int),
 * hash(float), %, bisect) are restated from their published algorithms
 * (CPython 3.12 Objects/longobject.c long_hash, Python/pyhash.c
 * _Py_HashDouble, Modules/_bisectmodule.c bisect_right); the reference does not
 * vendor them.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PYHASH_BITS 61
#define PYHASH_MOD ((((uint64_t)1) << PYHASH_BITS) - 1)

/* ------------------------------------------------------------------ a1: hash */

/* hash(int) for values in int64 range. dpark/portable_hash.pyx:61-62 -> CPython
 * long_hash: sign * (abs(x) mod (2**61-1)), and -1 -> -2. */
int64_t orc_hash_i64(int64_t x) {
    uint64_t a = x < 0 ? (uint64_t)0 - (uint64_t)x : (uint64_t)x;
    uint64_t r = a % PYHASH_MOD;
    int64_t h = x < 0 ? -(int64_t)r : (int64_t)r;
    return h == -1 ? -2 : h;
}

/* hash(np.uint64) / hash(int) for values in [0, 2**64). */
int64_t orc_hash_u64(uint64_t x) {
    return (int64_t)(x % PYHASH_MOD); /* never -1 */
}

/* hash(float). dpark/portable_hash.pyx:61-62 -> CPython _Py_HashDouble.
 * NaN is not supported (CPython >= 3.10 hashes NaN by object identity). */
int64_t orc_hash_f64(double v) {
    if (isinf(v)) return v > 0 ? 314159 : -314159;
    if (isnan(v)) return 0;
    int e;
    double m = frexp(v, &e);
    int sign = 1;
    if (m < 0) { sign = -1; m = -m; }
    uint64_t x = 0;
    while (m != 0.0) {
        x = ((x << 28) & PYHASH_MOD) | (x >> (PYHASH_BITS - 28));
        m *= 268435456.0; /* 2**28 */
        e -= 28;
        uint64_t y = (uint64_t)m;
        m -= (double)y;
        x += y;
        if (x >= PYHASH_MOD) x -= PYHASH_MOD;
    }
    e = e >= 0 ? e % PYHASH_BITS : PYHASH_BITS - 1 - ((-1 - e) % PYHASH_BITS);
    x = ((x << e) & PYHASH_MOD) | (x >> (PYHASH_BITS - e));
    int64_t h = sign > 0 ? (int64_t)x : -(int64_t)x;
    return h == -1 ? -2 : h;
}

/* string_hash over SIGNED chars. dpark/portable_hash.pyx:17-32 */
int64_t orc_hash_bytes(const uint8_t *s, int64_t len) {
    if (len == 0) return 0;
    uint64_t value = (uint64_t)(int64_t)(int8_t)s[0] << 7;
    for (int64_t i = 0; i < len; i++)
        value = (1000003ULL * value) ^ (uint64_t)(int64_t)(int8_t)s[i];
    value ^= (uint64_t)len;
    return (int64_t)value == -1 ? -2 : (int64_t)value;
}

/* unicode_hash over code points. dpark/portable_hash.pyx:34-48 */
int64_t orc_hash_codepoints(const uint32_t *s, int64_t len) {
    if (len == 0) return 0;
    uint64_t value = (uint64_t)s[0] << 7;
    for (int64_t i = 0; i < len; i++)
        value = (1000003ULL * value) ^ (uint64_t)s[i];
    value ^= (uint64_t)len;
    return (int64_t)value == -1 ? -2 : (int64_t)value;
}

/* unicode_hash of a str given as UTF-8 bytes (the columnar layout the product
 * uses): decode to code points on the fly, length = number of code points.
 * Assumes valid UTF-8 (as produced by str.encode('utf-8', 'surrogatepass')). */
int64_t orc_hash_utf8(const uint8_t *s, int64_t nbytes) {
    if (nbytes == 0) return 0;
    uint64_t value = 0;
    int64_t ncp = 0, i = 0;
    while (i < nbytes) {
        uint32_t c = s[i], cp;
        int extra;
        if (c < 0x80) { cp = c; extra = 0; }
        else if (c < 0xE0) { cp = c & 0x1F; extra = 1; }
        else if (c < 0xF0) { cp = c & 0x0F; extra = 2; }
        else { cp = c & 0x07; extra = 3; }
        for (int j = 1; j <= extra && i + j < nbytes; j++) cp = (cp << 6) | (s[i + j] & 0x3F);
        i += extra + 1;
        if (ncp == 0) value = (uint64_t)cp << 7;
        value = (1000003ULL * value) ^ (uint64_t)cp;
        ncp++;
    }
    value ^= (uint64_t)ncp;
    return (int64_t)value == -1 ? -2 : (int64_t)value;
}

/* tuple_hash given the item hashes. dpark/portable_hash.pyx:3-15 */
int64_t orc_hash_tuple(const int64_t *item_hashes, int64_t n) {
    uint64_t mul = 1000003ULL, value = 0x345678ULL;
    int64_t l = n;
    for (int64_t i = 0; i < n; i++) {
        l -= 1;
        value = (value ^ (uint64_t)item_hashes[i]) * mul;
        mul += (uint64_t)(int64_t)(82520 + l * 2);
    }
    value += 97531ULL;
    return (int64_t)value == -1 ? -2 : (int64_t)value;
}

/* --------------------------------------------------------- a2: getPartition */

/* HashPartitioner.getPartition. dpark/dependency.py:229-233.
 * Python % is floor-mod: result in [0, P).  With thresholds:
 * bisect.bisect(thresholds, h) = number of thresholds <= h. */
int32_t orc_partition(int64_t h, int32_t P, const int64_t *thresholds, int32_t nthr) {
    if (thresholds == NULL) {
        int64_t m = h % (int64_t)P;
        if (m < 0) m += P;
        return (int32_t)m;
    }
    int32_t lo = 0, hi = nthr;
    while (lo < hi) {
        int32_t mid = (lo + hi) / 2;
        if (h < thresholds[mid]) hi = mid; else lo = mid + 1;
    }
    return lo;
}

/* key kinds for the vector entry points */
enum { ORC_K_I64 = 0, ORC_K_I32 = 1, ORC_K_F64 = 2, ORC_K_U64 = 3, ORC_K_F32 = 4 };

static inline int64_t hash_key_at(const void *keys, int kind, int64_t i) {
    switch (kind) {
    case ORC_K_I64: return orc_hash_i64(((const int64_t *)keys)[i]);
    case ORC_K_I32: return orc_hash_i64((int64_t)((const int32_t *)keys)[i]);
    case ORC_K_F64: return orc_hash_f64(((const double *)keys)[i]);
    case ORC_K_U64: return orc_hash_u64(((const uint64_t *)keys)[i]);
    case ORC_K_F32: return orc_hash_f64((double)((const float *)keys)[i]);
    }
    return 0;
}

void orc_hash_vec(const void *keys, int kind, int64_t n, int64_t *out) {
    for (int64_t i = 0; i < n; i++) out[i] = hash_key_at(keys, kind, i);
}

/* mode 0: bytes (signed chars); mode 1: str stored as UTF-8 */
void orc_hash_bytes_vec(const uint8_t *data, const int64_t *offsets, int64_t n, int mode,
                        int64_t *out) {
    for (int64_t i = 0; i < n; i++) {
        const uint8_t *s = data + offsets[i];
        int64_t len = offsets[i + 1] - offsets[i];
        out[i] = mode == 0 ? orc_hash_bytes(s, len) : orc_hash_utf8(s, len);
    }
}

void orc_partition_vec(const int64_t *hash, int64_t n, int32_t P, const int64_t *thresholds,
                       int32_t nthr, int32_t *out) {
    for (int64_t i = 0; i < n; i++) out[i] = orc_partition(hash[i], P, thresholds, nthr);
}

/* ------------------------------------------- insertion-ordered int64 hash map */
/* Plays the role of the Python dict in task.py:209-226 and shuffle.py:600-608:
 * first-seen iteration order, upsert by key. */
typedef struct {
    int64_t *slots;  /* index into keys[], or -1 */
    int64_t cap;     /* power of two */
    int64_t *keys;
    int64_t n, room;
} omap;

static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    return x ^ (x >> 33);
}
static void omap_init(omap *m, int64_t hint) {
    m->cap = 16;
    while (m->cap < hint * 2) m->cap <<= 1;
    m->slots = (int64_t *)malloc(sizeof(int64_t) * m->cap);
    memset(m->slots, 0xff, sizeof(int64_t) * m->cap);
    m->room = hint > 16 ? hint : 16;
    m->keys = (int64_t *)malloc(sizeof(int64_t) * m->room);
    m->n = 0;
}
static void omap_free(omap *m) { free(m->slots); free(m->keys); }
static void omap_grow(omap *m) {
    m->cap <<= 1;
    m->slots = (int64_t *)realloc(m->slots, sizeof(int64_t) * m->cap);
    memset(m->slots, 0xff, sizeof(int64_t) * m->cap);
    for (int64_t i = 0; i < m->n; i++) {
        uint64_t s = mix64((uint64_t)m->keys[i]) & (uint64_t)(m->cap - 1);
        while (m->slots[s] >= 0) s = (s + 1) & (uint64_t)(m->cap - 1);
        m->slots[s] = i;
    }
}
/* returns index of key; *is_new set when inserted */
static inline int64_t omap_upsert(omap *m, int64_t key, int *is_new) {
    uint64_t s = mix64((uint64_t)key) & (uint64_t)(m->cap - 1);
    for (;;) {
        int64_t idx = m->slots[s];
        if (idx < 0) break;
        if (m->keys[idx] == key) { *is_new = 0; return idx; }
        s = (s + 1) & (uint64_t)(m->cap - 1);
    }
    if (m->n == m->room) {
        m->room *= 2;
        m->keys = (int64_t *)realloc(m->keys, sizeof(int64_t) * m->room);
    }
    m->keys[m->n] = key;
    m->slots[s] = m->n;
    *is_new = 1;
    int64_t idx = m->n++;
    if (m->n * 2 > m->cap) omap_grow(m);
    return idx;
}

/* combiner ops recognised for reduceByKey(func) (dpark/rdd.py:543-545:
 * Aggregator(identity, func, func)) */
enum { ORC_SUM = 0, ORC_MIN = 1, ORC_MAX = 2, ORC_PROD = 3, ORC_AND = 4, ORC_OR = 5, ORC_XOR = 6,
       ORC_FIRST = 7, ORC_LAST = 8 };
/* value kinds: int64 (wrapping == Python big-int as long as no overflow) or
 * float64 (the reference adds Python floats, i.e. doubles) */
enum { ORC_V_I64 = 0, ORC_V_F64 = 1 };

static inline int64_t op_i64(int op, int64_t a, int64_t b) {
    switch (op) {
    case ORC_SUM: return (int64_t)((uint64_t)a + (uint64_t)b);
    case ORC_MIN: return b < a ? b : a;   /* min(x, y): y only if strictly smaller */
    case ORC_MAX: return b > a ? b : a;
    case ORC_PROD: return (int64_t)((uint64_t)a * (uint64_t)b);
    case ORC_AND: return a & b;
    case ORC_OR: return a | b;
    case ORC_XOR: return a ^ b;
    case ORC_FIRST: return a;
    case ORC_LAST: return b;
    }
    return a;
}
static inline double op_f64(int op, double a, double b) {
    switch (op) {
    case ORC_SUM: return a + b;
    case ORC_MIN: return b < a ? b : a;
    case ORC_MAX: return b > a ? b : a;
    case ORC_PROD: return a * b;
    case ORC_FIRST: return a;
    case ORC_LAST: return b;
    }
    return a;
}

/* ----------------------------------------------------- a9: reduce-side merge */
/* DiskHashMerger._merge: combined[k] = mergeCombiners(old, v) if old is not
 * None else v; iteration = first-seen order.  dpark/shuffle.py:600-608, 610-612.
 * Also the map-side bucket upsert of task.py:221-226 (same shape).
 * keys widened to int64 by the caller; vals int64 or double per vkind.
 * Returns number of distinct keys; out arrays must hold n entries. */
int64_t orc_merge(const int64_t *keys, const void *vals, int vkind, int64_t n, int op,
                  int64_t *out_keys, void *out_vals) {
    omap m;
    omap_init(&m, n < 1024 ? 1024 : n / 2);
    int64_t *ov_i = (int64_t *)out_vals;
    double *ov_f = (double *)out_vals;
    for (int64_t i = 0; i < n; i++) {
        int is_new;
        int64_t idx = omap_upsert(&m, keys[i], &is_new);
        if (vkind == ORC_V_I64) {
            int64_t v = ((const int64_t *)vals)[i];
            ov_i[idx] = is_new ? v : op_i64(op, ov_i[idx], v);
        } else {
            double v = ((const double *)vals)[i];
            ov_f[idx] = is_new ? v : op_f64(op, ov_f[idx], v);
        }
    }
    memcpy(out_keys, m.keys, sizeof(int64_t) * m.n);
    int64_t nd = m.n;
    omap_free(&m);
    return nd;
}

/* ------------------------------------------------- a4: map-side ShuffleMapTask */
/* ShuffleMapTask._run inner loop, dpark/task.py:209-226: for each (k, v):
 * bucket = buckets[getPartition(k)]; bucket[k] = mergeValue(old, v) or
 * createCombiner(v).  Output: the P buckets concatenated bucket-major, each in
 * first-seen order; bucket_offsets[P+1].  `hash` are the portable_hash values
 * of the keys (so every key kind shares this loop).  Returns total rows out. */
int64_t orc_map_task(const int64_t *keys, const int64_t *hash, const void *vals, int vkind,
                     int64_t n, int32_t P, const int64_t *thresholds, int32_t nthr, int op,
                     int combine, int64_t *out_keys, void *out_vals, int64_t *bucket_offsets) {
    int32_t *pid = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
    int64_t *cnt = (int64_t *)calloc(P + 1, sizeof(int64_t));
    for (int64_t i = 0; i < n; i++) {
        pid[i] = orc_partition(hash[i], P, thresholds, nthr);
        cnt[pid[i] + 1]++;
    }
    for (int32_t p = 0; p < P; p++) cnt[p + 1] += cnt[p];
    /* stable split into per-bucket arrival order */
    int64_t *bk = (int64_t *)malloc(sizeof(int64_t) * (n ? n : 1));
    int64_t *bv = (int64_t *)malloc(sizeof(int64_t) * (n ? n : 1)); /* 8-byte cells either kind */
    int64_t *cur = (int64_t *)malloc(sizeof(int64_t) * P);
    memcpy(cur, cnt, sizeof(int64_t) * P);
    for (int64_t i = 0; i < n; i++) {
        int64_t d = cur[pid[i]]++;
        bk[d] = keys[i];
        bv[d] = ((const int64_t *)vals)[i];
    }
    int64_t total = 0;
    bucket_offsets[0] = 0;
    for (int32_t p = 0; p < P; p++) {
        int64_t b0 = cnt[p], bn = cnt[p + 1] - cnt[p];
        if (combine) {
            int64_t nd = orc_merge(bk + b0, bv + b0, vkind, bn, op, out_keys + total,
                                   (int64_t *)out_vals + total);
            total += nd;
        } else {
            memcpy(out_keys + total, bk + b0, sizeof(int64_t) * bn);
            memcpy((int64_t *)out_vals + total, bv + b0, sizeof(int64_t) * bn);
            total += bn;
        }
        bucket_offsets[p + 1] = total;
    }
    free(pid); free(cnt); free(bk); free(bv); free(cur);
    return total;
}

/* ------------------------------------------------ a10: ordered groupByKey merge */
/* OrderedGroupByDiskHashMerger, dpark/shuffle.py:626-646 with GroupByAggregator
 * (dpark/dependency.py:107-118): values of a key ordered by (map_id, arrival
 * within the map).  Input rows must be given concatenated in map_id order, each
 * map's rows in arrival order, so a stable group-by is exactly that order.
 * Output CSR: out_keys[nd] (first-seen order), out_offsets[nd+1], out_vals[n]
 * (8-byte cells).  Returns nd. */
int64_t orc_group(const int64_t *keys, const int64_t *vals, int64_t n, int64_t *out_keys,
                  int64_t *out_offsets, int64_t *out_vals) {
    omap m;
    omap_init(&m, n < 1024 ? 1024 : n / 2);
    int64_t *idx = (int64_t *)malloc(sizeof(int64_t) * (n ? n : 1));
    int64_t *cnt = (int64_t *)calloc(n + 2, sizeof(int64_t));
    for (int64_t i = 0; i < n; i++) {
        int is_new;
        idx[i] = omap_upsert(&m, keys[i], &is_new);
        cnt[idx[i] + 1]++;
    }
    int64_t nd = m.n;
    for (int64_t g = 0; g < nd; g++) cnt[g + 1] += cnt[g];
    memcpy(out_offsets, cnt, sizeof(int64_t) * (nd + 1));
    for (int64_t i = 0; i < n; i++) out_vals[cnt[idx[i]]++] = vals[i];
    memcpy(out_keys, m.keys, sizeof(int64_t) * nd);
    omap_free(&m);
    free(idx); free(cnt);
    return nd;
}
