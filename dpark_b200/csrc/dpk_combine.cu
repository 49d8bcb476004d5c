// dpk_combine.cu -- reduce side of the shuffle: DiskHashMerger._merge
// (dpark/shuffle.py:600-608): combined[k] = mergeCombiners(combined[k], v) over
// every row fetched for the reduce partitions this GPU owns.
//
// v1 design: one open-addressing table in HBM for all local partitions (keys of
// different partitions are different keys, so they can share a table); 16-byte
// slots {key bits, accumulator} so a probe + update touches one 32 B sector.
//   k_tbl_init    : slots <- {EMPTY, identity(op)}
//   k_tbl_insert  : per row: claim slot with atomicCAS on the key word (linear
//                   probing), then one native atomic on the accumulator.
//   k_tbl_compact : per occupied slot: recompute its partition, reserve an output
//                   index inside that partition's range (CTA-aggregated), write.
// Accumulators: int64 for integer values (exact while |sum| < 2^63, as the
// reference's big ints), float64 for float values (the reference adds Python
// floats).  Algorithmic bytes: (K+V) * (rows + distinct).
#include "dpk_common.cuh"
#include <type_traits>

namespace dpk {

constexpr int CB_THREADS = 256;
constexpr int64_t kEmpty = INT64_MIN;  // slot-free marker; a real key with these bits uses the side slot

struct __align__(16) Slot {
    int64_t key;
    int64_t acc;  // int64 or the bits of a double
};

template <typename KeyT> __device__ __forceinline__ int64_t key_bits(KeyT k);
template <> __device__ __forceinline__ int64_t key_bits<int64_t>(int64_t k) { return k; }
template <> __device__ __forceinline__ int64_t key_bits<int32_t>(int32_t k) { return (int64_t)k; }
template <> __device__ __forceinline__ int64_t key_bits<uint64_t>(uint64_t k) { return (int64_t)k; }
template <> __device__ __forceinline__ int64_t key_bits<double>(double k) {
    if (k == 0.0) k = 0.0;  // -0.0 == 0.0 is one dict key in Python
    return __double_as_longlong(k);
}
template <> __device__ __forceinline__ int64_t key_bits<float>(float k) {
    if (k == 0.0f) k = 0.0f;
    return (int64_t)__float_as_int(k);
}
template <typename KeyT> __device__ __forceinline__ KeyT key_from_bits(int64_t b);
template <> __device__ __forceinline__ int64_t key_from_bits<int64_t>(int64_t b) { return b; }
template <> __device__ __forceinline__ int32_t key_from_bits<int32_t>(int64_t b) { return (int32_t)b; }
template <> __device__ __forceinline__ uint64_t key_from_bits<uint64_t>(int64_t b) { return (uint64_t)b; }
template <> __device__ __forceinline__ double key_from_bits<double>(int64_t b) { return __longlong_as_double(b); }
template <> __device__ __forceinline__ float key_from_bits<float>(int64_t b) { return __int_as_float((int)b); }

// ---- accumulator ops (op is kernel-uniform, so the switch costs nothing) ------
template <typename AccT> struct Acc;

template <> struct Acc<int64_t> {
    static int64_t identity(int op) {
        switch (op) {
        case DPK_OP_MIN: return INT64_MAX;
        case DPK_OP_MAX: return INT64_MIN;
        case DPK_OP_PROD: return 1;
        case DPK_OP_AND: return -1;
        default: return 0;
        }
    }
    static bool supports(int op) { return op >= DPK_OP_SUM && op <= DPK_OP_XOR; }
    static __device__ __forceinline__ void apply(int op, int64_t *a, int64_t v) {
        switch (op) {
        case DPK_OP_SUM: atomicAdd((unsigned long long *)a, (unsigned long long)v); break;
        case DPK_OP_MIN: atomicMin((long long *)a, (long long)v); break;
        case DPK_OP_MAX: atomicMax((long long *)a, (long long)v); break;
        case DPK_OP_AND: atomicAnd((unsigned long long *)a, (unsigned long long)v); break;
        case DPK_OP_OR: atomicOr((unsigned long long *)a, (unsigned long long)v); break;
        case DPK_OP_XOR: atomicXor((unsigned long long *)a, (unsigned long long)v); break;
        default: {  // PROD (wrapping, like int64 multiply)
            unsigned long long old = *(volatile unsigned long long *)a, assumed;
            do {
                assumed = old;
                old = atomicCAS((unsigned long long *)a, assumed, assumed * (unsigned long long)v);
            } while (old != assumed);
        }
        }
    }
};

template <> struct Acc<double> {
    static int64_t identity(int op) {
        double d = 0.0;
        switch (op) {
        case DPK_OP_MIN: d = INFINITY; break;
        case DPK_OP_MAX: d = -INFINITY; break;
        case DPK_OP_PROD: d = 1.0; break;
        default: d = 0.0;
        }
        int64_t b;
        memcpy(&b, &d, 8);
        return b;
    }
    static bool supports(int op) { return op >= DPK_OP_SUM && op <= DPK_OP_PROD; }
    static __device__ __forceinline__ void apply(int op, int64_t *a, double v) {
        if (op == DPK_OP_SUM) {
            atomicAdd((double *)a, v);
            return;
        }
        unsigned long long old = *(volatile unsigned long long *)a, assumed;
        do {
            assumed = old;
            double cur = __longlong_as_double((long long)assumed), nv;
            if (op == DPK_OP_MIN) nv = v < cur ? v : cur;
            else if (op == DPK_OP_MAX) nv = v > cur ? v : cur;
            else nv = cur * v;
            if (__double_as_longlong(nv) == (long long)assumed) break;
            old = atomicCAS((unsigned long long *)a, assumed, (unsigned long long)__double_as_longlong(nv));
        } while (old != assumed);
    }
};

// ---- kernels ---------------------------------------------------------------
__global__ void __launch_bounds__(CB_THREADS)
k_tbl_init(Slot *__restrict__ table, int64_t slots, int64_t ident) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int4 fill = make_int4((int)(uint32_t)((uint64_t)kEmpty & 0xffffffffu), (int)(uint32_t)((uint64_t)kEmpty >> 32),
                                (int)(uint32_t)((uint64_t)ident & 0xffffffffu), (int)(uint32_t)((uint64_t)ident >> 32));
    for (; i < slots; i += stride) reinterpret_cast<int4 *>(table)[i] = fill;
}

template <typename KeyT, typename ValT, typename AccT>
__global__ void __launch_bounds__(CB_THREADS)
k_tbl_insert(const KeyT *__restrict__ keys, const ValT *__restrict__ vals, int64_t n, int op,
             Slot *__restrict__ table, uint64_t mask, Slot *__restrict__ side, int32_t *__restrict__ side_used) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int64_t kb = key_bits<KeyT>(keys[i]);
        const AccT v = (AccT)vals[i];
        Slot *s;
        if (kb == kEmpty) {
            s = side;
            *side_used = 1;
        } else {
            uint64_t h = mix64((uint64_t)kb) & mask;
            for (;;) {
                int64_t cur = __ldcg(&table[h].key);
                if (cur == kb) break;
                if (cur == kEmpty) {
                    unsigned long long prev = atomicCAS((unsigned long long *)&table[h].key,
                                                        (unsigned long long)kEmpty, (unsigned long long)kb);
                    if (prev == (unsigned long long)kEmpty || prev == (unsigned long long)kb) break;
                }
                h = (h + 1) & mask;
            }
            s = &table[h];
        }
        Acc<AccT>::apply(op, &s->acc, v);
    }
}

// slots [0, nslots) are the table; slot nslots is the side slot (valid iff *side_used)
template <typename KeyT>
__global__ void __launch_bounds__(CB_THREADS)
k_tbl_compact(const Slot *__restrict__ table, int64_t nslots, const int32_t *__restrict__ side_used,
              PartFn f, int32_t part_first, int32_t nparts, const int64_t *__restrict__ part_offsets,
              KeyT *__restrict__ out_keys, int64_t *__restrict__ out_vals,
              unsigned long long *__restrict__ out_counts) {
    extern __shared__ __align__(16) int32_t s_mem[];  // [nparts] counts, [nparts] 64-bit bases after
    int32_t *s_cnt = s_mem;
    long long *s_base = reinterpret_cast<long long *>(s_mem + ((nparts + 1) & ~1));
    const int lane = threadIdx.x & 31;
    constexpr int ITEMS = 4;
    const int64_t total = nslots + 1;
    const int64_t tile = (int64_t)CB_THREADS * ITEMS;
    for (int64_t t0 = (int64_t)blockIdx.x * tile; t0 < total; t0 += (int64_t)gridDim.x * tile) {
        for (int p = threadIdx.x; p < nparts; p += CB_THREADS) s_cnt[p] = 0;
        __syncthreads();
        int64_t kb[ITEMS], acc[ITEMS];
        int lp[ITEMS], rk[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            int64_t i = t0 + (int64_t)j * CB_THREADS + threadIdx.x;
            lp[j] = -1;
            if (i < total) {
                int4 raw = __ldcs(reinterpret_cast<const int4 *>(table) + i);
                kb[j] = (int64_t)(((uint64_t)(uint32_t)raw.y << 32) | (uint32_t)raw.x);
                acc[j] = (int64_t)(((uint64_t)(uint32_t)raw.w << 32) | (uint32_t)raw.z);
                bool occ = (i < nslots) ? (kb[j] != kEmpty) : (*side_used != 0);
                if (i == nslots) kb[j] = kEmpty;
                if (occ) lp[j] = f(KeyHash<KeyT>::of(key_from_bits<KeyT>(kb[j]))) - part_first;
            }
        }
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            unsigned m = __match_any_sync(0xffffffffu, lp[j]);
            int base = 0;
            const bool ok = lp[j] >= 0 && lp[j] < nparts;
            if (ok && lane == __ffs(m) - 1) base = atomicAdd(&s_cnt[lp[j]], __popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            rk[j] = base + __popc(m & ((1u << lane) - 1u));
        }
        __syncthreads();
        for (int p = threadIdx.x; p < nparts; p += CB_THREADS) {
            int c = s_cnt[p];
            s_base[p] = c ? (long long)atomicAdd(&out_counts[p], (unsigned long long)c) : 0;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            if (lp[j] >= 0 && lp[j] < nparts) {
                int64_t dst = part_offsets[lp[j]] + s_base[lp[j]] + rk[j];
                out_keys[dst] = key_from_bits<KeyT>(kb[j]);
                out_vals[dst] = acc[j];
            }
        }
        __syncthreads();
    }
}

static inline int64_t pow2ceil(int64_t x) {
    int64_t p = 1024;
    while (p < x) p <<= 1;
    return p;
}
static inline int64_t table_slots_for(int64_t n) { return pow2ceil(2 * (n > 0 ? n : 1)); }

static inline int grid_cap(int64_t items, int per_cta, int waves) {
    int64_t g = (items + per_cta - 1) / per_cta;
    int64_t cap = (int64_t)sm_count() * waves;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

template <typename KeyT, typename ValT, typename AccT>
static int dispatch_op(int op, const void *keys, const void *vals, int64_t n, Slot *table, int64_t slots,
                       Slot *side, int32_t *side_used, cudaStream_t st) {
    if (!Acc<AccT>::supports(op)) return fail(DPK_ERR_UNSUPPORTED, "op %d unsupported for this value kind", op);
    DPK_LAUNCH("tbl_init", st, k_tbl_init<<<grid_cap(slots + 1, CB_THREADS, 16), CB_THREADS, 0, st>>>(table, slots + 1, Acc<AccT>::identity(op)));
    if (n > 0) {
        DPK_LAUNCH("tbl_insert", st, k_tbl_insert<KeyT, ValT, AccT><<<grid_cap(n, CB_THREADS, 16), CB_THREADS, 0, st>>>(
            (const KeyT *)keys, (const ValT *)vals, n, op, table, (uint64_t)(slots - 1), side, side_used));
    }
    return DPK_OK;
}

template <typename KeyT>
static int dispatch_valkind(int val_kind, int op, const void *keys, const void *vals, int64_t n, Slot *table,
                            int64_t slots, Slot *side, int32_t *side_used, cudaStream_t st) {
    switch (val_kind) {
    case DPK_V_I64: return dispatch_op<KeyT, int64_t, int64_t>(op, keys, vals, n, table, slots, side, side_used, st);
    case DPK_V_I32: return dispatch_op<KeyT, int32_t, int64_t>(op, keys, vals, n, table, slots, side, side_used, st);
    case DPK_V_F64: return dispatch_op<KeyT, double, double>(op, keys, vals, n, table, slots, side, side_used, st);
    case DPK_V_F32: return dispatch_op<KeyT, float, double>(op, keys, vals, n, table, slots, side, side_used, st);
    }
    return fail(DPK_ERR_UNSUPPORTED, "value kind %d unsupported", val_kind);
}

template <typename KeyT>
static int run_combine(const void *keys, const void *vals, int val_kind, int64_t n, int op, const PartFn &f,
                       int32_t part_first, int32_t nparts, const int64_t *part_offsets, void *out_keys,
                       void *out_vals, int64_t *out_counts, void *ws, cudaStream_t st) {
    const int64_t slots = table_slots_for(n);
    Slot *table = (Slot *)ws;
    Slot *side = table + slots;
    int32_t *side_used = (int32_t *)(side + 1);
    DPK_CUDA_TRY(cudaMemsetAsync(side_used, 0, 16, st));
    DPK_CUDA_TRY(cudaMemsetAsync(out_counts, 0, (size_t)nparts * 8, st));
    int rc = dispatch_valkind<KeyT>(val_kind, op, keys, vals, n, table, slots, side, side_used, st);
    if (rc) return rc;
    size_t sh = (size_t)((nparts + 1) & ~1) * 4 + (size_t)nparts * 8;
    DPK_LAUNCH("tbl_compact", st, k_tbl_compact<KeyT><<<grid_cap(slots + 1, CB_THREADS * 4, 8), CB_THREADS, sh, st>>>(
        table, slots, side_used, f, part_first, nparts, part_offsets, (KeyT *)out_keys, (int64_t *)out_vals,
        (unsigned long long *)out_counts));
    return DPK_OK;
}

}  // namespace dpk

using namespace dpk;

extern "C" {

int64_t dpk_combine_workspace_bytes(int64_t n) {
    if (n < 0) n = 0;
    return (table_slots_for(n) + 2) * (int64_t)sizeof(Slot);
}

int dpk_combine(const void *keys, int key_kind, const void *vals, int val_kind, int64_t n, int op, int32_t P,
                const int64_t *thresholds, int32_t nthr, int32_t part_first, int32_t nparts,
                const int64_t *part_offsets, void *out_keys, void *out_vals, int64_t *out_counts, void *ws,
                int64_t ws_bytes, dpk_stream_t stream) {
    if (n < 0 || n >= ((int64_t)1 << 31)) return fail(DPK_ERR_INVALID, "n=%lld out of range [0, 2^31)", (long long)n);
    if (nparts < 1 || part_first < 0 || part_first + nparts > P)
        return fail(DPK_ERR_INVALID, "bad partition range first=%d n=%d P=%d", part_first, nparts, P);
    if (nparts > 4096) return fail(DPK_ERR_UNSUPPORTED, "nparts=%d > 4096", nparts);
    if (!part_offsets || !out_counts || !ws) return fail(DPK_ERR_INVALID, "NULL pointer");
    if (n > 0 && (!keys || !vals || !out_keys || !out_vals)) return fail(DPK_ERR_INVALID, "NULL pointer");
    if (ws_bytes < dpk_combine_workspace_bytes(n))
        return fail(DPK_ERR_WORKSPACE, "workspace needs %lld B, got %lld", (long long)dpk_combine_workspace_bytes(n), (long long)ws_bytes);
    PartFn f;
    int rc = make_partfn(P, thresholds, nthr, &f);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    switch (key_kind) {
    case DPK_K_I64: return run_combine<int64_t>(keys, vals, val_kind, n, op, f, part_first, nparts, part_offsets, out_keys, out_vals, out_counts, ws, st);
    case DPK_K_I32: return run_combine<int32_t>(keys, vals, val_kind, n, op, f, part_first, nparts, part_offsets, out_keys, out_vals, out_counts, ws, st);
    case DPK_K_F64: return run_combine<double>(keys, vals, val_kind, n, op, f, part_first, nparts, part_offsets, out_keys, out_vals, out_counts, ws, st);
    case DPK_K_U64: return run_combine<uint64_t>(keys, vals, val_kind, n, op, f, part_first, nparts, part_offsets, out_keys, out_vals, out_counts, ws, st);
    case DPK_K_F32: return run_combine<float>(keys, vals, val_kind, n, op, f, part_first, nparts, part_offsets, out_keys, out_vals, out_counts, ws, st);
    }
    return fail(DPK_ERR_UNSUPPORTED, "key kind %d is unhashable by portable_hash", key_kind);
}

}  // extern "C"
