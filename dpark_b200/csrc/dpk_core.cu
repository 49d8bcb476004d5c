// dpk_core.cu -- error plumbing, device info, and the stand-alone a1/a2 entry
// points (vector portable_hash, partition ids).  The fused hot kernels live in
// dpk_partition.cu (map side) and dpk_combine.cu (reduce side).
#include "dpk_common.cuh"

namespace dpk {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;
    }
    return cached;
}

// ---- profiling -------------------------------------------------------------
static long long g_launches = 0;
static bool g_prof_on = false;
static int g_prof_n = 0;
static cudaEvent_t g_ev0[DPK_PROF_MAX], g_ev1[DPK_PROF_MAX];
static bool g_ev_made[DPK_PROF_MAX];
static const char *g_prof_label[DPK_PROF_MAX];

ProfScope::ProfScope(const char *label, cudaStream_t s) : idx(-1), st(s) {
    g_launches++;
    if (g_prof_on && g_prof_n < DPK_PROF_MAX) {
        idx = g_prof_n++;
        if (!g_ev_made[idx]) {
            cudaEventCreate(&g_ev0[idx]);
            cudaEventCreate(&g_ev1[idx]);
            g_ev_made[idx] = true;
        }
        g_prof_label[idx] = label;
        cudaEventRecord(g_ev0[idx], st);
    }
}
ProfScope::~ProfScope() {
    if (idx >= 0) cudaEventRecord(g_ev1[idx], st);
}

int make_partfn(int32_t P, const int64_t *thresholds, int32_t nthr, int32_t sub_bits, PartFn *out) {
    if (P < 1) return fail(DPK_ERR_INVALID, "P must be >= 1, got %d", P);
    // the bucket limit belongs to the multisplit kernels (their shared-memory
    // histograms); plain getPartition (sub_bits == 0) works for any P
    if (sub_bits < 0 || sub_bits > 12 || (sub_bits > 0 && ((int64_t)P << sub_bits) > DPK_MAX_PARTITIONS))
        return fail(DPK_ERR_UNSUPPORTED, "P=%d with sub_bits=%d exceeds %d buckets", P, sub_bits, DPK_MAX_PARTITIONS);
    PartFn f;
    f.P = P; f.magic = 0; f.shift = 0; f.nthr = 0; f.thresholds = nullptr; f.sub_bits = sub_bits;
    f.row_hash = nullptr;
    if (thresholds != nullptr) {
        if (nthr != P - 1) return fail(DPK_ERR_INVALID, "thresholds need P-1=%d entries, got %d", P - 1, nthr);
        f.mode = 3; f.nthr = nthr; f.thresholds = thresholds;
    } else if (P == 1) {
        f.mode = 0;
    } else if ((P & (P - 1)) == 0) {
        f.mode = 1;
    } else {
        // round-up magic for floor(a / d), a < 2^64:  q = mulhi(m, a); t = ((a-q)>>1)+q; t >> s
        uint64_t d = (uint64_t)P;
        int s = 63 - __builtin_clzll(d);
        unsigned __int128 num = (unsigned __int128)1 << (64 + s);
        uint64_t m = (uint64_t)(num / d), rem = (uint64_t)(num % d);
        m += m;
        uint64_t twice = rem + rem;
        if (twice >= d || twice < rem) m += 1;
        f.mode = 2; f.magic = m + 1; f.shift = s;
    }
    *out = f;
    return DPK_OK;
}

template <typename KeyT>
__global__ void k_hash_keys(const KeyT *__restrict__ keys, int64_t n, int64_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = KeyHash<KeyT>::of(keys[i]);
}

__global__ void k_hash_bytes(const uint8_t *__restrict__ data, const int64_t *__restrict__ offsets,
                             int64_t n, int mode, int64_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        int64_t b = offsets[i], e = offsets[i + 1];
        out[i] = mode == DPK_BYTES_SIGNED ? hash_bytes_signed(data + b, e - b)
                                          : hash_utf8_codepoints(data + b, e - b);
    }
}

__global__ void k_hash_tuple(const int64_t *__restrict__ item_hash, int64_t n, int32_t arity, int64_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = hash_tuple_items(item_hash, n, i, arity);
}

__global__ void k_partition_ids(const int64_t *__restrict__ hash, int64_t n, PartFn f,
                                int32_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = f(hash[i]);
}

static inline int grid_for(int64_t n, int threads) {
    int64_t g = (n + threads - 1) / threads;
    int64_t cap = (int64_t)sm_count() * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace dpk

using namespace dpk;

extern "C" {

int dpk_abi_version(void) { return DPK_ABI_VERSION; }
const char *dpk_last_error(void) { return g_err; }

int dpk_device_info(int32_t *h_info) {
    if (!h_info) return fail(DPK_ERR_INVALID, "h_info is NULL");
    int dev = 0;
    DPK_CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp p;
    DPK_CUDA_TRY(cudaGetDeviceProperties(&p, dev));
    h_info[0] = p.multiProcessorCount;
    h_info[1] = p.major;
    h_info[2] = p.minor;
    h_info[3] = (int32_t)(p.l2CacheSize >> 20);
    return DPK_OK;
}

int dpk_hash_keys(const void *keys, int key_kind, int64_t n, int64_t *out_hash, dpk_stream_t stream) {
    if (n < 0) return fail(DPK_ERR_INVALID, "n < 0");
    if (n == 0) return DPK_OK;
    if (!keys || !out_hash) return fail(DPK_ERR_INVALID, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int g = grid_for(n, 256);
    switch (key_kind) {
    case DPK_K_I64: DPK_LAUNCH("hash_keys", st, k_hash_keys<int64_t><<<g, 256, 0, st>>>((const int64_t *)keys, n, out_hash)); break;
    case DPK_K_I32: DPK_LAUNCH("hash_keys", st, k_hash_keys<int32_t><<<g, 256, 0, st>>>((const int32_t *)keys, n, out_hash)); break;
    case DPK_K_F64: DPK_LAUNCH("hash_keys", st, k_hash_keys<double><<<g, 256, 0, st>>>((const double *)keys, n, out_hash)); break;
    case DPK_K_U64: DPK_LAUNCH("hash_keys", st, k_hash_keys<uint64_t><<<g, 256, 0, st>>>((const uint64_t *)keys, n, out_hash)); break;
    case DPK_K_F32: DPK_LAUNCH("hash_keys", st, k_hash_keys<float><<<g, 256, 0, st>>>((const float *)keys, n, out_hash)); break;
    default: return fail(DPK_ERR_UNSUPPORTED, "key kind %d is unhashable by portable_hash", key_kind);
    }
    return DPK_OK;
}

int dpk_hash_bytes(const uint8_t *data, const int64_t *offsets, int64_t n, int mode,
                   int64_t *out_hash, dpk_stream_t stream) {
    if (n < 0) return fail(DPK_ERR_INVALID, "n < 0");
    if (n == 0) return DPK_OK;
    if (!offsets || !out_hash) return fail(DPK_ERR_INVALID, "NULL pointer");
    if (mode != DPK_BYTES_SIGNED && mode != DPK_STR_UTF8) return fail(DPK_ERR_UNSUPPORTED, "bad bytes mode %d", mode);
    cudaStream_t st = (cudaStream_t)stream;
    DPK_LAUNCH("hash_bytes", st, k_hash_bytes<<<grid_for(n, 128), 128, 0, st>>>(data, offsets, n, mode, out_hash));
    return DPK_OK;
}

int dpk_hash_tuple(const int64_t *item_hash, int64_t n, int32_t arity, int64_t *out_hash, dpk_stream_t stream) {
    if (n < 0 || arity < 0 || arity > 64) return fail(DPK_ERR_INVALID, "bad n=%lld or arity=%d", (long long)n, arity);
    if (n == 0) return DPK_OK;
    if ((arity > 0 && !item_hash) || !out_hash) return fail(DPK_ERR_INVALID, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    DPK_LAUNCH("hash_tuple", st, k_hash_tuple<<<grid_for(n, 256), 256, 0, st>>>(item_hash, n, arity, out_hash));
    return DPK_OK;
}

int dpk_partition_ids(const int64_t *hash, int64_t n, int32_t P, const int64_t *thresholds,
                      int32_t nthr, int32_t *out_pid, dpk_stream_t stream) {
    if (n < 0) return fail(DPK_ERR_INVALID, "n < 0");
    PartFn f;
    int rc = make_partfn(P, thresholds, nthr, 0, &f);
    if (rc) return rc;
    if (n == 0) return DPK_OK;
    if (!hash || !out_pid) return fail(DPK_ERR_INVALID, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    DPK_LAUNCH("partition_ids", st, k_partition_ids<<<grid_for(n, 256), 256, 0, st>>>(hash, n, f, out_pid));
    return DPK_OK;
}

int64_t dpk_launch_count(void) { return g_launches; }
int dpk_prof_enable(int on) {
    g_prof_on = on != 0;
    if (on) g_prof_n = 0;
    return DPK_OK;
}
int dpk_prof_count(void) { return g_prof_n; }
int dpk_prof_get(int i, char *h_name, float *h_ms) {
    if (i < 0 || i >= g_prof_n || !h_name || !h_ms) return fail(DPK_ERR_INVALID, "bad profile index %d", i);
    DPK_CUDA_TRY(cudaEventSynchronize(g_ev1[i]));
    DPK_CUDA_TRY(cudaEventElapsedTime(h_ms, g_ev0[i], g_ev1[i]));
    snprintf(h_name, 64, "%s", g_prof_label[i]);
    return DPK_OK;
}

}  // extern "C"
