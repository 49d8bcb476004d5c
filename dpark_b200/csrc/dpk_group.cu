// dpk_group.cu -- reduce side of groupByKey: GroupByAggregator merged by
// OrderedGroupByDiskHashMerger (dpark/dependency.py:107-118, dpark/shuffle.py:626-646):
// for every key the list of its values, ordered by (map_id, arrival within the map).
//
// On the device a group-by is a STABLE sort of the received rows by key: the rows
// arrive ordered by (source rank = map block, position), a stable sort keeps that
// order inside every group.  The sort is LSD radix over the key bits, each pass the
// stable multisplit of dpk_partition.cu (dpk_radix_pass); digits in which all keys
// agree are skipped (k_key_or finds them).  After the passes, one more stable
// multisplit by reduce partition (dpk_partition with sub_bits = 0) makes the
// buffer partition-major.  The kernels here turn the sorted buffer into CSR:
//   k_heads_count / k_heads_scan / k_heads_write : out_keys[g], out_starts[g]
//   (row index of the first value of group g), out_starts[G] = n.
// Algorithmic bytes: (K+V)*N read + V*N written in group order + (K+8)*distinct.
#include "dpk_common.cuh"

namespace dpk {

constexpr int GR_THREADS = 256;
constexpr int GR_ITEMS = 16;
constexpr int GR_TILE = GR_THREADS * GR_ITEMS;

// OR over i of (keys[i] ^ keys[0]): the bit positions in which keys differ
__global__ void __launch_bounds__(GR_THREADS)
k_key_or(const int64_t *__restrict__ keys, int64_t n, unsigned long long *__restrict__ out) {
    const unsigned long long first = (unsigned long long)keys[0];
    unsigned long long acc = 0;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) acc |= (unsigned long long)keys[i] ^ first;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) acc |= __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0 && acc) atomicOr(out, acc);
}

__device__ __forceinline__ bool is_head(const int64_t *__restrict__ keys, int64_t i) {
    return i == 0 || keys[i] != keys[i - 1];
}

__global__ void __launch_bounds__(GR_THREADS)
k_heads_count(const int64_t *__restrict__ keys, int64_t n, int32_t *__restrict__ tile_counts) {
    __shared__ int s_w[GR_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * GR_TILE;
    int c = 0;
#pragma unroll
    for (int j = 0; j < GR_ITEMS; j++) {
        const int64_t i = base + (int64_t)j * GR_THREADS + threadIdx.x;
        if (i < n) c += is_head(keys, i);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < GR_THREADS / 32; w++) t += s_w[w];
        tile_counts[blockIdx.x] = t;
    }
}

// single CTA: exclusive scan of the per-tile head counts (int64 bases), total -> *ngroups
__global__ void __launch_bounds__(GR_THREADS)
k_heads_scan(const int32_t *__restrict__ tile_counts, int64_t T, int64_t *__restrict__ tile_base,
             int64_t *__restrict__ ngroups, int64_t *__restrict__ out_starts, int64_t n) {
    __shared__ long long s_part[GR_THREADS];
    const int64_t E = (T + GR_THREADS - 1) / GR_THREADS;
    const int64_t b0 = threadIdx.x * E, b1 = min(b0 + E, T);
    long long sum = 0;
    for (int64_t i = b0; i < b1; i++) sum += tile_counts[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    long long base = 0, tot = 0;
    for (int t = 0; t < GR_THREADS; t++) {
        if (t < (int)threadIdx.x) base += s_part[t];
        tot += s_part[t];
    }
    for (int64_t i = b0; i < b1; i++) {
        tile_base[i] = base;
        base += tile_counts[i];
    }
    if (threadIdx.x == 0) {
        *ngroups = tot;
        out_starts[tot] = n;
    }
}

__global__ void __launch_bounds__(GR_THREADS)
k_heads_write(const int64_t *__restrict__ keys, int64_t n, const int64_t *__restrict__ tile_base,
              int64_t *__restrict__ out_keys, int64_t *__restrict__ out_starts) {
    __shared__ int s_w[GR_THREADS / 32];
    // thread t owns the GR_ITEMS consecutive rows [base + t*GR_ITEMS, ...): heads stay in row order
    const int64_t base = (int64_t)blockIdx.x * GR_TILE + (int64_t)threadIdx.x * GR_ITEMS;
    unsigned flags = 0;
    int c = 0;
#pragma unroll
    for (int j = 0; j < GR_ITEMS; j++) {
        const int64_t i = base + j;
        if (i < n && is_head(keys, i)) { flags |= 1u << j; c++; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < warp; w++) wbase += s_w[w];
    int64_t g = tile_base[blockIdx.x] + wbase + inc - c;
#pragma unroll
    for (int j = 0; j < GR_ITEMS; j++) {
        if (flags & (1u << j)) {
            out_keys[g] = keys[base + j];
            out_starts[g] = base + j;
            g++;
        }
    }
}

__global__ void __launch_bounds__(GR_THREADS)
k_gather_i64(const int64_t *__restrict__ src, const int64_t *__restrict__ idx, int64_t n,
             int64_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = src[idx[i]];
}

}  // namespace dpk

using namespace dpk;

extern "C" {

int dpk_gather_i64(const int64_t *src, const int64_t *idx, int64_t n, int64_t *out, dpk_stream_t stream) {
    if (n < 0) return fail(DPK_ERR_INVALID, "n < 0");
    if (n == 0) return DPK_OK;
    if (!src || !idx || !out) return fail(DPK_ERR_INVALID, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t g = (n + GR_THREADS - 1) / GR_THREADS, cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    DPK_LAUNCH("gather_i64", st, k_gather_i64<<<(int)g, GR_THREADS, 0, st>>>(src, idx, n, out));
    return DPK_OK;
}

int dpk_key_or(const int64_t *keys, int64_t n, uint64_t *out_or, dpk_stream_t stream) {
    if (n < 0) return fail(DPK_ERR_INVALID, "n < 0");
    if (!out_or) return fail(DPK_ERR_INVALID, "out_or is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    DPK_CUDA_TRY(cudaMemsetAsync(out_or, 0, 8, st));
    if (n == 0) return DPK_OK;
    if (!keys) return fail(DPK_ERR_INVALID, "keys is NULL");
    int64_t g = (n + GR_TILE - 1) / GR_TILE, cap = (int64_t)sm_count() * 8;
    if (g > cap) g = cap;
    DPK_LAUNCH("key_or", st, k_key_or<<<(int)g, GR_THREADS, 0, st>>>(keys, n, (unsigned long long *)out_or));
    return DPK_OK;
}

int64_t dpk_group_heads_workspace_bytes(int64_t n) {
    int64_t T = (n + GR_TILE - 1) / GR_TILE + 1;
    return T * 4 + T * 8 + 256;
}

int dpk_group_heads(const int64_t *sorted_keys, int64_t n, int64_t *out_keys, int64_t *out_starts,
                    int64_t *out_ngroups, void *ws, int64_t ws_bytes, dpk_stream_t stream) {
    if (n < 0 || n >= ((int64_t)1 << 31)) return fail(DPK_ERR_INVALID, "n=%lld out of range [0, 2^31)", (long long)n);
    if (!out_starts || !out_ngroups || !ws) return fail(DPK_ERR_INVALID, "NULL pointer");
    if (ws_bytes < dpk_group_heads_workspace_bytes(n)) return fail(DPK_ERR_WORKSPACE, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t T = (n + GR_TILE - 1) / GR_TILE;
    int64_t *tile_base = (int64_t *)ws;
    int32_t *tile_counts = (int32_t *)(tile_base + T + 1);
    if (n > 0) {
        if (!sorted_keys || !out_keys) return fail(DPK_ERR_INVALID, "NULL pointer");
        DPK_LAUNCH("heads_count", st, k_heads_count<<<(int)T, GR_THREADS, 0, st>>>(sorted_keys, n, tile_counts));
    }
    DPK_LAUNCH("heads_scan", st, k_heads_scan<<<1, GR_THREADS, 0, st>>>(tile_counts, T, tile_base, out_ngroups, out_starts, n));
    if (n > 0)
        DPK_LAUNCH("heads_write", st, k_heads_write<<<(int)T, GR_THREADS, 0, st>>>(sorted_keys, n, tile_base, out_keys, out_starts));
    return DPK_OK;
}

}  // extern "C"
