// dpk_strings.cu -- key identity for variable-length keys (str / bytes).
// The reference's per-bucket dicts (dpark/task.py:221-226) and merge dict
// (dpark/shuffle.py:600-608) compare keys by VALUE.  On the device a row's key is
// (data, offsets[i]..offsets[i+1]); k_dict_encode maps every row to the index of a
// representative row with the same bytes, so the fixed-width machinery (partition,
// combine) can run on int64 ids without ever trusting the 64-bit hash as identity.
#include "dpk_common.cuh"

namespace dpk {

__global__ void __launch_bounds__(256)
k_dict_encode(const uint8_t *__restrict__ data, const int64_t *__restrict__ offsets,
              const int64_t *__restrict__ hash, int64_t n, long long *__restrict__ table, uint64_t mask,
              int64_t *__restrict__ rep) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int64_t h = hash[i];
        const int64_t b0 = offsets[i];
        const int64_t len = offsets[i + 1] - b0;
        uint64_t slot = mix64((uint64_t)h) & mask;
        int64_t r = i;
        for (;;) {
            long long cur = __ldcg(&table[slot]);
            if (cur < 0) {
                long long prev = (long long)atomicCAS((unsigned long long *)&table[slot], (unsigned long long)-1ll,
                                                      (unsigned long long)i);
                if (prev == -1ll) break;  // this row is the representative
                cur = prev;
            }
            if (hash[cur] == h) {
                const int64_t b1 = offsets[cur];
                if (offsets[cur + 1] - b1 == len) {
                    int64_t j = 0;
                    while (j < len && data[b0 + j] == data[b1 + j]) j++;
                    if (j == len) { r = cur; break; }
                }
            }
            slot = (slot + 1) & mask;
        }
        rep[i] = r;
    }
}

// ---- device text ingest -------------------------------------------------------------------------------------------
// TextFileRDD (dpark/rdd.py:1633-1711) hands out lines, and examples/wc.py splits each with `x.strip().split()`: for
// ASCII text the tokens of a byte range that starts and ends on line boundaries are exactly its maximal runs of
// non-whitespace bytes (newlines are whitespace; str.split() without arguments splits on ' ', \t \n \v \f \r and
// \x1c..\x1f).  Two passes over the bytes: count the token starts per 4096-byte block, then -- after the host-side scan
// of the block counts -- write (start, length) of every token in text order.  A byte >= 0x80 raises `flags` bit 0: the
// caller then leaves the split to the row-wise path (Unicode whitespace and decoding errors are Python's business).
constexpr int TK_THREADS = 256;   // TK_BYTES (16 bytes per thread): dpk_common.cuh
constexpr int TK_CHUNK = TK_THREADS * TK_BYTES;

// tok_ws / tok_starts16 (the per-thread token-start mask) live in dpk_common.cuh: __host__ __device__, so that
// tests/hostcheck.cu can run the same arithmetic on the CPU against Python's str.split()

__global__ void __launch_bounds__(TK_THREADS)
k_tok_count(const uint8_t *__restrict__ data, int64_t n, int64_t *__restrict__ block_counts, unsigned long long *__restrict__ flags) {
    __shared__ int s_w[TK_THREADS / 32];
    const int64_t i0 = ((int64_t)blockIdx.x * TK_THREADS + threadIdx.x) * TK_BYTES;
    bool hi = false;
    const uint32_t m = i0 < n ? tok_starts16(data, n, i0, &hi) : 0u;
    int c = __popc(m);
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    if (__any_sync(0xffffffffu, hi) && (threadIdx.x & 31) == 0) atomicOr(flags, 1ull);
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < TK_THREADS / 32; w++) t += s_w[w];
        block_counts[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(TK_THREADS)
k_tok_emit(const uint8_t *__restrict__ data, int64_t n, const int64_t *__restrict__ block_base, int64_t *__restrict__ starts,
           int64_t *__restrict__ lens) {
    __shared__ int s_w[TK_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i0 = ((int64_t)blockIdx.x * TK_THREADS + threadIdx.x) * TK_BYTES;
    bool hi = false;
    uint32_t m = i0 < n ? tok_starts16(data, n, i0, &hi) : 0u;
    const int c = __popc(m);
    int incl = c;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; w++) before += s_w[w];
    int64_t r = block_base[blockIdx.x] + before + incl - c;
    while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        const int64_t b = i0 + j;
        int64_t e = b + 1;
        while (e < n && !tok_ws(data[e])) e++;
        starts[r] = b;
        lens[r] = e - b;
        r++;
    }
}

// out[out_off[i] .. out_off[i] + lens[row]) = data[starts[row] ..), row = idx ? idx[i] : i  (token bytes made contiguous:
// the (data, offsets) form dpk_hash_bytes / dpk_dict_encode take; or the bytes of the distinct keys for the host)
__global__ void __launch_bounds__(256)
k_gather_bytes(const uint8_t *__restrict__ data, const int64_t *__restrict__ starts, const int64_t *__restrict__ lens,
               const int64_t *__restrict__ idx, int64_t m, const int64_t *__restrict__ out_off, uint8_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < m; i += stride) {
        const int64_t row = idx ? idx[i] : i;
        const uint8_t *src = data + starts[row];
        uint8_t *dst = out + out_off[i];
        const int64_t len = lens[row];
        for (int64_t j = 0; j < len; j++) dst[j] = src[j];
    }
}

static inline int64_t dict_slots(int64_t n) {
    int64_t s = 1024;
    while (s < 2 * n) s <<= 1;
    return s;
}

}  // namespace dpk

using namespace dpk;

extern "C" {

int64_t dpk_tokenize_blocks(int64_t n) { return n <= 0 ? 0 : (n + TK_CHUNK - 1) / TK_CHUNK; }

int dpk_tokenize_count(const uint8_t *data, int64_t n, int64_t *block_counts, int64_t *flags, dpk_stream_t stream) {
    if (n < 0) return fail(DPK_ERR_INVALID, "n=%lld < 0", (long long)n);
    if (n == 0) return DPK_OK;
    if (!data || !block_counts || !flags) return fail(DPK_ERR_INVALID, "NULL pointer");
    const int64_t nb = dpk_tokenize_blocks(n);
    if (nb >= ((int64_t)1 << 31)) return fail(DPK_ERR_INVALID, "text of %lld bytes is too long for one launch", (long long)n);
    cudaStream_t st = (cudaStream_t)stream;
    DPK_LAUNCH("tok_count", st, k_tok_count<<<(int)nb, TK_THREADS, 0, st>>>(data, n, block_counts, (unsigned long long *)flags));
    return DPK_OK;
}

int dpk_tokenize_emit(const uint8_t *data, int64_t n, const int64_t *block_base, int64_t *starts, int64_t *lens,
                      dpk_stream_t stream) {
    if (n < 0) return fail(DPK_ERR_INVALID, "n=%lld < 0", (long long)n);
    if (n == 0) return DPK_OK;
    if (!data || !block_base || !starts || !lens) return fail(DPK_ERR_INVALID, "NULL pointer");
    const int64_t nb = dpk_tokenize_blocks(n);
    if (nb >= ((int64_t)1 << 31)) return fail(DPK_ERR_INVALID, "text of %lld bytes is too long for one launch", (long long)n);
    cudaStream_t st = (cudaStream_t)stream;
    DPK_LAUNCH("tok_emit", st, k_tok_emit<<<(int)nb, TK_THREADS, 0, st>>>(data, n, block_base, starts, lens));
    return DPK_OK;
}

int dpk_gather_bytes(const uint8_t *data, const int64_t *starts, const int64_t *lens, const int64_t *idx, int64_t m,
                     const int64_t *out_off, uint8_t *out, dpk_stream_t stream) {
    if (m < 0) return fail(DPK_ERR_INVALID, "m=%lld < 0", (long long)m);
    if (m == 0) return DPK_OK;
    if (!data || !starts || !lens || !out_off || !out) return fail(DPK_ERR_INVALID, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t g = (m + 255) / 256, cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    DPK_LAUNCH("gather_bytes", st, k_gather_bytes<<<(int)g, 256, 0, st>>>(data, starts, lens, idx, m, out_off, out));
    return DPK_OK;
}

int64_t dpk_dict_encode_workspace_bytes(int64_t n) { return dict_slots(n < 0 ? 0 : n) * 8; }

int dpk_dict_encode(const uint8_t *data, const int64_t *offsets, const int64_t *hash, int64_t n,
                    int64_t *out_rep, void *ws, int64_t ws_bytes, dpk_stream_t stream) {
    if (n < 0 || n >= ((int64_t)1 << 31)) return fail(DPK_ERR_INVALID, "n=%lld out of range [0, 2^31)", (long long)n);
    if (n == 0) return DPK_OK;
    if (!offsets || !hash || !out_rep || !ws) return fail(DPK_ERR_INVALID, "NULL pointer");
    const int64_t slots = dict_slots(n);
    if (ws_bytes < slots * 8) return fail(DPK_ERR_WORKSPACE, "workspace needs %lld B, got %lld", (long long)(slots * 8), (long long)ws_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    DPK_CUDA_TRY(cudaMemsetAsync(ws, 0xff, (size_t)slots * 8, st));
    int64_t g = (n + 255) / 256, cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    DPK_LAUNCH("dict_encode", st, k_dict_encode<<<(int)g, 256, 0, st>>>(data, offsets, hash, n, (long long *)ws,
                                                                       (uint64_t)(slots - 1), out_rep));
    return DPK_OK;
}

}  // extern "C"
