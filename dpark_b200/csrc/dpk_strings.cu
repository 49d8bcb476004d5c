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

static inline int64_t dict_slots(int64_t n) {
    int64_t s = 1024;
    while (s < 2 * n) s <<= 1;
    return s;
}

}  // namespace dpk

using namespace dpk;

extern "C" {

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
