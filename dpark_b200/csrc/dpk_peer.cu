// dpk_peer.cu -- the exchange step of the shuffle as a PUSH over NVLink peer memory.
//
// The reference's reducers pull every map's bucket over files + HTTP (ShuffleFetcher,
// dpark/shuffle.py:309-420).  After the map-side multisplit the rows bound for one peer GPU
// are one contiguous block of the bucket-major buffer (buckets owned by a peer are adjacent),
// so the exchange is G block copies per column.  k_copy_segments moves all of them in ONE
// launch: a segment table (source address, destination address, bytes) built on the device
// from the gathered bucket counts -- no host sync -- and a grid of 4 CTAs per SM that walks
// 32 KB work items ROUND-ROBIN over the segments, so every peer link carries traffic at the
// same time (a segment-after-segment order would aim all GPUs at the same peer).
// Destinations are plain device addresses: this GPU's memory or a peer's receive buffer mapped
// into this address space (torch symmetric memory = cuMem allocations exchanged on one node).
//
// Algorithmic bytes: (K+V) per row read locally + (K+V) per row stored through NVLink; the
// bound is the NVLink egress/ingress of one GPU (900 GB/s per direction nominal), not HBM.
#include "dpk_common.cuh"

namespace dpk {

constexpr int CP_THREADS = 256;
constexpr int CP_MAX_SEG = 1024;
constexpr int64_t CP_ITEM = 32768;  // bytes per work item

template <typename T>
__device__ __forceinline__ void copy_item(const unsigned char *__restrict__ s, unsigned char *__restrict__ d,
                                          int64_t bytes) {
    const T *sp = reinterpret_cast<const T *>(s);
    T *dp = reinterpret_cast<T *>(d);
    const int n = (int)(bytes / (int64_t)sizeof(T));
    int i = threadIdx.x;
    for (; i + 3 * CP_THREADS < n; i += 4 * CP_THREADS) {  // 4 independent loads in flight per thread
        const T a = sp[i], b = sp[i + CP_THREADS], c = sp[i + 2 * CP_THREADS], e = sp[i + 3 * CP_THREADS];
        dp[i] = a;
        dp[i + CP_THREADS] = b;
        dp[i + 2 * CP_THREADS] = c;
        dp[i + 3 * CP_THREADS] = e;
    }
    for (; i < n; i += CP_THREADS) dp[i] = sp[i];
}

__global__ void __launch_bounds__(CP_THREADS)
k_copy_segments(const uint64_t *__restrict__ src_ptrs, const uint64_t *__restrict__ dst_ptrs,
                const int64_t *__restrict__ nbytes, int32_t nseg) {
    __shared__ long long s_max;
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    long long mine = 0;
    for (int s = threadIdx.x; s < nseg; s += CP_THREADS) mine = max(mine, (long long)nbytes[s]);
    if (mine > 0) atomicMax(&s_max, mine);
    __syncthreads();
    const int64_t items_per_seg = ((int64_t)s_max + CP_ITEM - 1) / CP_ITEM;
    const int64_t total = items_per_seg * nseg;
    for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
        const int seg = (int)(w % nseg);
        const int64_t off = (w / nseg) * CP_ITEM;
        const int64_t len = nbytes[seg];
        if (off >= len) continue;
        const int64_t bytes = min(CP_ITEM, len - off);
        const uint64_t sa = src_ptrs[seg] + (uint64_t)off, da = dst_ptrs[seg] + (uint64_t)off;
        const unsigned align = (unsigned)((sa | da | (uint64_t)bytes) & 15u);
        const unsigned char *s = reinterpret_cast<const unsigned char *>(sa);
        unsigned char *d = reinterpret_cast<unsigned char *>(da);
        if (align == 0) copy_item<uint4>(s, d, bytes);
        else if ((align & 7u) == 0) copy_item<uint2>(s, d, bytes);
        else if ((align & 3u) == 0) copy_item<uint32_t>(s, d, bytes);
        else copy_item<unsigned char>(s, d, bytes);
    }
}

}  // namespace dpk

using namespace dpk;

extern "C" int dpk_copy_segments(const uint64_t *src_ptrs, const uint64_t *dst_ptrs, const int64_t *nbytes,
                                 int32_t nseg, dpk_stream_t stream) {
    if (nseg < 0 || nseg > CP_MAX_SEG)
        return fail(DPK_ERR_INVALID, "nseg=%d out of range [0, %d]", nseg, CP_MAX_SEG);
    if (nseg == 0) return DPK_OK;
    if (!src_ptrs || !dst_ptrs || !nbytes) return fail(DPK_ERR_INVALID, "segment table is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    DPK_LAUNCH("copy_segments", st,
               k_copy_segments<<<sm_count() * 4, CP_THREADS, 0, st>>>(src_ptrs, dst_ptrs, nbytes, nseg));
    return DPK_OK;
}
