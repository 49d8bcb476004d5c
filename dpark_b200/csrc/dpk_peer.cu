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

// One launch instead of ~20 small tensor operations on the host side of every step: from the gathered counts matrix
// (the MapOutputTracker) to the segment table of this rank's pushes, the rows every rank will receive and the segment
// matrix of this rank's own partitions.  Single CTA; G <= 64 ranks, any F.
//   all_counts[S][F]   rows source s holds for fine bucket b (S = G sources, or G * H when a rank sends H groups)
//   my_src             this rank's source row (rank, or rank * H + group)
//   per_blk            fine buckets per destination block (ceil(P / G) << sub_bits); destination d owns
//                      [d * per_blk, min(F, (d + 1) * per_blk))
//   src0/src1, dst_base[c][G], elem0/elem1   column c (keys, values): address of my bucket-major buffer, of every
//                      rank's receive buffer, element size
// Outputs: src_ptrs / dst_ptrs / nbytes [ncols][G] (clamped so that nothing is written past `capacity` rows of a
// receive buffer), need_over = max(need_over, max_d rows d receives - capacity), seg_out[S][Fown] (Fown = my block).
__global__ void __launch_bounds__(256)
k_push_plan(const int64_t *__restrict__ all_counts, int32_t S, int32_t G, int32_t F, int32_t per_blk, int32_t my_src,
            int32_t my_rank, int32_t ncols, uint64_t src0, uint64_t src1, const uint64_t *__restrict__ dst_base,
            int32_t elem0, int32_t elem1, int64_t capacity, uint64_t *__restrict__ src_ptrs,
            uint64_t *__restrict__ dst_ptrs, int64_t *__restrict__ nbytes, long long *__restrict__ need_over,
            int64_t *__restrict__ seg_out) {
    extern __shared__ long long s_R[];   // [S][G] rows source s sends to destination d
    for (int i = threadIdx.x; i < S * G; i += blockDim.x) {
        const int s = i / G, d = i % G;
        const int b0 = min(F, d * per_blk), b1 = min(F, (d + 1) * per_blk);
        long long r = 0;
        for (int b = b0; b < b1; b++) r += all_counts[(int64_t)s * F + b];
        s_R[i] = r;
    }
    __syncthreads();
    if (threadIdx.x < G) {
        const int d = threadIdx.x;
        long long send_first = 0, dst_first = 0, total = 0;
        for (int dd = 0; dd < d; dd++) send_first += s_R[my_src * G + dd];
        for (int s = 0; s < S; s++) {
            if (s < my_src) dst_first += s_R[s * G + d];
            total += s_R[s * G + d];
        }
        long long rows = s_R[my_src * G + d];
        const long long room = capacity - dst_first;
        if (rows > room) rows = room > 0 ? room : 0;
        for (int c = 0; c < ncols; c++) {
            const long long e = c ? elem1 : elem0;
            src_ptrs[c * G + d] = (c ? src1 : src0) + (uint64_t)(send_first * e);
            dst_ptrs[c * G + d] = dst_base[c * G + d] + (uint64_t)(dst_first * e);
            nbytes[c * G + d] = rows * e;
        }
        if (total > capacity) atomicMax(need_over, total - capacity);
    }
    if (seg_out) {
        const int b0 = min(F, my_rank * per_blk), b1 = min(F, (my_rank + 1) * per_blk), fo = b1 - b0;
        for (int i = threadIdx.x; i < S * fo; i += blockDim.x) seg_out[i] = all_counts[(int64_t)(i / fo) * F + b0 + i % fo];
    }
}

// The plan of the FUSED scatter + exchange (dpk_partition_scatter_ptrs): for every fine bucket b of this rank's map
// output the address its rows go to -- the slot of (source = this rank, bucket b) in the OWNER's receive buffer, layout
// source-rank-major then bucket-major exactly as the push form delivers it -- plus this rank's segment matrix and the
// capacity flag.  A bucket that would end past `capacity` rows of its receive buffer is pointed at a local dump buffer
// instead (dump0/dump1: >= this rank's row count; position = the bucket's local bucket-major offset), so a too-small
// buffer can never be overrun; need_over reports it.  Single CTA; G <= 64 ranks, F <= 4096 buckets.
__global__ void __launch_bounds__(256)
k_fused_plan(const int64_t *__restrict__ all_counts, int32_t G, int32_t F, int32_t per_blk, int32_t my_rank, int32_t ncols,
             const uint64_t *__restrict__ dst_base, int32_t elem0, int32_t elem1, int64_t capacity, uint64_t dump0,
             uint64_t dump1, uint64_t *__restrict__ key_ptrs, uint64_t *__restrict__ val_ptrs,
             long long *__restrict__ need_over, int64_t *__restrict__ seg_out) {
    extern __shared__ long long s_fp[];
    long long *s_R = s_fp;             // [G][G] rows source s sends to destination d
    long long *s_mine = s_fp + G * G;  // [F] my rows per bucket
    for (int i = threadIdx.x; i < G * G; i += blockDim.x) {
        const int s = i / G, d = i % G;
        const int b0 = min(F, d * per_blk), b1 = min(F, (d + 1) * per_blk);
        long long r = 0;
        for (int b = b0; b < b1; b++) r += all_counts[(int64_t)s * F + b];
        s_R[i] = r;
    }
    for (int b = threadIdx.x; b < F; b += blockDim.x) s_mine[b] = all_counts[(int64_t)my_rank * F + b];
    __syncthreads();
    if (threadIdx.x < G) {
        const int d = threadIdx.x;
        long long dst_first = 0, total = 0, local_first = 0;
        for (int dd = 0; dd < d; dd++) local_first += s_R[my_rank * G + dd];
        for (int s = 0; s < G; s++) {
            if (s < my_rank) dst_first += s_R[s * G + d];
            total += s_R[s * G + d];
        }
        const int b0 = min(F, d * per_blk), b1 = min(F, (d + 1) * per_blk);
        long long run = 0;
        for (int b = b0; b < b1; b++) {
            const long long c = s_mine[b];
            const bool fits = dst_first + run + c <= capacity;
            key_ptrs[b] = fits ? dst_base[d] + (uint64_t)((dst_first + run) * elem0) : dump0 + (uint64_t)((local_first + run) * elem0);
            if (ncols > 1)
                val_ptrs[b] = fits ? dst_base[G + d] + (uint64_t)((dst_first + run) * elem1)
                                   : dump1 + (uint64_t)((local_first + run) * elem1);
            run += c;
        }
        if (total > capacity) atomicMax(need_over, total - capacity);
    }
    if (seg_out) {
        const int b0 = min(F, my_rank * per_blk), b1 = min(F, (my_rank + 1) * per_blk), fo = b1 - b0;
        for (int i = threadIdx.x; i < G * fo; i += blockDim.x) seg_out[i] = all_counts[(int64_t)(i / fo) * F + b0 + i % fo];
    }
}

}  // namespace dpk

using namespace dpk;

extern "C" int dpk_fused_plan(const int64_t *all_counts, int32_t nranks, int32_t nbuckets, int32_t per_block, int32_t my_rank,
                              int32_t ncols, const uint64_t *dst_base, int32_t key_bytes, int32_t val_bytes, int64_t capacity,
                              uint64_t dump_keys, uint64_t dump_vals, uint64_t *key_ptrs, uint64_t *val_ptrs,
                              int64_t *need_over, int64_t *seg_out, dpk_stream_t stream) {
    if (nranks < 1 || nranks > 64 || nbuckets < 1 || nbuckets > 4096 || ncols < 1 || ncols > 2 || per_block < 0 ||
        my_rank < 0 || my_rank >= nranks)
        return fail(DPK_ERR_INVALID, "bad fused plan shape: %d ranks, %d buckets, %d columns", nranks, nbuckets, ncols);
    if (!all_counts || !dst_base || !key_ptrs || (ncols > 1 && !val_ptrs) || !need_over || !dump_keys || (ncols > 1 && !dump_vals))
        return fail(DPK_ERR_INVALID, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t sh = ((size_t)nranks * nranks + nbuckets) * sizeof(long long);
    DPK_CUDA_TRY(cudaFuncSetAttribute(k_fused_plan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    DPK_LAUNCH("fused_plan", st, k_fused_plan<<<1, 256, sh, st>>>(all_counts, nranks, nbuckets, per_block, my_rank, ncols, dst_base,
                                                               key_bytes, val_bytes, capacity, dump_keys, dump_vals, key_ptrs,
                                                               val_ptrs, (long long *)need_over, seg_out));
    return DPK_OK;
}

extern "C" int dpk_push_plan(const int64_t *all_counts, int32_t nsrc, int32_t nranks, int32_t nbuckets, int32_t per_block,
                             int32_t my_src, int32_t my_rank, int32_t ncols, uint64_t src_keys, uint64_t src_vals,
                             const uint64_t *dst_base, int32_t key_bytes, int32_t val_bytes, int64_t capacity, uint64_t *src_ptrs,
                             uint64_t *dst_ptrs, int64_t *nbytes, int64_t *need_over, int64_t *seg_out,
                             dpk_stream_t stream) {
    if (nranks < 1 || nranks > 64 || nsrc < nranks || nsrc > 4096 || ncols < 1 || ncols > 2 || per_block < 0)
        return fail(DPK_ERR_INVALID, "bad push plan shape: %d sources, %d ranks, %d columns", nsrc, nranks, ncols);
    if (!all_counts || !dst_base || !src_ptrs || !dst_ptrs || !nbytes || !need_over)
        return fail(DPK_ERR_INVALID, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t sh = (size_t)nsrc * nranks * sizeof(long long);
    if (sh > 48 * 1024) return fail(DPK_ERR_UNSUPPORTED, "push plan: %d x %d sources x ranks exceed shared memory", nsrc, nranks);
    DPK_LAUNCH("push_plan", st, k_push_plan<<<1, 256, sh, st>>>(all_counts, nsrc, nranks, nbuckets, per_block, my_src, my_rank,
                                                             ncols, src_keys, src_vals, dst_base, key_bytes, val_bytes, capacity, src_ptrs,
                                                             dst_ptrs, nbytes, (long long *)need_over, seg_out));
    return DPK_OK;
}

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
