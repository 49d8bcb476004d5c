// dpk_aggregate3.cuh -- reduce-side implementation 2, final stage, register-pipelined form (included by
// dpk_combine.cu after dpk_aggregate2.cuh; dpk_set_option("agg_pipe", 1)).
//
// k_smem_aggregate2 runs every fine bucket as a chain of dependent latencies: ticket -> row range -> row loads
// (DRAM) -> barrier -> insert -> barrier -> output reservation (L2 atomic) -> barrier -> write-out.  Its r02 profile
// shows 27-30 % issue utilisation with the stall samples at the barriers behind the loads, the inserts and the
// reservation: three resident CTAs per SM are not enough independent chains to cover them.  This kernel takes the
// row loads OFF the chain:
//
//   * rows are not staged in shared memory.  A thread keeps its (up to) 8 rows of the bucket in REGISTERS and inserts
//     from there; the tag table still holds "row index + 1", and a row that meets an occupied slot compares its key
//     with keys[r0 + tag - 1] read through L1 (the CTA has just loaded those lines; ~5 % of the rows need it for
//     C2-like data).  Values are accumulated per CLAIMING row in a shared-memory array acc[row] that is kept at the
//     operation's identity between buckets: every row adds its value into acc[row that owns its key] with a native
//     shared atomic (the owner too), and the owner resets its entry when it writes the result out.
//   * as soon as a thread's rows are inserted their value registers are dead, and the NEXT bucket's rows are loaded
//     into registers right there -- the ticket and row range of the bucket after next are fetched by thread 0 at the
//     same time -- so the DRAM latency of bucket i+1 hides behind the barriers, the output reservation and the
//     write-out of bucket i.  No barrier between "rows arrive" and "insert".
//   * buckets that do not fit one window (more than AG2_CAP rows: hot keys) are appended to a list and merged by
//     k_smem_aggregate2 (staged windows + resident compaction) in a second launch that takes its tickets from the list.
//
// Output ranges are reserved with the atomic cursor (see dpk_aggregate2.cuh); shared memory per CTA: 16 KB of tags
// + 16 KB of accumulators.
#pragma once

struct Ag3Shared {
    int fb[3];                    // tickets: current, next (rows being prefetched), the one after (being fetched)
    long long r0[3], r1[3];
    unsigned long long excl;
    int wcnt[AG2_ITEMS * AG2_WARPS];
    int total;
};

template <typename KeyT, typename ValT, typename AccT, int MINB, bool BATCHED>
__global__ void __launch_bounds__(AG2_THREADS, MINB)
k_smem_aggregate3(const KeyT *__restrict__ keys, const ValT *__restrict__ vals, int op, int64_t ident,
                  const int64_t *__restrict__ fine_off, int32_t nfine, int32_t fine_per_part,
                  const int64_t *__restrict__ part_offsets, KeyT *__restrict__ out_keys,
                  int64_t *__restrict__ out_vals, long long *__restrict__ out_counts,
                  int *__restrict__ work_counter, int *__restrict__ big_list, int *__restrict__ big_count) {
    extern __shared__ __align__(16) long long s_dyn3[];   // [TAGS] u32 tags | [CAP] accumulators
    uint32_t *s_tag = reinterpret_cast<uint32_t *>(s_dyn3);
    long long *s_acc = s_dyn3 + AG2_TAGS / 2;
    const uint32_t tag_base = (uint32_t)__cvta_generic_to_shared(s_tag);
    const uint32_t acc_base = (uint32_t)__cvta_generic_to_shared(s_acc);
    __shared__ Ag3Shared sh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    auto clear_tags = [&]() {
        uint4 *t4 = reinterpret_cast<uint4 *>(s_tag);
#pragma unroll
        for (int i = 0; i < AG2_TAGS / 4 / AG2_THREADS; i++) t4[i * AG2_THREADS + threadIdx.x] = make_uint4(0, 0, 0, 0);
    };
    auto take_ticket = [&](int slot) {   // thread 0 only
        const int t = atomicAdd(work_counter, 1);
        sh.fb[slot] = t;
        long long a = 0, b = 0;
        if (t < nfine) { a = fine_off[t]; b = fine_off[t + 1]; }
        sh.r0[slot] = a;
        sh.r1[slot] = b;
    };
    // a thread's rows of the bucket [g0, g0 + n): row j * THREADS + tid for j < ITEMS (n <= CAP)
    auto load_rows = [&](int64_t g0, int n, KeyT (&kk)[AG2_ITEMS], ValT (&vv)[AG2_ITEMS]) {
#pragma unroll
        for (int j = 0; j < AG2_ITEMS; j++) {
            const int i = j * AG2_THREADS + (int)threadIdx.x;
            if (i < n) { kk[j] = keys[g0 + i]; vv[j] = vals[g0 + i]; }
        }
    };
    auto scan_claims = [&]() {   // warp 0: exclusive prefix of the 64 (item, warp) claim counts in place
        int a = sh.wcnt[lane], b = sh.wcnt[lane + 32];
        int ia = a, ib = b;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int ta = __shfl_up_sync(0xffffffffu, ia, d), tb = __shfl_up_sync(0xffffffffu, ib, d);
            if (lane >= d) { ia += ta; ib += tb; }
        }
        const int suma = __shfl_sync(0xffffffffu, ia, 31);
        sh.wcnt[lane] = ia - a;
        sh.wcnt[lane + 32] = suma + ib - b;
        if (lane == 31) sh.total = suma + ib;
    };

    clear_tags();
    for (int i = threadIdx.x; i < AG2_CAP; i += AG2_THREADS) s_acc[i] = (long long)ident;
    if (threadIdx.x == 0) { take_ticket(0); take_ticket(1); }
    __syncthreads();
    KeyT k[AG2_ITEMS], nk[AG2_ITEMS];
    ValT v[AG2_ITEMS], nv[AG2_ITEMS];
    {
        const int fb0 = sh.fb[0];
        const int64_t n0 = sh.r1[0] - sh.r0[0];
        if (fb0 < nfine && n0 <= AG2_CAP) load_rows(sh.r0[0], (int)n0, k, v);
    }
    for (int it = 0;; it++) {
        const int cur = it % 3, nxt = (it + 1) % 3, aft = (it + 2) % 3;
        const int fb = sh.fb[cur];
        if (fb >= nfine) break;                                     // uniform: tickets are handed out in order
        const int64_t r0 = sh.r0[cur], r1 = sh.r1[cur];
        const int p = fb / fine_per_part;
        const int64_t pbase = part_offsets[p];
        const bool big = r1 - r0 > AG2_CAP;
        const int n = big ? 0 : (int)(r1 - r0);
        if (big && threadIdx.x == 0) big_list[atomicAdd(big_count, 1)] = fb;   // merged by k_smem_aggregate2 afterwards
        // ticket of the bucket after next: the atomic is ISSUED here and consumed at the end of the iteration, its
        // round trip to L2 must not delay warp 0 on the way to the barriers
        int tk = 0;
        long long tk0 = 0, tk1 = 0;
        if (threadIdx.x == 0) tk = atomicAdd(work_counter, 1);

        // ---- insert from registers
        unsigned mine = 0;
        uint32_t offs[2] = {0u, 0u};
        int lc = 0;
        if constexpr (!BATCHED) {   // one probe loop per row (measured faster than four rows in flight: fewer instructions, no spills)
#pragma unroll
            for (int j = 0; j < AG2_ITEMS; j++) {
                const int idx = j * AG2_THREADS + (int)threadIdx.x;
                bool claimed = false;
                if (idx < n) {
                    const long long kbj = key_bits<KeyT>(k[j]);
                    uint32_t h = slot_hash32((uint64_t)kbj) & (AG2_TAGS - 1);
                    uint32_t tgt = (uint32_t)idx;
                    for (;;) {
                        uint32_t t = sm_ld_u32(tag_base + h * 4u);
                        if (t == 0u) {
                            t = sm_cas_u32(tag_base + h * 4u, 0u, (uint32_t)idx + 1u);
                            if (t == 0u) { claimed = true; break; }
                        }
                        if (key_bits<KeyT>(keys[r0 + (int64_t)(t - 1u)]) == kbj) { tgt = t - 1u; break; }
                        h = (h + 1u) & (AG2_TAGS - 1);
                    }
                    sm_apply<AccT>(op, acc_base + tgt * 8u, s_acc + tgt, (AccT)v[j]);
                }
                const unsigned cmj = __ballot_sync(0xffffffffu, claimed);
                if (lane == j) lc = __popc(cmj);
                offs[j >> 2] |= (uint32_t)__popc(cmj & lt) << (8 * (j & 3));
                if (claimed) mine |= 1u << j;
            }
        } else {
#pragma unroll
        for (int g = 0; g < AG2_ITEMS; g += 4) {
            long long kb[4];
            uint32_t h[4], t[4], tgt[4];
            bool live[4], claimed[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int idx = (g + u) * AG2_THREADS + (int)threadIdx.x;
                live[u] = idx < n;
                kb[u] = live[u] ? key_bits<KeyT>(k[g + u]) : 0ll;
                h[u] = slot_hash32((uint64_t)kb[u]) & (AG2_TAGS - 1);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) t[u] = live[u] ? sm_ld_u32(tag_base + h[u] * 4u) : 0xffffffffu;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int idx = (g + u) * AG2_THREADS + (int)threadIdx.x;
                claimed[u] = false;
                tgt[u] = (uint32_t)idx;
                if (t[u] == 0u) {
                    t[u] = sm_cas_u32(tag_base + h[u] * 4u, 0u, (uint32_t)idx + 1u);
                    claimed[u] = t[u] == 0u;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int idx = (g + u) * AG2_THREADS + (int)threadIdx.x;
                if (live[u] && !claimed[u]) {   // occupied: the owner's key comes through L1 (this CTA just loaded the line)
                    uint32_t hh = h[u], tt = t[u];
                    for (;;) {
                        if (tt == 0u) {
                            tt = sm_cas_u32(tag_base + hh * 4u, 0u, (uint32_t)idx + 1u);
                            if (tt == 0u) { claimed[u] = true; break; }
                        }
                        if (key_bits<KeyT>(keys[r0 + (int64_t)(tt - 1u)]) == kb[u]) { tgt[u] = tt - 1u; break; }
                        hh = (hh + 1u) & (AG2_TAGS - 1);
                        tt = sm_ld_u32(tag_base + hh * 4u);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (live[u]) sm_apply<AccT>(op, acc_base + tgt[u] * 8u, s_acc + tgt[u], (AccT)v[g + u]);
                const int j = g + u;
                const unsigned cmj = __ballot_sync(0xffffffffu, claimed[u]);
                if (lane == j) lc = __popc(cmj);
                offs[j >> 2] |= (uint32_t)__popc(cmj & lt) << (8 * (j & 3));
                if (claimed[u]) mine |= 1u << j;
            }
        }
        }
        // ---- the value registers are dead: fetch the NEXT bucket's rows now; thread 0 also fetches the ticket after it
        {
            const int nfb = sh.fb[nxt];
            const int64_t nn = sh.r1[nxt] - sh.r0[nxt];
            if (nfb < nfine && nn <= AG2_CAP) load_rows(sh.r0[nxt], (int)nn, nk, nv);
        }
        if (threadIdx.x == 0 && tk < nfine) { tk0 = fine_off[tk]; tk1 = fine_off[tk + 1]; }
        if (lane < AG2_ITEMS) sh.wcnt[lane * AG2_WARPS + warp] = lc;
        __syncthreads();                                            // (B) inserts done, claim counts written
        if (warp == 0) {
            scan_claims();
            __syncwarp();
            if (lane == 0)
                sh.excl = atomicAdd(reinterpret_cast<unsigned long long *>(&out_counts[p]), (unsigned long long)sh.total);
        } else {
            clear_tags();
        }
        __syncthreads();                                            // (D) output range known
        const int64_t obase = pbase + (int64_t)sh.excl;
#pragma unroll
        for (int j = 0; j < AG2_ITEMS; j++) {
            if (mine & (1u << j)) {
                const int idx = j * AG2_THREADS + (int)threadIdx.x;
                const int64_t o = obase + sh.wcnt[j * AG2_WARPS + warp] + ag2_off(offs, j);
                out_keys[o] = key_from_bits<KeyT>(key_bits<KeyT>(k[j]));
                out_vals[o] = s_acc[idx];
                s_acc[idx] = (long long)ident;                      // keep the accumulators at the identity between buckets
            }
        }
        if (warp == 0) clear_tags();
#pragma unroll
        for (int j = 0; j < AG2_ITEMS; j++) { k[j] = nk[j]; v[j] = nv[j]; }
        if (threadIdx.x == 0) { sh.fb[aft] = tk; sh.r0[aft] = tk0; sh.r1[aft] = tk1; }
        __syncthreads();                                            // (A) write-out reads, resets and tag clear done
    }
}
