// dpk_aggregate.cuh -- reduce-side implementation 2, final stage (included by
// dpk_combine.cu after Slot / key_bits / Acc are defined).
//
// The L2 round trip per probe is what bounds implementations 0/1 (ncu: long-scoreboard
// stalls on the probe loop, atomic units ~7 % busy).  Here every first-level bucket has
// been split once more (seg_multisplit, other hash bits) into fine buckets of ~1.5 k
// rows, and one CTA merges a fine bucket in a 4096-slot table in SHARED memory: a probe
// costs ~30 cycles instead of ~600.  Table accesses are explicit shared-space PTX
// (ld.volatile.shared / atom.shared.cas / red.shared.add), not generic atomics.
//
// Output placement without atomics on the critical path: fine buckets are handed out
// in order (atomic work counter), and the position of a fine bucket's output inside its
// partition is the running sum of the distinct counts of the fine buckets before it --
// a chained scan with decoupled look-back over fb_state[] (AGGREGATE then INCLUSIVE
// words, 2 flag bits + 62 value bits).  The partition's final count is the inclusive
// value of its last fine bucket.
//
// A fine bucket with more distinct keys than the table holds is processed in
// hash-disjoint passes (m, r): rows with ((mix >> 40) & (m-1)) == r, split on demand
// (probe sequences longer than AG_MAX_PROBE declare the pass overflowed).
#pragma once

constexpr int AG_THREADS = 256;
constexpr int AG_CAP = 4096;
constexpr int AG_LIMIT = AG_CAP - AG_CAP / 8;  // upper end for "rows per fine bucket" settings
constexpr int AG_MAX_PROBE = 96;
constexpr int AG_STACK = 96;
constexpr int AG_UNROLL = 4;

constexpr unsigned long long AG_FLAG_AGG = 1ull << 62;
constexpr unsigned long long AG_FLAG_INC = 2ull << 62;
constexpr unsigned long long AG_VAL_MASK = (1ull << 62) - 1;

// ---- shared-space table primitives (32-bit shared addresses) -------------------
__device__ __forceinline__ long long sm_ld_volatile(uint32_t a) {
    long long v;
    asm volatile("ld.volatile.shared.b64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ long long sm_cas(uint32_t a, long long cmp, long long val) {
    long long old;
    asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(a), "l"(cmp), "l"(val) : "memory");
    return old;
}
__device__ __forceinline__ void sm_red_add_u64(uint32_t a, long long v) {
    asm volatile("red.shared.add.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}
__device__ __forceinline__ void sm_red_add_f64(uint32_t a, double v) {
    asm volatile("red.shared.add.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
// accumulate v into the shared-memory accumulator at 32-bit address a (generic pointer g for the rare ops)
template <typename AccT>
__device__ __forceinline__ void sm_apply(int op, uint32_t a, long long *g, AccT v) {
    if (op == DPK_OP_SUM) {
        if constexpr (std::is_same<AccT, double>::value) sm_red_add_f64(a, v);
        else sm_red_add_u64(a, (long long)v);
    } else {
        Acc<AccT>::apply(op, (int64_t *)g, v);
    }
}

// exclusive prefix of the distinct counts of fine buckets [first_fb, fb) -- warp 0 only
__device__ __forceinline__ unsigned long long ag_look_back(const unsigned long long *state, int first_fb, int fb) {
    const int lane = threadIdx.x & 31;
    unsigned long long excl = 0;
    int idx = fb - 1;
    while (idx >= first_fb) {
        const int my = idx - lane;
        unsigned long long s;
        do {
            s = my >= first_fb ? *(volatile const unsigned long long *)&state[my] : AG_FLAG_INC;
        } while (__any_sync(0xffffffffu, (s >> 62) == 0));
        const unsigned inc_mask = __ballot_sync(0xffffffffu, (s >> 62) == 2);
        const int stop = inc_mask ? __ffs(inc_mask) - 1 : 32;  // nearest predecessor with an inclusive value
        unsigned long long v = lane <= stop ? (s & AG_VAL_MASK) : 0;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        excl += v;
        if (inc_mask) break;
        idx -= 32;
    }
    return excl;
}

// insert one row; returns false when the probe sequence got too long (table too full)
template <typename AccT>
__device__ __forceinline__ bool ag_insert(uint32_t key_base, uint32_t acc_base, long long *s_acc, int op, int64_t kb,
                                          uint64_t mx, AccT v) {
    uint32_t h = (uint32_t)mx & (AG_CAP - 1);
#pragma unroll 1
    for (int steps = 0; steps < AG_MAX_PROBE; steps++) {
        const uint32_t ka = key_base + h * 8u;
        long long cur = sm_ld_volatile(ka);
        if (cur == kEmpty) cur = sm_cas(ka, kEmpty, kb);   // old value: kEmpty = claimed, kb = someone else claimed it
        if (cur == kb || cur == kEmpty) {
            sm_apply<AccT>(op, acc_base + h * 8u, s_acc + h, v);
            return true;
        }
        h = (h + 1) & (AG_CAP - 1);
    }
    return false;
}

template <typename KeyT, typename ValT, typename AccT>
__global__ void __launch_bounds__(AG_THREADS)
k_smem_aggregate(const KeyT *__restrict__ keys, const ValT *__restrict__ vals, int op, int64_t ident,
                 const int64_t *__restrict__ fine_off, int32_t nfine, int32_t fine_per_part,
                 const int64_t *__restrict__ part_offsets, KeyT *__restrict__ out_keys,
                 int64_t *__restrict__ out_vals, unsigned long long *__restrict__ out_counts,
                 unsigned long long *__restrict__ fb_state, int *__restrict__ work_counter) {
    extern __shared__ __align__(16) long long s_dyn[];  // [AG_CAP] keys | [AG_CAP] accumulators
    long long *s_key = s_dyn;
    long long *s_acc = s_dyn + AG_CAP;
    const uint32_t key_base = (uint32_t)__cvta_generic_to_shared(s_key);
    const uint32_t acc_base = (uint32_t)__cvta_generic_to_shared(s_acc);
    __shared__ int s_fb, s_overflow, s_side_used, s_sp;
    __shared__ long long s_side_acc;
    __shared__ unsigned long long s_excl;
    __shared__ int s_stack_m[AG_STACK], s_stack_r[AG_STACK];
    __shared__ int s_wsum[AG_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        if (threadIdx.x == 0) s_fb = atomicAdd(work_counter, 1);  // in-order hand-out
        __syncthreads();
        const int fb = s_fb;
        if (fb >= nfine) break;
        const int64_t r0 = fine_off[fb], r1 = fine_off[fb + 1];
        const int p = fb / fine_per_part;
        const int first_fb = p * fine_per_part;
        const int64_t pbase = part_offsets[p];
        unsigned long long written = 0;   // distinct pairs of this fine bucket written so far (uniform)
        bool have_excl = false;
        if (threadIdx.x == 0) { s_stack_m[0] = 1; s_stack_r[0] = 0; s_sp = r1 > r0 ? 1 : 0; }
        __syncthreads();
        while (s_sp > 0) {
            __syncthreads();  // everyone has seen s_sp > 0
            const int m = s_stack_m[s_sp - 1], r = s_stack_r[s_sp - 1];
            __syncthreads();
            if (threadIdx.x == 0) { s_sp--; s_overflow = 0; s_side_used = 0; s_side_acc = ident; }
            for (int i = threadIdx.x; i < AG_CAP; i += AG_THREADS) { s_key[i] = kEmpty; s_acc[i] = ident; }
            __syncthreads();
            bool ok = true;
            const int64_t nfull = (r1 - r0) / ((int64_t)AG_THREADS * AG_UNROLL) * ((int64_t)AG_THREADS * AG_UNROLL);
            // ---- full batches: AG_UNROLL independent row loads in flight per thread, no bounds checks
            for (int64_t base = r0; base < r0 + nfull; base += (int64_t)AG_THREADS * AG_UNROLL) {
                KeyT kreg[AG_UNROLL];
                ValT vreg[AG_UNROLL];
#pragma unroll
                for (int u = 0; u < AG_UNROLL; u++) {
                    kreg[u] = keys[base + u * AG_THREADS + threadIdx.x];
                    vreg[u] = vals[base + u * AG_THREADS + threadIdx.x];
                }
#pragma unroll
                for (int u = 0; u < AG_UNROLL; u++) {
                    const int64_t kb = key_bits<KeyT>(kreg[u]);
                    const uint64_t mx = mix64((uint64_t)kb);
                    if (m > 1 && (int)((mx >> 40) & (uint64_t)(m - 1)) != r) continue;
                    if (kb == kEmpty) {
                        s_side_used = 1;
                        Acc<AccT>::apply(op, (int64_t *)&s_side_acc, (AccT)vreg[u]);
                    } else {
                        ok &= ag_insert<AccT>(key_base, acc_base, s_acc, op, kb, mx, (AccT)vreg[u]);
                    }
                }
                if (!ok) s_overflow = 1;
                if (*(volatile int *)&s_overflow) break;
            }
            // ---- tail
            for (int64_t i = r0 + nfull + threadIdx.x; i < r1; i += AG_THREADS) {
                const int64_t kb = key_bits<KeyT>(keys[i]);
                const uint64_t mx = mix64((uint64_t)kb);
                if (m > 1 && (int)((mx >> 40) & (uint64_t)(m - 1)) != r) continue;
                const AccT v = (AccT)vals[i];
                if (kb == kEmpty) {
                    s_side_used = 1;
                    Acc<AccT>::apply(op, (int64_t *)&s_side_acc, v);
                } else if (!ag_insert<AccT>(key_base, acc_base, s_acc, op, kb, mx, v)) {
                    s_overflow = 1;
                }
            }
            __syncthreads();
            if (s_overflow) {  // uniform after the barrier: split this pass in two and retry
                if (threadIdx.x == 0 && s_sp + 2 <= AG_STACK) {
                    s_stack_m[s_sp] = m * 2; s_stack_r[s_sp] = r; s_sp++;
                    s_stack_m[s_sp] = m * 2; s_stack_r[s_sp] = r + m; s_sp++;
                }
                __syncthreads();
                continue;
            }
            // ---- count + rank the occupied slots
            constexpr int PER = AG_CAP / AG_THREADS;
            long long kslot[PER];
            int c = 0;
#pragma unroll
            for (int j = 0; j < PER; j++) {
                kslot[j] = s_key[threadIdx.x * PER + j];
                c += kslot[j] != kEmpty;
            }
            int inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            if (lane == 31) s_wsum[warp] = inc;
            __syncthreads();
            int wbase = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < AG_THREADS / 32; w++) {
                const int t = s_wsum[w];
                if (w < warp) wbase += t;
                tot += t;
            }
            const int side = s_side_used ? 1 : 0;
            // ---- where does this fine bucket start inside its partition?  (first pass only)
            if (!have_excl) {
                if (warp == 0) {
                    // single-pass buckets publish their count right away so successors do not wait for our look-back
                    if (lane == 0 && s_sp == 0) atomicExch(&fb_state[fb], AG_FLAG_AGG | (unsigned long long)(tot + side));
                    const unsigned long long e = ag_look_back(fb_state, first_fb, fb);
                    if (lane == 0) s_excl = e;
                }
                __syncthreads();
                have_excl = true;
            }
            int64_t dst = pbase + (int64_t)(s_excl + written) + wbase + (inc - c);
#pragma unroll
            for (int j = 0; j < PER; j++) {
                if (kslot[j] != kEmpty) {
                    out_keys[dst] = key_from_bits<KeyT>(kslot[j]);
                    out_vals[dst] = s_acc[threadIdx.x * PER + j];
                    dst++;
                }
            }
            if (side && threadIdx.x == 0) {
                out_keys[pbase + (int64_t)(s_excl + written) + tot] = key_from_bits<KeyT>(kEmpty);
                out_vals[pbase + (int64_t)(s_excl + written) + tot] = s_side_acc;
            }
            written += (unsigned long long)(tot + side);
            __syncthreads();
        }
        // ---- publish the inclusive value (empty buckets too, so chains stay short)
        if (warp == 0) {
            unsigned long long e;
            if (have_excl) e = s_excl;
            else e = ag_look_back(fb_state, first_fb, fb);
            if (lane == 0) {
                __threadfence();
                atomicExch(&fb_state[fb], AG_FLAG_INC | (e + written));
                if (fb == first_fb + fine_per_part - 1) out_counts[p] = e + written;
            }
        }
        __syncthreads();
    }
}
