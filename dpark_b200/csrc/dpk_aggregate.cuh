// dpk_aggregate.cuh -- reduce-side implementation 2, final stage (included by
// dpk_combine.cu after Slot / key_bits / Acc are defined).
//
// The L2 round trip per probe is what bounds implementations 0/1 (ncu: long-scoreboard
// stalls on the probe loop, atomic units ~7 % busy).  Here every first-level bucket has
// been split once more (seg_multisplit, other hash bits) into fine buckets of ~1.5-2 k
// rows, and one CTA merges a fine bucket in a 4096-slot table in SHARED memory: a probe
// costs ~30 cycles instead of ~600.  Table accesses are explicit shared-space PTX
// (ld.volatile.shared / atom.shared.cas / atom.shared.add), not generic atomics.
//
// Claim list: a row that claims a free slot appends the slot index to a per-CTA list
// (warp-aggregated counter), so the number of distinct keys is known the moment the
// inserts are done, the output is written by walking the list (coalesced, no sweep of
// the 4096 slots, no block scan) and only the touched slots are reset for the next bucket.
//
// Output placement without atomics on the critical path: fine buckets are handed out
// in order (atomic work counter), and the position of a fine bucket's output inside its
// partition is the running sum of the distinct counts of the fine buckets before it --
// a chained scan with decoupled look-back over fb_state[] (AGGREGATE then INCLUSIVE
// words, 2 flag bits + 62 value bits).  The partition's final count is the inclusive
// value of its last fine bucket.
//
// A fine bucket with more distinct keys than the table holds takes the slow path:
// hash-disjoint passes (m, r), rows with ((mix >> 40) & (m-1)) == r, split on demand.
#pragma once

constexpr int AG_THREADS = 256;
constexpr int AG_CAP = 4096;
constexpr int AG_LIMIT = AG_CAP - AG_CAP / 8;  // distinct keys a pass may hold
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
// 128-bit compare-and-swap on a whole {key, accumulator} slot (native ATOMS.CAS.128 on sm_100): a row
// that finds a free slot claims it AND deposits its value in one atomic; returns the old key word.
__device__ __forceinline__ long long sm_cas_slot(uint32_t a, long long ckey, long long cacc, long long nkey, long long nacc) {
    long long olo;
    [[maybe_unused]] long long ohi;  // the old accumulator word: not needed, the key word tells who owns the slot
    asm volatile("{\n\t.reg .b128 c, s, o;\n\tmov.b128 c, {%3, %4};\n\tmov.b128 s, {%5, %6};\n\t"
                 "atom.shared.cas.b128 o, [%2], c, s;\n\tmov.b128 {%0, %1}, o;\n\t}"
                 : "=l"(olo), "=l"(ohi) : "r"(a), "l"(ckey), "l"(cacc), "l"(nkey), "l"(nacc) : "memory");
    return olo;
}
// 64-bit integer add into shared memory.  sm_100 has no native 64-bit shared-memory add
// (red.shared.add.u64 compiles to an ATOMS.CAST.SPIN.64 retry loop), so the add is done on
// the two 32-bit halves with native atomics: the low half returns its old value, from which
// this row's carry follows exactly; the high half gets (hi + carry) when that is non-zero.
// Additions commute and every carry is accounted for by the row that produced it, so the
// final 64-bit word is exact under any interleaving.
__device__ __forceinline__ void sm_red_add_u64(uint32_t a, long long v) {
    const uint32_t vlo = (uint32_t)(unsigned long long)v, vhi = (uint32_t)((unsigned long long)v >> 32);
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(vlo) : "memory");
    const uint32_t addhi = vhi + ((uint32_t)(old + vlo) < old ? 1u : 0u);
    if (addhi) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a + 4u), "r"(addhi) : "memory");
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

struct AgShared {
    int fb, overflow, side_used, sp, nclaim;
    long long side_acc;
    unsigned long long excl;
    int stack_m[AG_STACK], stack_r[AG_STACK];
};

// insert the rows [r0, r1) whose pass id matches (m, r).  Whole warps walk the rows together
// (predicated on validity) so the claim ballot below is always converged.
template <typename KeyT, typename ValT, typename AccT, bool WIDE>
__device__ __forceinline__ void ag_insert_rows(const KeyT *__restrict__ keys, const ValT *__restrict__ vals,
                                               int64_t r0, int64_t r1, int m, int r, int op, long long ident,
                                               uint32_t tab_base, long long *s_tab, uint16_t *s_list, AgShared &sh) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t step = (int64_t)AG_THREADS * AG_UNROLL;
    for (int64_t base = r0; base < r1; base += step) {
        KeyT kreg[AG_UNROLL];
        ValT vreg[AG_UNROLL];
#pragma unroll
        for (int u = 0; u < AG_UNROLL; u++) {  // AG_UNROLL independent loads in flight per thread
            const int64_t i = base + u * AG_THREADS + threadIdx.x;
            if (i < r1) { kreg[u] = keys[i]; vreg[u] = vals[i]; }
        }
#pragma unroll
        for (int u = 0; u < AG_UNROLL; u++) {
            const int64_t i = base + u * AG_THREADS + threadIdx.x;
            const int64_t kb = i < r1 ? key_bits<KeyT>(kreg[u]) : 0;
            const uint64_t mx = mix64((uint64_t)kb);
            bool live = i < r1 && (m == 1 || (int)((mx >> 40) & (uint64_t)(m - 1)) == r);
            if (live && kb == kEmpty) {  // the key whose bits equal the free-slot marker: side accumulator
                sh.side_used = 1;
                Acc<AccT>::apply(op, (int64_t *)&sh.side_acc, (AccT)vreg[u]);
                live = false;
            }
            bool claimed = false;
            uint32_t h = (uint32_t)mx & (AG_CAP - 1);
            if (live) {
                bool placed = false;
#pragma unroll 1
                for (int steps = 0; steps < AG_MAX_PROBE; steps++) {
                    const uint32_t ka = tab_base + h * 16u;
                    long long cur = sm_ld_volatile(ka);
                    if (cur == kEmpty) {  // old value: kEmpty = we claimed it, kb = a peer did
                        if constexpr (WIDE) {
                            // the first value of a key IS its combiner (createCombiner = identity function,
                            // dpark/rdd.py:303-327): claim the slot and deposit the value in one atomic
                            long long first;
                            if constexpr (std::is_same<AccT, double>::value) first = __double_as_longlong((double)vreg[u]);
                            else first = (long long)vreg[u];
                            cur = sm_cas_slot(ka, kEmpty, ident, kb, first);
                        } else {
                            cur = sm_cas(ka, kEmpty, kb);
                        }
                        claimed = cur == kEmpty;
                    }
                    if (cur == kb || claimed) { placed = true; break; }
                    h = (h + 1) & (AG_CAP - 1);
                }
                if (placed) {
                    if (!(WIDE && claimed)) sm_apply<AccT>(op, tab_base + h * 16u + 8u, s_tab + 2 * h + 1, (AccT)vreg[u]);
                } else {
                    sh.overflow = 1;  // table too full for this pass: it will be split
                }
            }
            // ---- append the newly claimed slots to the claim list (one atomic per warp)
            const unsigned cm = __ballot_sync(0xffffffffu, claimed);
            if (cm) {
                const int leader = __ffs(cm) - 1;
                int pos = 0;
                if (lane == leader) pos = atomicAdd(&sh.nclaim, __popc(cm));
                pos = __shfl_sync(0xffffffffu, pos, leader);
                if (claimed) {
                    const int at = pos + __popc(cm & lt);
                    if (at < AG_CAP) s_list[at] = (uint16_t)h;
                }
                if (pos + __popc(cm) > AG_LIMIT) sh.overflow = 1;
            }
        }
        // warp-uniform exit (the ballots above need whole warps)
        if (__any_sync(0xffffffffu, *(volatile int *)&sh.overflow != 0)) return;
    }
}

template <typename KeyT, typename ValT, typename AccT, bool WIDE>
__global__ void __launch_bounds__(AG_THREADS)
k_smem_aggregate(const KeyT *__restrict__ keys, const ValT *__restrict__ vals, int op, int64_t ident,
                 const int64_t *__restrict__ fine_off, int32_t nfine, int32_t fine_per_part,
                 const int64_t *__restrict__ part_offsets, KeyT *__restrict__ out_keys,
                 int64_t *__restrict__ out_vals, unsigned long long *__restrict__ out_counts,
                 unsigned long long *__restrict__ fb_state, int *__restrict__ work_counter) {
    extern __shared__ __align__(16) long long s_dyn[];  // [AG_CAP] {key, accumulator} slots | [AG_CAP] u16 claim list
    longlong2 *s_slot = reinterpret_cast<longlong2 *>(s_dyn);
    uint16_t *s_list = reinterpret_cast<uint16_t *>(s_dyn + 2 * AG_CAP);
    const uint32_t tab_base = (uint32_t)__cvta_generic_to_shared(s_dyn);
    const longlong2 kFree = make_longlong2(kEmpty, ident);
    __shared__ AgShared sh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // the table is kept clean between fine buckets: writing a bucket out resets exactly the slots it touched
    for (int i = threadIdx.x; i < AG_CAP; i += AG_THREADS) s_slot[i] = kFree;
    if (threadIdx.x == 0) { sh.overflow = 0; sh.side_used = 0; sh.side_acc = ident; sh.nclaim = 0; }
    for (;;) {
        if (threadIdx.x == 0) sh.fb = atomicAdd(work_counter, 1);  // in-order hand-out
        __syncthreads();                                            // (A) also: previous write-out finished, table clean
        const int fb = sh.fb;
        if (fb >= nfine) break;
        const int64_t r0 = fine_off[fb], r1 = fine_off[fb + 1];
        const int p = fb / fine_per_part;
        const int first_fb = p * fine_per_part;
        const int64_t pbase = part_offsets[p];

        // ---- fast path: one pass over all rows
        ag_insert_rows<KeyT, ValT, AccT, WIDE>(keys, vals, r0, r1, 1, 0, op, ident, tab_base, s_dyn, s_list, sh);
        __syncthreads();                                            // (B)
        if (sh.overflow == 0) {                                     // uniform
            const int cnt = sh.nclaim, side = sh.side_used ? 1 : 0;
            if (warp == 0) {
                if (lane == 0) atomicExch(&fb_state[fb], AG_FLAG_AGG | (unsigned long long)(cnt + side));
                const unsigned long long e = ag_look_back(fb_state, first_fb, fb);
                if (lane == 0) {
                    sh.excl = e;
                    atomicExch(&fb_state[fb], AG_FLAG_INC | (e + (unsigned long long)(cnt + side)));
                    if (fb == first_fb + fine_per_part - 1) out_counts[p] = e + (unsigned long long)(cnt + side);
                }
            }
            __syncthreads();                                        // (D)
            const int64_t obase = pbase + (int64_t)sh.excl;
            for (int j = threadIdx.x; j < cnt; j += AG_THREADS) {   // coalesced: list order is output order
                const int s = s_list[j];
                const longlong2 e = s_slot[s];
                out_keys[obase + j] = key_from_bits<KeyT>(e.x);
                out_vals[obase + j] = e.y;
                s_slot[s] = kFree;
            }
            if (threadIdx.x == 0) {
                if (side) {
                    out_keys[obase + cnt] = key_from_bits<KeyT>(kEmpty);
                    out_vals[obase + cnt] = sh.side_acc;
                    sh.side_used = 0;
                    sh.side_acc = ident;
                }
                sh.nclaim = 0;
            }
            continue;  // barrier (A) of the next iteration orders the resets before the next inserts
        }

        // ---- slow path: hash-disjoint passes, split on demand
        unsigned long long written = 0, excl = 0;  // uniform
        bool have_excl = false;
        if (threadIdx.x == 0) { sh.stack_m[0] = 2; sh.stack_r[0] = 0; sh.stack_m[1] = 2; sh.stack_r[1] = 1; sh.sp = 2; }
        __syncthreads();
        while (sh.sp > 0) {
            __syncthreads();  // everyone has seen sp > 0
            const int m = sh.stack_m[sh.sp - 1], r = sh.stack_r[sh.sp - 1];
            __syncthreads();
            if (threadIdx.x == 0) { sh.sp--; sh.overflow = 0; sh.side_used = 0; sh.side_acc = ident; sh.nclaim = 0; }
            for (int i = threadIdx.x; i < AG_CAP; i += AG_THREADS) s_slot[i] = kFree;
            __syncthreads();
            ag_insert_rows<KeyT, ValT, AccT, WIDE>(keys, vals, r0, r1, m, r, op, ident, tab_base, s_dyn, s_list, sh);
            __syncthreads();
            if (sh.overflow) {  // uniform after the barrier: split this pass in two and retry
                if (threadIdx.x == 0 && sh.sp + 2 <= AG_STACK) {
                    sh.stack_m[sh.sp] = m * 2; sh.stack_r[sh.sp] = r; sh.sp++;
                    sh.stack_m[sh.sp] = m * 2; sh.stack_r[sh.sp] = r + m; sh.sp++;
                }
                __syncthreads();
                continue;
            }
            const int cnt = sh.nclaim, side = sh.side_used ? 1 : 0;
            if (!have_excl) {  // multi-pass buckets publish only their inclusive value, at the end
                if (warp == 0) {
                    const unsigned long long e = ag_look_back(fb_state, first_fb, fb);
                    if (lane == 0) sh.excl = e;
                }
                __syncthreads();
                excl = sh.excl;
                have_excl = true;
            }
            const int64_t obase = pbase + (int64_t)(excl + written);
            for (int j = threadIdx.x; j < cnt; j += AG_THREADS) {
                const int s = s_list[j];
                const longlong2 e = s_slot[s];
                out_keys[obase + j] = key_from_bits<KeyT>(e.x);
                out_vals[obase + j] = e.y;
            }
            if (side && threadIdx.x == 0) {
                out_keys[obase + cnt] = key_from_bits<KeyT>(kEmpty);
                out_vals[obase + cnt] = sh.side_acc;
            }
            written += (unsigned long long)(cnt + side);
            __syncthreads();
        }
        // leave the table clean and publish the inclusive value
        for (int i = threadIdx.x; i < AG_CAP; i += AG_THREADS) s_slot[i] = kFree;
        if (warp == 0) {
            unsigned long long e = have_excl ? excl : ag_look_back(fb_state, first_fb, fb);
            if (lane == 0) {
                sh.overflow = 0; sh.side_used = 0; sh.side_acc = ident; sh.nclaim = 0;
                atomicExch(&fb_state[fb], AG_FLAG_INC | (e + written));
                if (fb == first_fb + fine_per_part - 1) out_counts[p] = e + written;
            }
        }
    }
}
