// smem_ops.cu -- B200 microbenchmarks behind the round-2 kernel designs.
// Measures, per SM, the cost of the shared-memory primitives the multisplit and the reduce-side merge
// can be built from: native shared atomics (add / CAS 32, 64, 128 bit), plain random LDS/STS, per-thread
// private byte counters, MATCH.ANY, and the TMA bulk copies (cp.async.bulk) in both directions for the
// run lengths a 4096-row tile produces.  Not product code: nvcc -arch=sm_100a -O3 -o smem_ops smem_ops.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int THREADS = 256;
constexpr int SLOTS = 4096;
constexpr int ITERS = 2048;

__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

enum Op { ADD32 = 0, CAS32, CAS64, CAS128, LDS32, LDS64, STS64, STS128, BYTECNT, MATCH8, ADDF64, ADD32_HALF, CAS32_HALF, EXCH32, NOPS };
static const char *kNames[] = {"atoms.add.u32 random", "atoms.cas.b32 random", "atoms.cas.b64 random", "atoms.cas.b128 random",
                               "lds.32 random", "lds.64 random", "sts.64 random", "sts.128 random",
                               "private byte counter (lds.u8+sts.u8)", "match.any 8-bit", "red.shared.add.f64 random",
                               "atoms.add.u32 16 lanes active", "atoms.cas.b32 16 lanes active", "atoms.exch.b32 random"};

template <int OP>
__global__ void __launch_bounds__(THREADS) k_op(long long *out_cycles, unsigned *sink) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t *s32 = (uint32_t *)smem;
    for (int i = threadIdx.x; i < SLOTS * 4; i += THREADS) s32[i] = 0;
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    uint32_t seed = blockIdx.x * 7919u + threadIdx.x * 104729u + 17u;
    unsigned acc = 0;
    const int lane = threadIdx.x & 31;
    const long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; it++) {
        const uint32_t h = lcg(seed) & (SLOTS - 1);
        if (OP == ADD32) {
            unsigned old;
            asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(base + h * 4) : "memory");
            acc += old;
        } else if (OP == ADD32_HALF) {
            if (lane & 1) {
                unsigned old;
                asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(base + h * 4) : "memory");
                acc += old;
            }
        } else if (OP == EXCH32) {
            unsigned old;
            asm volatile("atom.shared.exch.b32 %0, [%1], %2;" : "=r"(old) : "r"(base + h * 4), "r"(h) : "memory");
            acc += old;
        } else if (OP == CAS32) {
            unsigned old;
            asm volatile("atom.shared.cas.b32 %0, [%1], 0, %2;" : "=r"(old) : "r"(base + h * 4), "r"(h + 1) : "memory");
            acc += old;
        } else if (OP == CAS32_HALF) {
            if (lane & 1) {
                unsigned old;
                asm volatile("atom.shared.cas.b32 %0, [%1], 0, %2;" : "=r"(old) : "r"(base + h * 4), "r"(h + 1) : "memory");
                acc += old;
            }
        } else if (OP == CAS64) {
            unsigned long long old;
            asm volatile("atom.shared.cas.b64 %0, [%1], 0, %2;" : "=l"(old) : "r"(base + h * 8), "l"((unsigned long long)h + 1) : "memory");
            acc += (unsigned)old;
        } else if (OP == CAS128) {
            long long olo, ohi;
            asm volatile("{\n\t.reg .b128 c, s, o;\n\tmov.b128 c, {%3, %4};\n\tmov.b128 s, {%5, %6};\n\t"
                         "atom.shared.cas.b128 o, [%2], c, s;\n\tmov.b128 {%0, %1}, o;\n\t}"
                         : "=l"(olo), "=l"(ohi) : "r"(base + h * 16), "l"(0ll), "l"(0ll), "l"((long long)h + 1), "l"(7ll) : "memory");
            acc += (unsigned)olo;
        } else if (OP == LDS32) {
            unsigned v;
            asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + h * 4) : "memory");
            acc += v;
        } else if (OP == LDS64) {
            unsigned long long v;
            asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(base + h * 8) : "memory");
            acc += (unsigned)v;
        } else if (OP == STS64) {
            asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(base + h * 8), "l"((unsigned long long)h) : "memory");
        } else if (OP == STS128) {
            asm volatile("st.volatile.shared.v2.u64 [%0], {%1, %2};" ::"r"(base + h * 16), "l"((unsigned long long)h), "l"(1ull) : "memory");
        } else if (OP == BYTECNT) {
            // counter of bucket b (256 buckets) for this thread: row b, word = lane + 32 * (warp / 4), byte = warp % 4
            const uint32_t b = h & 255, w = threadIdx.x >> 5;
            const uint32_t a = base + b * 256 + ((lane + 32 * (w >> 2)) << 2) + (w & 3);
            unsigned c;
            asm volatile("ld.volatile.shared.u8 %0, [%1];" : "=r"(c) : "r"(a) : "memory");
            asm volatile("st.volatile.shared.u8 [%0], %1;" ::"r"(a), "r"(c + 1) : "memory");
            acc += c;
        } else if (OP == MATCH8) {
            acc += __match_any_sync(0xffffffffu, h & 255);
        } else if (OP == ADDF64) {
            asm volatile("red.shared.add.f64 [%0], %1;" ::"r"(base + h * 8), "d"(1.0) : "memory");
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345u) *sink = acc;
}

// ---- TMA bulk copies ---------------------------------------------------------------
__device__ __forceinline__ void bulk_store(void *gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// every CTA writes `rounds` tiles of 64 KB from shared memory to its own region of global memory as runs of
// `run` bytes, each run at a different (pseudo-random, 128 B aligned) place of a `span`-byte window, like the
// copy-out of a multisplit tile.  mode 0: one elected thread per warp issues the bulk stores of its share of runs;
// mode 1: all 32 lanes of every warp issue; mode 2: plain per-thread 8-byte stores (consecutive threads ->
// consecutive addresses inside a run) for comparison.
__global__ void __launch_bounds__(THREADS) k_bulk_store(unsigned char *g, size_t span, int run, int rounds, int mode, long long *out_cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int TILE = 65536;
    for (int i = threadIdx.x; i < TILE / 8; i += THREADS) ((unsigned long long *)smem)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    unsigned char *mine = g + (size_t)blockIdx.x * span;
    const int nruns = TILE / run;
    const size_t places = span / run;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t seed = blockIdx.x * 7919u + 13u;
    const long long t0 = clock64();
    for (int r = 0; r < rounds; r++) {
        if (mode == 0) {
            if (lane == 0)
                for (int i = warp; i < nruns; i += THREADS / 32) {
                    const size_t place = ((size_t)(r * 2654435761u + i * 40503u + seed)) % places;
                    bulk_store(mine + place * run, sbase + i * run, run);
                }
            if (lane == 0) { bulk_commit(); bulk_wait_read0(); }
            __syncthreads();
        } else if (mode == 1) {
            for (int i = threadIdx.x; i < nruns; i += THREADS) {
                const size_t place = ((size_t)(r * 2654435761u + i * 40503u + seed)) % places;
                bulk_store(mine + place * run, sbase + i * run, run);
            }
            bulk_commit(); bulk_wait_read0();
            __syncthreads();
        } else {
            for (int e = threadIdx.x; e < TILE / 8; e += THREADS) {
                const int i = (e * 8) / run, o = (e * 8) % run;
                const size_t place = ((size_t)(r * 2654435761u + i * 40503u + seed)) % places;
                *(unsigned long long *)(mine + place * run + o) = ((unsigned long long *)smem)[e];
            }
            __syncthreads();
        }
    }
    if (mode != 2 && lane == 0) bulk_wait0();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
}

// every CTA streams `rounds` tiles of `tile` bytes from global memory through a 2-stage mbarrier ring
__global__ void __launch_bounds__(THREADS) k_bulk_load(const unsigned char *g, size_t per_cta, int tile, long long *out_cycles, unsigned *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar[2];
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t b0 = (uint32_t)__cvta_generic_to_shared(&bar[0]);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b0 + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned char *mine = g + (size_t)blockIdx.x * per_cta;
    const int rounds = (int)(per_cta / tile);
    unsigned acc = 0;
    const long long t0 = clock64();
    auto issue = [&](int r) {
        const uint32_t bar_a = b0 + (r & 1) * 8, dst = sbase + (r & 1) * tile;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(tile) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(mine + (size_t)r * tile), "r"(tile), "r"(bar_a) : "memory");
    };
    if (threadIdx.x == 0) { issue(0); if (rounds > 1) issue(1); }
    for (int r = 0; r < rounds; r++) {
        const uint32_t bar_a = b0 + (r & 1) * 8, par = (r >> 1) & 1;
        unsigned done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar_a), "r"(par) : "memory");
        const unsigned long long *t = (const unsigned long long *)(smem + (r & 1) * tile);
        for (int i = threadIdx.x; i < tile / 8; i += THREADS) acc += (unsigned)t[i];
        __syncthreads();
        if (threadIdx.x == 0 && r + 2 < rounds) issue(r + 2);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345u) *sink = acc;
}

template <int OP>
static void run_op(int ctas_per_sm, int sms, long long *d_cyc, unsigned *d_sink) {
    const int grid = sms * ctas_per_sm;
    const size_t sh = SLOTS * 16;
    CK(cudaFuncSetAttribute(k_op<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_op<OP><<<grid, THREADS, sh>>>(d_cyc, d_sink);
    CK(cudaEventRecord(e0));
    k_op<OP><<<grid, THREADS, sh>>>(d_cyc, d_sink);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    long long *h = (long long *)malloc(grid * 8);
    CK(cudaMemcpy(h, d_cyc, grid * 8, cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < grid; i++) avg += h[i];
    avg /= grid;
    free(h);
    const double warp_instr_per_sm = (double)ITERS * (THREADS / 32) * ctas_per_sm;
    printf("%-40s ctas/sm=%d  %8.1f cycles per warp-instr per SM (clock64)   %7.2f ns/warp-instr/SM (events)\n",
           kNames[OP], ctas_per_sm, avg / ((double)ITERS * (THREADS / 32)), ms * 1e6 / warp_instr_per_sm);
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
    long long *d_cyc;
    unsigned *d_sink;
    CK(cudaMalloc(&d_cyc, 8 * 4096));
    CK(cudaMalloc(&d_sink, 4));
    for (int c : {1, 2, 3}) {
        run_op<ADD32>(c, sms, d_cyc, d_sink);
        run_op<ADD32_HALF>(c, sms, d_cyc, d_sink);
        run_op<EXCH32>(c, sms, d_cyc, d_sink);
        run_op<CAS32>(c, sms, d_cyc, d_sink);
        run_op<CAS32_HALF>(c, sms, d_cyc, d_sink);
        run_op<CAS64>(c, sms, d_cyc, d_sink);
        run_op<CAS128>(c, sms, d_cyc, d_sink);
        run_op<ADDF64>(c, sms, d_cyc, d_sink);
        run_op<LDS32>(c, sms, d_cyc, d_sink);
        run_op<LDS64>(c, sms, d_cyc, d_sink);
        run_op<STS64>(c, sms, d_cyc, d_sink);
        run_op<STS128>(c, sms, d_cyc, d_sink);
        run_op<BYTECNT>(c, sms, d_cyc, d_sink);
        run_op<MATCH8>(c, sms, d_cyc, d_sink);
    }
    // ---- bulk stores: 64 KB tile per round as runs of `run` bytes
    {
        const size_t span = 8u << 20;  // 8 MB window per CTA
        unsigned char *g;
        for (int cps : {1, 2}) {
            const int grid = sms * cps;
            CK(cudaMalloc(&g, span * grid));
            CK(cudaFuncSetAttribute(k_bulk_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
            for (int mode : {0, 1, 2})
                for (int run : {64, 128, 256, 512, 1024, 4096, 16384}) {
                    const int rounds = 64;
                    cudaEvent_t e0, e1;
                    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
                    k_bulk_store<<<grid, THREADS, 65536>>>(g, span, run, 4, mode, d_cyc);
                    CK(cudaEventRecord(e0));
                    k_bulk_store<<<grid, THREADS, 65536>>>(g, span, run, rounds, mode, d_cyc);
                    CK(cudaEventRecord(e1));
                    CK(cudaDeviceSynchronize());
                    float ms;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    const double bytes = (double)grid * rounds * 65536;
                    printf("bulk store smem->global  mode=%d (%s) ctas/sm=%d run=%5d B : %8.1f GB/s chip, %6.1f ns per run per SM\n", mode,
                           mode == 0 ? "1 lane/warp issues" : mode == 1 ? "all lanes issue" : "st.global.u64", cps, run,
                           bytes / (ms * 1e-3) / 1e9, ms * 1e6 / ((double)rounds * (65536 / run) * cps));
                }
            CK(cudaFree(g));
        }
    }
    // ---- bulk loads: stream 64 MB per CTA... (kept small: 16 MB per CTA)
    {
        for (int cps : {1, 2})
            for (int tile : {8192, 16384, 32768, 65536}) {
                if (cps * 2 * tile > 200 * 1024) continue;
                const int grid = sms * cps;
                const size_t per = 16u << 20;
                unsigned char *g;
                CK(cudaMalloc(&g, per * grid));
                CK(cudaMemset(g, 1, per * grid));
                CK(cudaFuncSetAttribute(k_bulk_load, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * tile));
                cudaEvent_t e0, e1;
                CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
                k_bulk_load<<<grid, THREADS, 2 * tile>>>(g, per, tile, d_cyc, d_sink);
                CK(cudaEventRecord(e0));
                k_bulk_load<<<grid, THREADS, 2 * tile>>>(g, per, tile, d_cyc, d_sink);
                CK(cudaEventRecord(e1));
                CK(cudaDeviceSynchronize());
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                printf("bulk load global->smem  ctas/sm=%d tile=%6d B x2 stages: %8.1f GB/s chip\n", cps, tile,
                       (double)per * grid / (ms * 1e-3) / 1e9);
                CK(cudaFree(g));
            }
    }
    return 0;
}
