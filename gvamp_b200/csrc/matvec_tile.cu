// matvec_tile.cu -- the main kernels of the bed mat-vecs for sm_100a (tile walks) with their pre- and post-kernels.
//
// Both products (reference data::Ax data.cpp:848-1011 and dot_product/ATx data.cpp:728-835) run on ONE
// skeleton.  The unit of data movement is a TILE = 32 marker groups x 1 stripe = 32 rows of 128 B =
// 4 KB, CONTIGUOUS in the striped layout (gvb_internal.cuh), fetched by one TMA bulk copy
// (cp.async.bulk -> mbarrier) issued by lane 0 of the warp that consumes it.  A warp reads its tile
// along XOR-skewed diagonals -- at step tau lane l reads word (row, col) with
//
//      X.v   : row = tau ^ l (marker group), col = l       (lane = byte position, 4 individuals)
//      X^T.u : row = l (marker group),       col = tau ^ l (lane = marker group, 4 markers)
//
// so the 32 lanes always (a) hit 32 different banks of the bed tile and (b) need 32 different slots
// of the 256-entry lookup table of the step (slot = bank, tables are [256 entries][32 slots] int32):
// every shared-memory access of the kernel is conflict free, the accumulators never move between
// lanes (4 per lane, no cross-lane reduction at all), and the staging of the bed costs one 4 KB TMA
// instead of 32 four-byte cp.async per lane (what limited the gen-1 X.v kernel: lg_throttle).
//
//   X.v   table H_T[e][s]  = sum_q rint(val_{4g+q}(c_q(e)) * scale), g = 32 T + s   (ax_build_kernel)
//         index e of individual k = ((w & 0x03030303<<2k) * (0x01041040>>2k)) >> 24   LOP3 + IMAD
//   X^T.u table T_t[B][c]  = sum_k a(code_k(B)) * U_{4(32t+c)+k}                       (atx_build_kernel)
//         index B = byte q of the word (marker 4g+q)                                    (nothing)
//   address = e*256 + buffer*128 + slot*4: one PRMT merges the index byte with a loop-invariant
//   per-lane slot byte (8 registers hold the 32 slot bytes); the table base is a link-time constant
//   folded into the LDS immediate.  Per 4 genotypes: [LOP3 + IMAD +] PRMT + LDS + IADD/IMAD.
//
// A CTA = NW consumer warps + one producer warp.  The consumers walk their tiles step by step and share the table tile of the step
// (X.v: warps = NW stripes, same marker tile; X^T.u: warps = NW marker tiles, same stripe).  Two ways to stage the tables:
//   * pair mode (ax_pair_kernel / atx_pair_kernel, the default for X^T.u and for X.v on the twin): the tables of two consecutive steps
//     are interleaved in global memory the way they are addressed in shared memory, one 64 KB region = one pair = ONE cp.async.bulk issued
//     by one elected thread; two regions double-buffer pairs of steps; 12 consumer warps x 2 bed stages fill the 227 KB (PairLayout);
//   * producer warp (ax_tile_kernel / atx_tile_kernel; the gathering X.v kernel, which needs its 15 warps): 16-byte cp.async into a
//     two-step double buffer interleaved at 128 B.
// Producer and consumers meet on two pairs of mbarriers (full / empty per buffer or region) instead of a __syncthreads per step, so a
// warp whose bed tile arrives late does not stall the others.  Bed tiles are per-warp, NS stages deep.  Persistent CTAs pull
// rectangular work items from an atomic counter; items are ordered so that CTAs running at the same time share their table tiles in L2.
//
// Arithmetic: fixed point, exact after quantisation, with one power-of-two scale class per stripe / marker tile (see "Scale classes"
// below; pre- and post-kernels at the end of this file).
#include "gvb_internal.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier; the bed is read once per
// sweep, so it is tagged evict-first and leaves the L2 to the lookup tables
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the stage was just read through the generic proxy
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
// arrive on an mbarrier once all cp.async copies this thread has issued so far have landed (the barrier's count includes
// these arrivals: .noinc)
__device__ __forceinline__ void cp_async_arrive_on(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned sel) {
    unsigned d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
// acc + val * one with a run-time `one` == 1: an IMAD (fma pipe) instead of an IADD3 (alu pipe); the alu pipe
// already carries the LOP3/PRMT of every lookup (B300_MICROARCH: both pipes issue one warp per 2 cycles)
__device__ __forceinline__ int mad_one(int val, int one, int acc) {
    int d;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(val), "r"(one), "r"(acc));
    return d;
}

constexpr int TAB_REGION = 65536;   // [256 entries][2 buffers][32 slots] int32: entry stride 256 B
constexpr int TILE_BYTES = 4096;    // 32 groups x 128 B
constexpr int TILE_WORDS = 1024;

template <int NW, int NS>
struct TileCfg {
    static constexpr int THREADS = NW * 32 + 32;   // NW consumer warps + one producer warp that stages the lookup tables
    static constexpr int SMEM = TAB_REGION + NW * NS * TILE_BYTES + TILE_BYTES;   // + slack for the 4 KB alignment of the bed stages
};

// The table double buffer is a two-stage producer / consumer pipeline on mbarriers instead of a __syncthreads per step: the
// producer warp refills buffer b as soon as all NW consumer warps have left it (tab_empty[b]) and a consumer warp starts a step
// as soon as ITS bed tile and the step's table have landed (tab_full[b]), so a warp whose tile arrives late no longer stalls the
// others.  Measured on B200 (27.5 GB shard): 15 consumers + producer 5.87 / 5.93 TB/s (X.v / X^T.u) against 5.79 / 5.76 TB/s for 16
// warps with a __syncthreads per step.  (Dropping the per-step barrier altogether - unsafe, timing only - gave 6.37 TB/s, so the
// warps still pay for moving within one step of each other; a third table buffer would not fit the 227 KB.  Also measured and
// not kept: 16 + 1 warps (the register file then allows 96 registers instead of 128: no gain), bed tiles staged through registers
// with two TMA copies in flight per warp (slower), an L2 prefetch of the bed tiles 1-8 steps ahead (no gain), table rows as 256
// TMA bulk copies of 128 B (4.7x slower: the copy engine is the limit).)
struct TabPipe {
    uint32_t full, empty;   // shared-memory addresses of tab_full[2] / tab_empty[2] (8 bytes apart)
    uint32_t fills0, fills1;   // how often buffer 0 / 1 has been filled since kernel start (uniform over the CTA)
    template <int B>
    __device__ __forceinline__ uint32_t& fills() { return B ? fills1 : fills0; }
};

// slot bytes of the XOR-skewed walk, 4 per register: byte (tau & 3) of spack[tau >> 2] = ((tau ^ lane) & 31) * 4
__device__ __forceinline__ void make_spack(unsigned (&spack)[8], int lane) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
        unsigned x = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) x |= ((unsigned)((((4 * j + b) ^ lane) & 31) * 4)) << (8 * b);
        spack[j] = x;
    }
}

// producer warp: table tile of one step -> interleaved buffer `buf` of the table region, 64 x 16 bytes per lane.  In global memory the
// tiles of two consecutive steps are interleaved the same way ([pair][256 entries][2 steps][32 slots], gvb_tab_index): `src` points
// at the step's half of entry 0, entries are 64 ints apart
__device__ __forceinline__ void stage_table(uint32_t tab_sm, int buf, const int* __restrict__ src, int lane) {
    const uint32_t dst = tab_sm + buf * 128;
#pragma unroll 8
    for (int c = lane; c < 2048; c += 32) cp_async16(dst + (c >> 3) * 256 + (c & 7) * 16, src + (c >> 3) * 64 + (c & 7) * 4);
}
__device__ __forceinline__ void tab_pipe_init(TabPipe& tp, unsigned long long* s_tab, int consumers) {
    tp.full = smem_u32(&s_tab[0]);
    tp.empty = smem_u32(&s_tab[2]);
    tp.fills0 = tp.fills1 = 0;
    if (threadIdx.x == 0) {
        mbar_init(tp.full, 32);
        mbar_init(tp.full + 8, 32);
        mbar_init(tp.empty, consumers);
        mbar_init(tp.empty + 8, consumers);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
}
// producer: wait until the consumers have left buffer b, refill it, let the copies arrive on tab_full[b]
template <int B>
__device__ __forceinline__ void tab_produce(TabPipe& tp, uint32_t tab_sm, const int* __restrict__ src, int lane) {
    mbar_wait(tp.empty + 8 * B, (tp.fills<B>() & 1) ^ 1);   // a fresh barrier passes the wait for the phase "before the first"
    stage_table(tab_sm, B, src, lane);
    cp_async_arrive_on(tp.full + 8 * B);
    tp.fills<B>()++;
}
// producer warp: the tables of the n steps of a work item (it starts on a pair boundary), buffers alternating from 0
__device__ __forceinline__ void tab_produce_item(TabPipe& tp, uint32_t tab_sm, const int* __restrict__ src, int n, int lane) {
    for (int i = 0; i < n; i += 2) {
        tab_produce<0>(tp, tab_sm, src + (long)i * 8192, lane);               // pair i/2, half 0
        if (i + 1 < n) tab_produce<1>(tp, tab_sm, src + (long)i * 8192 + 32, lane);   // pair i/2, half 1
    }
}

// work_counter[0] hands out items, work_counter[1] counts the CTAs that ran dry: the last one re-arms both for the
// next launch on the stream (no memset node between the sweeps)
__device__ __forceinline__ void release_work_counter(int* work_counter) {
    if (threadIdx.x == 0) {
        const int done = atomicAdd(work_counter + 1, 1);
        if (done == (int)gridDim.x - 1) {
            work_counter[0] = 0;
            work_counter[1] = 0;
            __threadfence();
        }
    }
}

// ===================================================================================================
//  X . v : warp = stripe, lane = byte position (4 individuals), walk over the marker tiles of a chunk
// ===================================================================================================
// MADK of the four accumulators take their additions as IMAD (fma pipe), the others as IADD3 (alu pipe, two lookups per
// instruction): the alu pipe also carries the LOP3 / PRMT of every lookup and is the busiest unit of this kernel
// TW: the words come from the individual-major twin (twin.cu): byte k IS the index of individual k, no gather
template <int MADK, bool TW>
__device__ __forceinline__ void ax_consume(const char* __restrict__ tb, uint32_t bb, const unsigned (&spack)[8], int one, int (&a32)[4]) {
    // tb: the step's table inside the table region -- a compile-time offset in the two-buffer form, a warp-uniform register in pair mode
#pragma unroll
    for (int tau = 0; tau < 32; tau++) {
        const uint32_t w = lds32(bb ^ (uint32_t)(tau << 7));   // row tau ^ lane, column lane
#pragma unroll
        for (int k = 0; k < 4; k++) {
            // the 2-bit codes of individual k in the four marker bytes -> one 8-bit index in the top byte
            const unsigned prod = TW ? w : (w & (0x03030303u << (2 * k))) * (0x01041040u >> (2 * k));
            // address = index * 256 + slot * 4 (bytes 2,3 = sign replication of a slot byte < 128 = 0)
            const unsigned a = prmt(prod, spack[tau >> 2], (TW ? (0xCC00u | (k << 4)) : 0xCC30u) | (4 + (tau & 3)));
            const int val = *reinterpret_cast<const int*>(tb + a);
            a32[k] = (k >= 4 - MADK) ? mad_one(val, one, a32[k]) : a32[k] + val;
        }
    }
}

template <int NW, int NS, int MADK, bool TW>
__global__ void __launch_bounds__(NW * 32 + 32, 1)
ax_tile_kernel(const uint32_t* __restrict__ bed, const int* __restrict__ tabv, long Mg_pad, long stripe0, long n_stripes, int n_sblocks, int n_gchunks,
               int tiles_per_chunk, int* __restrict__ work_counter, unsigned long long* __restrict__ acc_out, int one, const int* __restrict__ skip,
               const int* __restrict__ shifts) {
    if (skip && *skip) return;   // device-side predicate of a speculatively enqueued CG iteration (cg.cu): uniform over the grid
    extern __shared__ __align__(1024) char smem[];
    __shared__ int s_item;
    __shared__ __align__(8) unsigned long long s_bar[NW * NS];
    __shared__ __align__(8) unsigned long long s_tab[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool producer = warp == NW;
    const uint32_t tab_sm = smem_u32(smem);
    const uint32_t bed_sm = ((tab_sm + TAB_REGION + 4095u) & ~4095u) + (producer ? 0 : warp) * (NS * TILE_BYTES);   // 4 KB aligned: XOR addressing
    const uint32_t bar0 = smem_u32(&s_bar[(producer ? 0 : warp) * NS]);
    if (lane == 0 && !producer) {
#pragma unroll
        for (int s = 0; s < NS; s++) mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    TabPipe tp;
    tab_pipe_init(tp, s_tab, NW);
    const uint64_t pol = policy_evict_first();
    unsigned spack[8];
    make_spack(spack, lane);
    const uint32_t lane_off = lane * 132;   // row `lane`, column `lane`
    uint32_t n_fill = 0, n_use = 0;         // tiles requested / consumed by this warp since kernel start (stage = n % NS)
    const int n_items = n_sblocks * n_gchunks;
    const long n_tiles = Mg_pad / 32;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(work_counter, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= n_items) break;
        // marker-chunk major order: CTAs running at the same time share the chunk's tables in L2
        const int gc = item / n_sblocks, sb = item % n_sblocks;
        const long tile_lo = (long)gc * tiles_per_chunk;
        const int nt = (int)min((long)tiles_per_chunk, n_tiles - tile_lo);
        const int* tsrc = tabv + tile_lo * 8192;

        if (producer) {
            tab_produce_item(tp, tab_sm, tsrc, nt, lane);
            continue;
        }
        // stripes [stripe0, stripe0 + n_stripes) of the matrix behind `bed` (a partial twin covers a prefix of the stripes, the one
        // matrix the rest: two launches with different TW)
        const long t = stripe0 + (long)sb * NW + warp;
        const bool active = (long)sb * NW + warp < n_stripes;
        const uint32_t* bsrc = bed + ((active ? t : stripe0) * Mg_pad + tile_lo * 32) * 32;
        auto issue_bed = [&](int i) {
            if (active && i < nt) {
                if (lane == 0) {
                    const uint32_t s = n_fill % NS;
                    mbar_expect_tx(bar0 + 8 * s, TILE_BYTES);
                    bulk_g2s(bed_sm + s * TILE_BYTES, bsrc + (long)i * TILE_WORDS, TILE_BYTES, bar0 + 8 * s, pol);
                }
                n_fill++;
            }
        };
#pragma unroll
        for (int s = 0; s < NS - 1; s++) issue_bed(s);
        long long acc64[4] = {0, 0, 0, 0};

        for (int i0 = 0; i0 < nt; i0 += 2) {
#define AX_STEP(B)                                                                           \
    if (i0 + B < nt) {                                                                       \
        const int i = i0 + B;                                                                \
        issue_bed(i + NS - 1); /* its stage was consumed in step i-1 by this same warp */    \
        const int sh = __ldg(shifts + tile_lo + i); /* scale class of this table tile */     \
        mbar_wait(tp.full + 8 * B, tp.fills<B>() & 1); /* table i has landed */              \
        tp.fills<B>()++;                                                                     \
        if (active) {                                                                        \
            const uint32_t s = n_use % NS;                                                   \
            mbar_wait(bar0 + 8 * s, (n_use / NS) & 1);                                       \
            n_use++;                                                                         \
            int a32[4] = {0, 0, 0, 0};                                                       \
            ax_consume<MADK, TW>(smem + B * 128, bed_sm + s * TILE_BYTES + lane_off, spack, one, a32); \
            _Pragma("unroll") for (int k = 0; k < 4; k++) acc64[k] += (long long)a32[k] << sh; \
        }                                                                                    \
        __syncwarp();                                                                        \
        if (lane == 0) mbar_arrive(tp.empty + 8 * B); /* this warp has left table buffer B */ \
    }
            AX_STEP(0)
            AX_STEP(1)
#undef AX_STEP
        }
        if (active) {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (acc64[k] != 0) atomicAdd(acc_out + (t * 32 + lane) * 4 + k, (unsigned long long)acc64[k]);
        }
    }
    release_work_counter(work_counter);
}

// ===================================================================================================
//  X^T . u : warp = marker tile (32 groups), lane = marker group (4 markers), walk over the stripes of a chunk
// ===================================================================================================
template <bool USE_MAD>
__device__ __forceinline__ void atx_consume(const char* __restrict__ tb, uint32_t bb, const unsigned (&spack)[8], int one, int (&a32)[4]) {
#pragma unroll
    for (int tau = 0; tau < 32; tau++) {
        const uint32_t w = lds32(bb ^ (uint32_t)(tau << 2));   // row lane, column tau ^ lane
#pragma unroll
        for (int q = 0; q < 4; q++) {
            // address = byte_q(w) * 256 + slot * 4
            const unsigned a = prmt(w, spack[tau >> 2], 0xCC00u | (q << 4) | (4 + (tau & 3)));
            const int val = *reinterpret_cast<const int*>(tb + a);
            a32[q] = USE_MAD ? mad_one(val, one, a32[q]) : a32[q] + val;
        }
    }
}

// MODE 0: plain sums (X^T.u).  MODE 1: every table entry packs three 10-bit counters (marker statistics, stats.cu); a
// stripe adds at most 32 x 4 = 128 to each of them, and the per-stripe flush widens them to three 21-bit counters of
// the int64 accumulator.
template <int NW, int NS, bool USE_MAD, int MODE>
__global__ void __launch_bounds__(NW * 32 + 32, 1)
atx_tile_kernel(const uint32_t* __restrict__ bed, const int* __restrict__ tab, long Mg_pad, long n_stripes, int n_gblocks, int n_schunks,
                int stripes_per_chunk, int* __restrict__ work_counter, unsigned long long* __restrict__ acc_out, int one, const int* __restrict__ skip,
                const int* __restrict__ shifts) {
    if (skip && *skip) return;
    extern __shared__ __align__(1024) char smem[];
    __shared__ int s_item;
    __shared__ __align__(8) unsigned long long s_bar[NW * NS];
    __shared__ __align__(8) unsigned long long s_tab[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool producer = warp == NW;
    const uint32_t tab_sm = smem_u32(smem);
    const uint32_t bed_sm = ((tab_sm + TAB_REGION + 4095u) & ~4095u) + (producer ? 0 : warp) * (NS * TILE_BYTES);
    const uint32_t bar0 = smem_u32(&s_bar[(producer ? 0 : warp) * NS]);
    if (lane == 0 && !producer) {
#pragma unroll
        for (int s = 0; s < NS; s++) mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    TabPipe tp;
    tab_pipe_init(tp, s_tab, NW);
    const uint64_t pol = policy_evict_first();
    unsigned spack[8];
    make_spack(spack, lane);
    const uint32_t lane_off = lane * 132;
    uint32_t n_fill = 0, n_use = 0;
    const int n_items = n_gblocks * n_schunks;
    const long n_tiles = Mg_pad / 32;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(work_counter, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= n_items) break;
        // stripe-chunk major order: CTAs running at the same time share the chunk's tables in L2
        const int sc = item / n_gblocks, gb = item % n_gblocks;
        const long t_lo = (long)sc * stripes_per_chunk;
        const int ns = (int)min((long)stripes_per_chunk, n_stripes - t_lo);
        const int* tsrc = tab + t_lo * 8192;

        if (producer) {
            tab_produce_item(tp, tab_sm, tsrc, ns, lane);
            continue;
        }
        const long T = (long)gb * NW + warp;
        const bool active = T < n_tiles;
        const uint32_t* bsrc = bed + (t_lo * Mg_pad + (active ? T : 0) * 32) * 32;   // + i * Mg_pad * 32 words per stripe
        auto issue_bed = [&](int i) {
            if (active && i < ns) {
                if (lane == 0) {
                    const uint32_t s = n_fill % NS;
                    mbar_expect_tx(bar0 + 8 * s, TILE_BYTES);
                    bulk_g2s(bed_sm + s * TILE_BYTES, bsrc + (long)i * Mg_pad * 32, TILE_BYTES, bar0 + 8 * s, pol);
                }
                n_fill++;
            }
        };
#pragma unroll
        for (int s = 0; s < NS - 1; s++) issue_bed(s);
        long long acc64[4] = {0, 0, 0, 0};

        for (int i0 = 0; i0 < ns; i0 += 2) {
#define ATX_STEP(B)                                                                            \
    if (i0 + B < ns) {                                                                         \
        const int i = i0 + B;                                                                  \
        issue_bed(i + NS - 1);                                                                 \
        const int sh = (MODE == 0 && shifts) ? __ldg(shifts + t_lo + i) : 0; /* scale class of this stripe's table */ \
        mbar_wait(tp.full + 8 * B, tp.fills<B>() & 1);                                         \
        tp.fills<B>()++;                                                                       \
        if (active) {                                                                          \
            const uint32_t s = n_use % NS;                                                     \
            mbar_wait(bar0 + 8 * s, (n_use / NS) & 1);                                         \
            n_use++;                                                                           \
            int a32[4] = {0, 0, 0, 0};                                                         \
            atx_consume<USE_MAD>(smem + B * 128, bed_sm + s * TILE_BYTES + lane_off, spack, one, a32); \
            _Pragma("unroll") for (int q = 0; q < 4; q++) {                                    \
                if (MODE == 0) {                                                               \
                    acc64[q] += (long long)a32[q] << sh;                                       \
                } else {                                                                       \
                    const unsigned a = (unsigned)a32[q];                                       \
                    acc64[q] += (long long)((unsigned long long)(a & 0x3FFu) | ((unsigned long long)((a >> 10) & 0x3FFu) << 21) | \
                                            ((unsigned long long)(a >> 20) << 42));            \
                }                                                                              \
            }                                                                                  \
        }                                                                                      \
        __syncwarp();                                                                          \
        if (lane == 0) mbar_arrive(tp.empty + 8 * B);                                          \
    }
            ATX_STEP(0)
            ATX_STEP(1)
#undef ATX_STEP
        }
        if (active) {
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (acc64[q] != 0) atomicAdd(acc_out + (T * 32 + lane) * 4 + q, (unsigned long long)acc64[q]);
        }
    }
    release_work_counter(work_counter);
}


// ===================================================================================================
//  Table staging by TMA ("pair" mode).  The producer warp of the kernels above moves every 32 KB table tile with 2048 cp.async of 16
//  bytes: 256 shared-memory store wavefronts per step on the load/store pipe that is the limiter of both walks (measured: skipping the
//  staging is worth 3.7 %), and two table buffers keep the consumer warps within one step of each other (dropping the barrier: +8 %).
//  Here the tables of two consecutive steps are INTERLEAVED IN GLOBAL MEMORY the way the consumers address them in shared memory
//  ([pair][256 entries][2 steps][32 slots], gvb_tab_index): one table region (64 KB = two steps) is a contiguous run that ONE
//  cp.async.bulk fills -- no load/store-pipe traffic, one elected thread instead of a warp -- and two regions double-buffer PAIRS of
//  steps, so a warp may run up to three steps ahead of the slowest one.  Shared memory: 128 KB of tables + NW x NS x 4 KB of bed tiles.
// ===================================================================================================
constexpr int PAIR_REGION = 65536;   // [256 entries][2 steps][32 slots] int32
#ifdef GVB_EXP_TAB_COPY_BYTES            // timing experiment only (wrong results): what a smaller L2 -> shared-memory table stream would buy
constexpr int PAIR_COPY_BYTES = GVB_EXP_TAB_COPY_BYTES;
#else
constexpr int PAIR_COPY_BYTES = PAIR_REGION;
#endif

// Shared memory of a pair-mode CTA, all of it dynamic so that 128 KB of tables + 24 bed tiles (12 consumer warps x 2) fit the 227 KB:
//   [2 table regions at the start of the segment: a link-time constant base, so the lookups are LDS [R + UR]]
//   [pad up to the next 4 KB boundary][NW x NS bed tiles, 4 KB aligned: XOR addressing][control block]
// the 256-byte control block (work item, table barriers, bed barriers) sits in the pad when the pad is large enough (the dynamic
// segment starts 1 KB into the window: 3 KB of pad), behind the bed tiles otherwise.
constexpr int PAIR_CTRL = 256;
constexpr int SMEM_MAX = 232448;   // 227 KB
template <int NW, int NS>
struct PairCfg {
    static_assert(NW * NS <= 24, "control block holds 24 bed barriers");
    static constexpr int THREADS = NW * 32 + 32;
    static constexpr int WANT = 2 * PAIR_REGION + NW * NS * TILE_BYTES + TILE_BYTES + PAIR_CTRL;
    static constexpr int SMEM = WANT < SMEM_MAX ? WANT : SMEM_MAX;
};
struct PairLayout {
    uint32_t bed, tab, ctrl;   // shared-memory addresses
    uint32_t tab_off, ctrl_off;   // the same as offsets from the dynamic segment (generic pointers)
};
template <int NW, int NS>
__device__ __forceinline__ PairLayout pair_layout(const char* smem) {
    PairLayout L;
    const uint32_t base = smem_u32(smem);
    L.tab = base;
    const uint32_t tab_end = base + 2 * PAIR_REGION;
    L.bed = (tab_end + 4095u) & ~4095u;
    L.ctrl = (L.bed - tab_end >= (uint32_t)PAIR_CTRL) ? tab_end : L.bed + NW * NS * TILE_BYTES;
    if (L.ctrl + PAIR_CTRL > base + PairCfg<NW, NS>::SMEM || L.bed + NW * NS * TILE_BYTES > base + PairCfg<NW, NS>::SMEM) __trap();   // layout does not fit
    L.tab_off = 0;
    L.ctrl_off = L.ctrl - base;
    return L;
}

__device__ __forceinline__ void bulk_g2s_plain(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

struct PairPipe {
    uint32_t full, empty;      // shared-memory addresses of full[2] / empty[2]
    uint32_t fills0, fills1;   // how often region 0 / 1 has been filled since kernel start (uniform over the CTA)
    template <int R>
    __device__ __forceinline__ uint32_t& fills() { return R ? fills1 : fills0; }
};
__device__ __forceinline__ void pair_pipe_init(PairPipe& pp, uint32_t s_tab, int consumers) {
    pp.full = s_tab;
    pp.empty = s_tab + 16;
    pp.fills0 = pp.fills1 = 0;
    if (threadIdx.x == 0) {
        mbar_init(pp.full, 1);
        mbar_init(pp.full + 8, 1);
        mbar_init(pp.empty, consumers);
        mbar_init(pp.empty + 8, consumers);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
}
// producer (one lane): the table pairs of a work item of n steps, regions alternating from 0; src = first pair of the item
__device__ __forceinline__ void pair_produce_item(PairPipe& pp, uint32_t tab_sm, const int* __restrict__ src, int n, int lane) {
    const int npairs = (n + 1) >> 1;
    for (int p = 0; p < npairs; p++) {
        const int r = p & 1;
        uint32_t& f = r ? pp.fills1 : pp.fills0;
        if (lane == 0) {
            mbar_wait(pp.empty + 8 * r, (f & 1) ^ 1);   // a fresh barrier passes the wait for the phase "before the first"
            mbar_expect_tx(pp.full + 8 * r, PAIR_COPY_BYTES);
            bulk_g2s_plain(tab_sm + r * PAIR_REGION, src + (long)p * (PAIR_REGION / 4), PAIR_COPY_BYTES, pp.full + 8 * r);
        }
        f++;
    }
}


// ---- pair mode addressing.  address = table region base (link-time constant, LDS immediate) + rgn * 65536 + entry * 256 + half * 128
// + slot * 4, produced by ONE PRMT per lookup like in the two-buffer form: the per-lane slot registers hold three slot bytes and, in
// byte 3, the region number; (half * 128) rides in bit 7 of every slot byte.  PRMT picks byte 0 = slot byte, byte 1 = index byte,
// byte 2 = region byte, byte 3 = sign replication of the region byte (= 0).  One loop body serves every (region, half): after each step
// the eleven slot registers are XORed with a constant (toggle bit 7 of the slot bytes; after an odd step also the region byte).
__device__ __forceinline__ void make_spack3(unsigned (&sp)[11], int lane) {
#pragma unroll
    for (int j = 0; j < 11; j++) {
        unsigned x = 0;
#pragma unroll
        for (int b = 0; b < 3; b++)
            if (3 * j + b < 32) x |= ((unsigned)((((3 * j + b) ^ lane) & 31) * 4)) << (8 * b);
        sp[j] = x;
    }
}

template <int MADK, bool TW>
__device__ __forceinline__ void ax_consume_p(const char* __restrict__ tabc, uint32_t bb, const unsigned (&sp)[11], int one, int (&a32)[4]) {
#pragma unroll
    for (int tau = 0; tau < 32; tau++) {
        const uint32_t w = lds32(bb ^ (uint32_t)(tau << 7));   // row tau ^ lane, column lane
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const unsigned prod = TW ? w : (w & (0x03030303u << (2 * k))) * (0x01041040u >> (2 * k));
            const unsigned a = prmt(prod, sp[tau / 3], (TW ? (0xF700u | (k << 4)) : 0xF730u) | (4 + (tau % 3)));
            const int val = *reinterpret_cast<const int*>(tabc + a);
            a32[k] = (k >= 4 - MADK) ? mad_one(val, one, a32[k]) : a32[k] + val;
        }
    }
}

template <bool USE_MAD>
__device__ __forceinline__ void atx_consume_p(const char* __restrict__ tabc, uint32_t bb, const unsigned (&sp)[11], int one, int (&a32)[4]) {
#pragma unroll
    for (int tau = 0; tau < 32; tau++) {
        const uint32_t w = lds32(bb ^ (uint32_t)(tau << 2));   // row lane, column tau ^ lane
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const unsigned a = prmt(w, sp[tau / 3], 0xF700u | (q << 4) | (4 + (tau % 3)));
            const int val = *reinterpret_cast<const int*>(tabc + a);
            a32[q] = USE_MAD ? mad_one(val, one, a32[q]) : a32[q] + val;
        }
    }
}

// one step of a consumer warp in pair mode: step i of the item lives in region (i >> 1) & 1, half i & 1 (both carried by the slot
// registers, see above: the loop body is the same code for every step, ~1200 instructions = 19 KB, inside the 32 KB L1.5 I-cache)
#define PAIR_STEP(N_STEPS, CONSUME, FLUSH)                                                       \
    {                                                                                            \
        const int rgn = (i >> 1) & 1;                                                            \
        issue_bed(i + NS - 1);                                                                   \
        if ((i & 1) == 0) {                                                                      \
            const uint32_t f = rgn ? pp.fills1 : pp.fills0;                                      \
            mbar_wait(pp.full + 8 * rgn, f & 1);                                                 \
            if (rgn) pp.fills1++; else pp.fills0++;                                              \
        }                                                                                        \
        const int sh = shifts ? __ldg(shifts + step_lo + i) : 0;                                 \
        if (active) {                                                                            \
            const uint32_t s = n_use % NS;                                                       \
            mbar_wait(bar0 + 8 * s, (n_use / NS) & 1);                                           \
            n_use++;                                                                             \
            int a32[4] = {0, 0, 0, 0};                                                           \
            CONSUME;                                                                             \
            FLUSH;                                                                               \
        }                                                                                        \
        {                                                                                        \
            const unsigned flip = (i & 1) ? 0x01808080u : 0x00808080u;                           \
            _Pragma("unroll") for (int j = 0; j < 11; j++) sp[j] ^= flip;                        \
        }                                                                                        \
        if ((i & 1) || i == N_STEPS - 1) {                                                       \
            __syncwarp();                                                                        \
            if (lane == 0) mbar_arrive(pp.empty + 8 * rgn);                                      \
        }                                                                                        \
    }

template <int NW, int NS, int MADK, bool TW>
__global__ void __launch_bounds__(NW * 32 + 32, 1)
ax_pair_kernel(const uint32_t* __restrict__ bed, const int* __restrict__ tabv, long Mg_pad, long stripe0, long n_stripes, int n_sblocks, int n_gchunks,
               int tiles_per_chunk, int* __restrict__ work_counter, unsigned long long* __restrict__ acc_out, int one, const int* __restrict__ skip,
               const int* __restrict__ shifts) {
    if (skip && *skip) return;
    extern __shared__ __align__(1024) char smem[];
    const PairLayout lay = pair_layout<NW, NS>(smem);
    volatile int* s_item_p = reinterpret_cast<volatile int*>(smem + lay.ctrl_off);       // control block: work item at +0,
    const char* tabs = smem;                                                                // table barriers at +8, bed barriers at +64
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool producer = warp == NW;
    const uint32_t tab_sm = lay.tab;
    const uint32_t bed_sm = lay.bed + (producer ? 0 : warp) * (NS * TILE_BYTES);
    const uint32_t bar0 = lay.ctrl + 64 + 8 * ((producer ? 0 : warp) * NS);
    if (lane == 0 && !producer) {
#pragma unroll
        for (int s = 0; s < NS; s++) mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    PairPipe pp;
    pair_pipe_init(pp, lay.ctrl + 8, NW);
    const uint64_t pol = policy_evict_first();
    unsigned sp[11];
    const uint32_t lane_off = lane * 132;
    uint32_t n_fill = 0, n_use = 0;
    const int n_items = n_sblocks * n_gchunks;
    const long n_tiles = Mg_pad / 32;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) *s_item_p = atomicAdd(work_counter, 1);
        __syncthreads();
        const int item = *s_item_p;
        if (item >= n_items) break;
        const int gc = item / n_sblocks, sb = item % n_sblocks;
        const long step_lo = (long)gc * tiles_per_chunk;   // even: a chunk starts on a pair boundary
        const int nt = (int)min((long)tiles_per_chunk, n_tiles - step_lo);
        if (producer) {
            pair_produce_item(pp, tab_sm, tabv + (step_lo >> 1) * (PAIR_REGION / 4), nt, lane);
            continue;
        }
        const long t = stripe0 + (long)sb * NW + warp;
        const bool active = (long)sb * NW + warp < n_stripes;
        const uint32_t* bsrc = bed + ((active ? t : stripe0) * Mg_pad + step_lo * 32) * 32;
        auto issue_bed = [&](int i) {
            if (active && i < nt) {
                if (lane == 0) {
                    const uint32_t s = n_fill % NS;
                    mbar_expect_tx(bar0 + 8 * s, TILE_BYTES);
                    bulk_g2s(bed_sm + s * TILE_BYTES, bsrc + (long)i * TILE_WORDS, TILE_BYTES, bar0 + 8 * s, pol);
                }
                n_fill++;
            }
        };
#pragma unroll
        for (int s = 0; s < NS - 1; s++) issue_bed(s);
        long long acc64[4] = {0, 0, 0, 0};
        make_spack3(sp, lane);   // region 0, half 0
#define AX_CONSUME ax_consume_p<MADK, TW>(tabs, bed_sm + s * TILE_BYTES + lane_off, sp, one, a32)
#define AX_FLUSH _Pragma("unroll") for (int k = 0; k < 4; k++) acc64[k] += (long long)a32[k] << sh
#pragma unroll 1
        for (int i = 0; i < nt; i++) PAIR_STEP(nt, AX_CONSUME, AX_FLUSH)
#undef AX_CONSUME
#undef AX_FLUSH
        if (active) {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (acc64[k] != 0) atomicAdd(acc_out + (t * 32 + lane) * 4 + k, (unsigned long long)acc64[k]);
        }
    }
    release_work_counter(work_counter);
}

template <int NW, int NS, bool USE_MAD, int MODE>
__global__ void __launch_bounds__(NW * 32 + 32, 1)
atx_pair_kernel(const uint32_t* __restrict__ bed, const int* __restrict__ tab, long Mg_pad, long n_stripes, int n_gblocks, int n_schunks,
                int stripes_per_chunk, int* __restrict__ work_counter, unsigned long long* __restrict__ acc_out, int one, const int* __restrict__ skip,
                const int* __restrict__ shifts_in) {
    if (skip && *skip) return;
    extern __shared__ __align__(1024) char smem[];
    const PairLayout lay = pair_layout<NW, NS>(smem);
    volatile int* s_item_p = reinterpret_cast<volatile int*>(smem + lay.ctrl_off);       // control block: work item at +0,
    const char* tabs = smem;                                                                // table barriers at +8, bed barriers at +64
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool producer = warp == NW;
    const uint32_t tab_sm = lay.tab;
    const uint32_t bed_sm = lay.bed + (producer ? 0 : warp) * (NS * TILE_BYTES);
    const uint32_t bar0 = lay.ctrl + 64 + 8 * ((producer ? 0 : warp) * NS);
    if (lane == 0 && !producer) {
#pragma unroll
        for (int s = 0; s < NS; s++) mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    PairPipe pp;
    pair_pipe_init(pp, lay.ctrl + 8, NW);
    const uint64_t pol = policy_evict_first();
    unsigned sp[11];
    const uint32_t lane_off = lane * 132;
    uint32_t n_fill = 0, n_use = 0;
    const int n_items = n_gblocks * n_schunks;
    const long n_tiles = Mg_pad / 32;
    const int* shifts = MODE == 0 ? shifts_in : nullptr;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) *s_item_p = atomicAdd(work_counter, 1);
        __syncthreads();
        const int item = *s_item_p;
        if (item >= n_items) break;
        const int sc = item / n_gblocks, gb = item % n_gblocks;
        const long step_lo = (long)sc * stripes_per_chunk;   // even
        const int ns = (int)min((long)stripes_per_chunk, n_stripes - step_lo);
        if (producer) {
            pair_produce_item(pp, tab_sm, tab + (step_lo >> 1) * (PAIR_REGION / 4), ns, lane);
            continue;
        }
        const long T = (long)gb * NW + warp;
        const bool active = T < n_tiles;
        const uint32_t* bsrc = bed + (step_lo * Mg_pad + (active ? T : 0) * 32) * 32;   // + i * Mg_pad * 32 words per stripe
        auto issue_bed = [&](int i) {
            if (active && i < ns) {
                if (lane == 0) {
                    const uint32_t s = n_fill % NS;
                    mbar_expect_tx(bar0 + 8 * s, TILE_BYTES);
                    bulk_g2s(bed_sm + s * TILE_BYTES, bsrc + (long)i * Mg_pad * 32, TILE_BYTES, bar0 + 8 * s, pol);
                }
                n_fill++;
            }
        };
#pragma unroll
        for (int s = 0; s < NS - 1; s++) issue_bed(s);
        long long acc64[4] = {0, 0, 0, 0};
        make_spack3(sp, lane);   // region 0, half 0
#define ATX_CONSUME atx_consume_p<USE_MAD>(tabs, bed_sm + s * TILE_BYTES + lane_off, sp, one, a32)
#define ATX_FLUSH                                                                                              \
    _Pragma("unroll") for (int q = 0; q < 4; q++) {                                                            \
        if (MODE == 0) {                                                                                       \
            acc64[q] += (long long)a32[q] << sh;                                                               \
        } else {                                                                                               \
            const unsigned a = (unsigned)a32[q];                                                               \
            acc64[q] += (long long)((unsigned long long)(a & 0x3FFu) | ((unsigned long long)((a >> 10) & 0x3FFu) << 21) | \
                                    ((unsigned long long)(a >> 20) << 42));                                    \
        }                                                                                                      \
    }
#pragma unroll 1
        for (int i = 0; i < ns; i++) PAIR_STEP(ns, ATX_CONSUME, ATX_FLUSH)
#undef ATX_CONSUME
#undef ATX_FLUSH
        if (active) {
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (acc64[q] != 0) atomicAdd(acc_out + (T * 32 + lane) * 4 + q, (unsigned long long)acc64[q]);
        }
    }
    release_work_counter(work_counter);
}
#undef PAIR_STEP

// ===================================================================================================
//  Two right-hand sides per bed read ("dual" X.v): out0 = X.v0 and out1 = X.v1 from ONE pass over the shard.  The tables of the two
//  products are interleaved entry by entry, [step][256 entries][32 slots][2 rhs] int32, so that one LDS.64 returns both summands of a
//  lookup and the index arithmetic (the gather of the 2-bit codes, the PRMT) is shared.  A 64 KB region now holds ONE step; the two
//  regions double-buffer single steps.  Every int32 window, shift and int64 sum is the one the single-product kernel forms for that
//  right-hand side: the results are bit-identical to two gvb_ax_tile calls (test_dual_ax_equals_two_sweeps).
// ===================================================================================================
__device__ __forceinline__ void make_spack_dual(unsigned (&sp)[11], int lane) {
#pragma unroll
    for (int j = 0; j < 11; j++) {
        unsigned x = 0;
#pragma unroll
        for (int b = 0; b < 3; b++)
            if (3 * j + b < 32) x |= ((unsigned)((((3 * j + b) ^ lane) & 31) * 8)) << (8 * b);
        sp[j] = x;
    }
}

template <bool TW>
__device__ __forceinline__ void ax_consume_dual(const char* __restrict__ tabc, uint32_t bb, const unsigned (&sp)[11], int (&a0)[4], int (&a1)[4]) {
#pragma unroll
    for (int tau = 0; tau < 32; tau++) {
        const uint32_t w = lds32(bb ^ (uint32_t)(tau << 7));   // row tau ^ lane, column lane
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const unsigned prod = TW ? w : (w & (0x03030303u << (2 * k))) * (0x01041040u >> (2 * k));
            const unsigned a = prmt(prod, sp[tau / 3], (TW ? (0xF700u | (k << 4)) : 0xF730u) | (4 + (tau % 3)));
            const int2 val = *reinterpret_cast<const int2*>(tabc + a);
            a0[k] += val.x;
            a1[k] += val.y;
        }
    }
}

template <int NW, int NS, bool TW>
__global__ void __launch_bounds__(NW * 32 + 32, 1)
ax_dual_kernel(const uint32_t* __restrict__ bed, const int* __restrict__ tabv2, long Mg_pad, long stripe0, long n_stripes, int n_sblocks, int n_gchunks,
               int tiles_per_chunk, int* __restrict__ work_counter, unsigned long long* __restrict__ acc_out0, unsigned long long* __restrict__ acc_out1,
               const int* __restrict__ skip, const int* __restrict__ shifts0, const int* __restrict__ shifts1) {
    if (skip && *skip) return;
    extern __shared__ __align__(1024) char smem[];
    const PairLayout lay = pair_layout<NW, NS>(smem);
    volatile int* s_item_p = reinterpret_cast<volatile int*>(smem + lay.ctrl_off);
    const char* tabs = smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool producer = warp == NW;
    const uint32_t tab_sm = lay.tab;
    const uint32_t bed_sm = lay.bed + (producer ? 0 : warp) * (NS * TILE_BYTES);
    const uint32_t bar0 = lay.ctrl + 64 + 8 * ((producer ? 0 : warp) * NS);
    if (lane == 0 && !producer) {
#pragma unroll
        for (int s = 0; s < NS; s++) mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    PairPipe pp;
    pair_pipe_init(pp, lay.ctrl + 8, NW);
    const uint64_t pol = policy_evict_first();
    unsigned sp[11];
    const uint32_t lane_off = lane * 132;
    uint32_t n_fill = 0, n_use = 0;
    const int n_items = n_sblocks * n_gchunks;
    const long n_tiles = Mg_pad / 32;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) *s_item_p = atomicAdd(work_counter, 1);
        __syncthreads();
        const int item = *s_item_p;
        if (item >= n_items) break;
        const int gc = item / n_sblocks, sb = item % n_sblocks;
        const long step_lo = (long)gc * tiles_per_chunk;
        const int nt = (int)min((long)tiles_per_chunk, n_tiles - step_lo);
        if (producer) {   // one 64 KB region per step, regions alternating from 0
            const int* src = tabv2 + step_lo * (PAIR_REGION / 4);
            for (int i = 0; i < nt; i++) {
                const int r = i & 1;
                uint32_t& f = r ? pp.fills1 : pp.fills0;
                if (lane == 0) {
                    mbar_wait(pp.empty + 8 * r, (f & 1) ^ 1);
                    mbar_expect_tx(pp.full + 8 * r, PAIR_REGION);
                    bulk_g2s_plain(tab_sm + r * PAIR_REGION, src + (long)i * (PAIR_REGION / 4), PAIR_REGION, pp.full + 8 * r);
                }
                f++;
            }
            continue;
        }
        const long t = stripe0 + (long)sb * NW + warp;
        const bool active = (long)sb * NW + warp < n_stripes;
        const uint32_t* bsrc = bed + ((active ? t : stripe0) * Mg_pad + step_lo * 32) * 32;
        auto issue_bed = [&](int i) {
            if (active && i < nt) {
                if (lane == 0) {
                    const uint32_t s = n_fill % NS;
                    mbar_expect_tx(bar0 + 8 * s, TILE_BYTES);
                    bulk_g2s(bed_sm + s * TILE_BYTES, bsrc + (long)i * TILE_WORDS, TILE_BYTES, bar0 + 8 * s, pol);
                }
                n_fill++;
            }
        };
#pragma unroll
        for (int s = 0; s < NS - 1; s++) issue_bed(s);
        long long acc0[4] = {0, 0, 0, 0}, acc1[4] = {0, 0, 0, 0};
        make_spack_dual(sp, lane);   // region 0
#pragma unroll 1
        for (int i = 0; i < nt; i++) {
            const int rgn = i & 1;
            issue_bed(i + NS - 1);
            {
                const uint32_t f = rgn ? pp.fills1 : pp.fills0;
                mbar_wait(pp.full + 8 * rgn, f & 1);
                if (rgn) pp.fills1++; else pp.fills0++;
            }
            const int sh0 = __ldg(shifts0 + step_lo + i), sh1 = __ldg(shifts1 + step_lo + i);
            if (active) {
                const uint32_t s = n_use % NS;
                mbar_wait(bar0 + 8 * s, (n_use / NS) & 1);
                n_use++;
                int a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0};
                ax_consume_dual<TW>(tabs, bed_sm + s * TILE_BYTES + lane_off, sp, a0, a1);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    acc0[k] += (long long)a0[k] << sh0;
                    acc1[k] += (long long)a1[k] << sh1;
                }
            }
#pragma unroll
            for (int j = 0; j < 11; j++) sp[j] ^= 0x01000000u;   // the other region
            __syncwarp();
            if (lane == 0) mbar_arrive(pp.empty + 8 * rgn);
        }
        if (active) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (acc0[k] != 0) atomicAdd(acc_out0 + (t * 32 + lane) * 4 + k, (unsigned long long)acc0[k]);
                if (acc1[k] != 0) atomicAdd(acc_out1 + (t * 32 + lane) * 4 + k, (unsigned long long)acc1[k]);
            }
        }
    }
    release_work_counter(work_counter);
}

// The dual X^T.u: out0 = X^T.u0 and out1 = X^T.u1 from one pass over the shard, the per-stripe tables interleaved like those of the dual
// X.v ([stripe][256 entries][32 slots][2 rhs]).  Bit-identical to two gvb_atx_tile calls on a shard without missing genotypes.
__device__ __forceinline__ void atx_consume_dual(const char* __restrict__ tabc, uint32_t bb, const unsigned (&sp)[11], int (&a0)[4], int (&a1)[4]) {
#pragma unroll
    for (int tau = 0; tau < 32; tau++) {
        const uint32_t w = lds32(bb ^ (uint32_t)(tau << 2));   // row lane, column tau ^ lane
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const unsigned a = prmt(w, sp[tau / 3], 0xF700u | (q << 4) | (4 + (tau % 3)));
            const int2 val = *reinterpret_cast<const int2*>(tabc + a);
            a0[q] += val.x;
            a1[q] += val.y;
        }
    }
}

template <int NW, int NS>
__global__ void __launch_bounds__(NW * 32 + 32, 1)
atx_dual_kernel(const uint32_t* __restrict__ bed, const int* __restrict__ tab2, long Mg_pad, long n_stripes, int n_gblocks, int n_schunks,
                int stripes_per_chunk, int* __restrict__ work_counter, unsigned long long* __restrict__ acc_out0, unsigned long long* __restrict__ acc_out1,
                const int* __restrict__ skip, const int* __restrict__ shifts0, const int* __restrict__ shifts1) {
    if (skip && *skip) return;
    extern __shared__ __align__(1024) char smem[];
    const PairLayout lay = pair_layout<NW, NS>(smem);
    volatile int* s_item_p = reinterpret_cast<volatile int*>(smem + lay.ctrl_off);
    const char* tabs = smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool producer = warp == NW;
    const uint32_t tab_sm = lay.tab;
    const uint32_t bed_sm = lay.bed + (producer ? 0 : warp) * (NS * TILE_BYTES);
    const uint32_t bar0 = lay.ctrl + 64 + 8 * ((producer ? 0 : warp) * NS);
    if (lane == 0 && !producer) {
#pragma unroll
        for (int s = 0; s < NS; s++) mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    PairPipe pp;
    pair_pipe_init(pp, lay.ctrl + 8, NW);
    const uint64_t pol = policy_evict_first();
    unsigned sp[11];
    const uint32_t lane_off = lane * 132;
    uint32_t n_fill = 0, n_use = 0;
    const int n_items = n_gblocks * n_schunks;
    const long n_tiles = Mg_pad / 32;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) *s_item_p = atomicAdd(work_counter, 1);
        __syncthreads();
        const int item = *s_item_p;
        if (item >= n_items) break;
        const int sc = item / n_gblocks, gb = item % n_gblocks;
        const long step_lo = (long)sc * stripes_per_chunk;
        const int ns = (int)min((long)stripes_per_chunk, n_stripes - step_lo);
        if (producer) {   // one 64 KB region per stripe, regions alternating from 0
            const int* src = tab2 + step_lo * (PAIR_REGION / 4);
            for (int i = 0; i < ns; i++) {
                const int r = i & 1;
                uint32_t& f = r ? pp.fills1 : pp.fills0;
                if (lane == 0) {
                    mbar_wait(pp.empty + 8 * r, (f & 1) ^ 1);
                    mbar_expect_tx(pp.full + 8 * r, PAIR_REGION);
                    bulk_g2s_plain(tab_sm + r * PAIR_REGION, src + (long)i * (PAIR_REGION / 4), PAIR_REGION, pp.full + 8 * r);
                }
                f++;
            }
            continue;
        }
        const long T = (long)gb * NW + warp;
        const bool active = T < n_tiles;
        const uint32_t* bsrc = bed + (step_lo * Mg_pad + (active ? T : 0) * 32) * 32;   // + i * Mg_pad * 32 words per stripe
        auto issue_bed = [&](int i) {
            if (active && i < ns) {
                if (lane == 0) {
                    const uint32_t s = n_fill % NS;
                    mbar_expect_tx(bar0 + 8 * s, TILE_BYTES);
                    bulk_g2s(bed_sm + s * TILE_BYTES, bsrc + (long)i * Mg_pad * 32, TILE_BYTES, bar0 + 8 * s, pol);
                }
                n_fill++;
            }
        };
#pragma unroll
        for (int s = 0; s < NS - 1; s++) issue_bed(s);
        long long acc0[4] = {0, 0, 0, 0}, acc1[4] = {0, 0, 0, 0};
        make_spack_dual(sp, lane);   // region 0
#pragma unroll 1
        for (int i = 0; i < ns; i++) {
            const int rgn = i & 1;
            issue_bed(i + NS - 1);
            {
                const uint32_t f = rgn ? pp.fills1 : pp.fills0;
                mbar_wait(pp.full + 8 * rgn, f & 1);
                if (rgn) pp.fills1++; else pp.fills0++;
            }
            const int sh0 = __ldg(shifts0 + step_lo + i), sh1 = __ldg(shifts1 + step_lo + i);
            if (active) {
                const uint32_t s = n_use % NS;
                mbar_wait(bar0 + 8 * s, (n_use / NS) & 1);
                n_use++;
                int a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0};
                atx_consume_dual(tabs, bed_sm + s * TILE_BYTES + lane_off, sp, a0, a1);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    acc0[q] += (long long)a0[q] << sh0;
                    acc1[q] += (long long)a1[q] << sh1;
                }
            }
#pragma unroll
            for (int j = 0; j < 11; j++) sp[j] ^= 0x01000000u;   // the other region
            __syncwarp();
            if (lane == 0) mbar_arrive(pp.empty + 8 * rgn);
        }
        if (active) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (acc0[q] != 0) atomicAdd(acc_out0 + (T * 32 + lane) * 4 + q, (unsigned long long)acc0[q]);
                if (acc1[q] != 0) atomicAdd(acc_out1 + (T * 32 + lane) * 4 + q, (unsigned long long)acc1[q]);
            }
        }
    }
    release_work_counter(work_counter);
}

// measured defaults (profiles/r02_sweep_tuning.txt)
#define GVB_DEFAULT_TMA_WALK 1
#define GVB_DEFAULT_TMA_GATHER 0
#define GVB_DEFAULT_PAIR_SHAPE 1

struct TileTune {
    int tma_walk;    // X^T.u and X.v on the twin: 1 = pair mode (tables by TMA bulk copy), 0 = producer-warp cp.async (two 32 KB buffers)
    int tma_gather;  // the gathering X.v kernel, likewise
    int pair_shape;  // pair mode: 0 = 11 consumer warps x 2 bed stages, 1 = 12 x 2, 2 = 7 x 3 (128 KB of tables leave room for 24 bed tiles)
    int variant;   // 0: 15 consumer warps x 2 stages, 1: 12 consumer warps x 3 stages (+ the producer warp)
    int use_mad;   // X.v: accumulators (0..4) fed by IMAD instead of IADD3; X^T.u: 4 -> all four, else none
    int ax_tiles_per_chunk, atx_stripes_per_chunk;
    int twin_mad;  // X.v on the twin: one accumulator fed by IMAD (1) or none (0)
};

TileTune tune_from_env() {
    TileTune t{GVB_DEFAULT_TMA_WALK, GVB_DEFAULT_TMA_GATHER, GVB_DEFAULT_PAIR_SHAPE, 0, 1, 0, 0, 0};   // one IMAD-fed accumulator measured best for X.v on B200 (5.30 vs 5.17 TB/s with none)   // chunk lengths 0: chosen per launch from the problem size (pick_chunk)
    if (const char* e = getenv("GVB_TAB")) t.tma_walk = t.tma_gather = strcmp(e, "cpasync") != 0;
    if (const char* e = getenv("GVB_GATHER_TAB")) t.tma_gather = strcmp(e, "cpasync") != 0;
    if (const char* e = getenv("GVB_PAIR_SHAPE")) t.pair_shape = std::max(0, std::min(2, atoi(e)));
    if (const char* e = getenv("GVB_TILE_VARIANT")) t.variant = atoi(e);
    if (const char* e = getenv("GVB_TILE_MAD")) t.use_mad = std::max(0, std::min(4, atoi(e)));
    if (const char* e = getenv("GVB_TWIN_MAD")) t.twin_mad = atoi(e);
    if (const char* e = getenv("GVB_AX_TPC")) t.ax_tiles_per_chunk = std::max(1, atoi(e));
    if (const char* e = getenv("GVB_ATX_SPC")) t.atx_stripes_per_chunk = std::max(1, atoi(e));
    return t;
}
TileTune tune() { return tune_from_env(); }

// steps (tiles of X.v / stripes of X^T.u) per work item: long enough to amortise the pipeline fill and to let the CTAs
// of a wave share table tiles in L2 (<= 64), short enough that a small matrix still yields >= 4 items per SM
int pick_chunk(int requested, long rows, long steps, int sm_count) {
    if (requested > 0) return requested + (requested & 1);   // even: a chunk starts on a pair boundary of the tables
    long per = (rows * steps + 4l * sm_count - 1) / (4l * sm_count);
    per = std::max(4l, std::min(64l, per));
    return (int)(per + (per & 1));
}   // read per launch: the tests vary the chunking within one process

template <int NW, int NS, int MADK, bool TW>
int launch_ax(gvb_ctx* c, unsigned long long* accN, long stripe0, long n_stripes) {
    using Cfg = TileCfg<NW, NS>;
    auto kern = ax_tile_kernel<NW, NS, MADK, TW>;
    static unsigned long long attr_done = 0;   // one bit per device: the attribute is per device, a process may hold several contexts
    if (!(attr_done >> (c->device & 63) & 1ull)) {
        GVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        attr_done |= 1ull << (c->device & 63);
    }
    if (n_stripes <= 0) return GVB_OK;
    const long n_tiles = c->Mg_pad / 32;
    int n_sblocks = (int)((n_stripes + NW - 1) / NW);
    const int tpc = pick_chunk(tune().ax_tiles_per_chunk, n_sblocks, n_tiles, c->sm_count);
    int n_gchunks = (int)((n_tiles + tpc - 1) / tpc);
    int grid = std::min(n_sblocks * n_gchunks, c->sm_count);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, c->stream>>>(TW ? c->bed_twin : c->bed, c->tab_v, c->Mg_pad, stripe0, n_stripes, n_sblocks, n_gchunks, tpc, c->work_counter,
                                                       accN, 1, c->skip, c->shift_v);
    GVB_LAUNCHED(c);
    return GVB_OK;
}

template <int NW, int NS, int MADK, bool TW>
int launch_ax_pair(gvb_ctx* c, unsigned long long* accN, long stripe0, long n_stripes) {
    using Cfg = PairCfg<NW, NS>;
    auto kern = ax_pair_kernel<NW, NS, MADK, TW>;
    static unsigned long long attr_done = 0;
    if (!(attr_done >> (c->device & 63) & 1ull)) {
        GVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        attr_done |= 1ull << (c->device & 63);
    }
    if (n_stripes <= 0) return GVB_OK;
    const long n_tiles = c->Mg_pad / 32;
    int n_sblocks = (int)((n_stripes + NW - 1) / NW);
    const int tpc = pick_chunk(tune().ax_tiles_per_chunk, n_sblocks, n_tiles, c->sm_count);
    int n_gchunks = (int)((n_tiles + tpc - 1) / tpc);
    int grid = std::min(n_sblocks * n_gchunks, c->sm_count);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, c->stream>>>(TW ? c->bed_twin : c->bed, c->tab_v, c->Mg_pad, stripe0, n_stripes, n_sblocks, n_gchunks, tpc, c->work_counter,
                                                       accN, 1, c->skip, c->shift_v);
    GVB_LAUNCHED(c);
    return GVB_OK;
}

template <int NW, int NS, bool TW>
int launch_ax_dual(gvb_ctx* c, unsigned long long* acc0, unsigned long long* acc1, long stripe0, long n_stripes) {
    using Cfg = PairCfg<NW, NS>;
    auto kern = ax_dual_kernel<NW, NS, TW>;
    static unsigned long long attr_done = 0;
    if (!(attr_done >> (c->device & 63) & 1ull)) {
        GVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        attr_done |= 1ull << (c->device & 63);
    }
    if (n_stripes <= 0) return GVB_OK;
    const long n_tiles = c->Mg_pad / 32;
    int n_sblocks = (int)((n_stripes + NW - 1) / NW);
    const int tpc = pick_chunk(tune().ax_tiles_per_chunk, n_sblocks, n_tiles, c->sm_count);
    int n_gchunks = (int)((n_tiles + tpc - 1) / tpc);
    int grid = std::min(n_sblocks * n_gchunks, c->sm_count);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, c->stream>>>(TW ? c->bed_twin : c->bed, c->tab_v2, c->Mg_pad, stripe0, n_stripes, n_sblocks, n_gchunks, tpc, c->work_counter,
                                                       acc0, acc1, c->skip, c->shift_v, c->shift_v2);
    GVB_LAUNCHED(c);
    return GVB_OK;
}

template <int NW, int NS>
int launch_atx_dual(gvb_ctx* c, unsigned long long* acc0, unsigned long long* acc1) {
    using Cfg = PairCfg<NW, NS>;
    auto kern = atx_dual_kernel<NW, NS>;
    static unsigned long long attr_done = 0;
    if (!(attr_done >> (c->device & 63) & 1ull)) {
        GVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        attr_done |= 1ull << (c->device & 63);
    }
    const long n_tiles = c->Mg_pad / 32;
    int n_gblocks = (int)((n_tiles + NW - 1) / NW);
    const int spc = pick_chunk(tune().atx_stripes_per_chunk, n_gblocks, c->n_stripes, c->sm_count);
    int n_schunks = (int)((c->n_stripes + spc - 1) / spc);
    int grid = std::min(n_gblocks * n_schunks, c->sm_count);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, c->stream>>>(c->bed, c->tab_u2, c->Mg_pad, c->n_stripes, n_gblocks, n_schunks, spc, c->work_counter, acc0, acc1, c->skip,
                                                       c->shift_u, c->shift_u2);
    GVB_LAUNCHED(c);
    return GVB_OK;
}

template <int NW, int NS, bool USE_MAD, int MODE>
int launch_atx_pair(gvb_ctx* c, const int* tab, unsigned long long* acc, const int* shifts) {
    using Cfg = PairCfg<NW, NS>;
    auto kern = atx_pair_kernel<NW, NS, USE_MAD, MODE>;
    static unsigned long long attr_done = 0;
    if (!(attr_done >> (c->device & 63) & 1ull)) {
        GVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        attr_done |= 1ull << (c->device & 63);
    }
    const long n_tiles = c->Mg_pad / 32;
    int n_gblocks = (int)((n_tiles + NW - 1) / NW);
    const int spc = pick_chunk(tune().atx_stripes_per_chunk, n_gblocks, c->n_stripes, c->sm_count);
    int n_schunks = (int)((c->n_stripes + spc - 1) / spc);
    int grid = std::min(n_gblocks * n_schunks, c->sm_count);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, c->stream>>>(c->bed, tab, c->Mg_pad, c->n_stripes, n_gblocks, n_schunks, spc, c->work_counter, acc, 1, c->skip, shifts);
    GVB_LAUNCHED(c);
    return GVB_OK;
}

template <int NW, int NS, bool USE_MAD, int MODE>
int launch_atx(gvb_ctx* c, const int* tab, unsigned long long* acc, const int* shifts) {
    using Cfg = TileCfg<NW, NS>;
    auto kern = atx_tile_kernel<NW, NS, USE_MAD, MODE>;
    static unsigned long long attr_done = 0;   // one bit per device: the attribute is per device, a process may hold several contexts
    if (!(attr_done >> (c->device & 63) & 1ull)) {
        GVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        attr_done |= 1ull << (c->device & 63);
    }
    const long n_tiles = c->Mg_pad / 32;
    int n_gblocks = (int)((n_tiles + NW - 1) / NW);
    const int spc = pick_chunk(tune().atx_stripes_per_chunk, n_gblocks, c->n_stripes, c->sm_count);
    int n_schunks = (int)((c->n_stripes + spc - 1) / spc);
    int grid = std::min(n_gblocks * n_schunks, c->sm_count);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, c->stream>>>(c->bed, tab, c->Mg_pad, c->n_stripes, n_gblocks, n_schunks, spc, c->work_counter, acc, 1, c->skip, shifts);
    GVB_LAUNCHED(c);
    return GVB_OK;
}


// ===================================================================================================
//  pre- and post-kernels of a sweep (fixed-point scales, lookup tables, conversion back to FP64)
// ===================================================================================================
// scal[] layout (device doubles): [0] scale of X^T.u, [1] its inverse, [2] scale of X.v, [3] its inverse,
// [8] / [9]: running maxima (bit patterns of non-negative doubles, ordered like unsigned integers) of the X^T.u / X.v
// magnitude bounds; the finish kernel of a sweep re-arms its maximum to 0 for the next sweep.
// [10] / [11]: 1.0 when the input vector of the running X^T.u / X.v sweep holds a NaN or an infinity.  The reference's FP64 sums then
// turn every output into NaN (0 * NaN in the LUT products of data.cpp:766 / :975); fixed point cannot carry a NaN, so the
// sweep carries this flag from its build kernel to its finish kernel instead.
// Scale classes.  One global power-of-two scale makes the resolution of EVERY table entry that of the largest one: a single outlier
// in u (or one tile of 128 clustered large effects in v) would cost all other entries its magnitude in precision.  The build
// kernels therefore give every stripe (X^T.u) / marker tile (X.v) its own class k in {0, 1, 2} from its OWN magnitude bound b:
// k = 2 if b <= B 2^-12, 1 if b <= B 2^-6, else 0 (B = the global bound), and quantise it with scale_0 * 2^(6k) -- the int32 window
// stays safe because b * 2^(6k) <= B.  A step of the walk covers exactly one stripe / tile, so the main kernels flush their int32
// window into the int64 accumulator shifted left by (12 - 6k): everything is summed at the finest scale scale_0 * 2^12, still exact
// integer arithmetic, at the cost of one shift per 32 lookups.  Vectors without outliers fall into class 0 throughout and give the
// bits of the single-scale form.  Head room: |window| < 2^31, shift <= 12, <= 2^16 steps per output -> |sum| < 2^59.
#define GVB_CLASS_BITS 6
#define GVB_CLASS_MAX 2
#define SCAL_ATX_BOUND 8
#define SCAL_AX_BOUND 9
#define SCAL_ATX_BAD 10
#define SCAL_AX_BAD 11

// power-of-two scale such that `window` table entries of magnitude <= bound * scale (+ slack) cannot overflow an int32
__device__ __forceinline__ double scale_for(double bound, double window, double slack) {
    double s = 1.0;
    if (bound > 0.0 && isfinite(bound)) {
        double limit = (2147483647.0 / window - slack) / bound;
        int e;
        frexp(limit, &e);   // limit = f * 2^e, f in [0.5, 1)  ->  2^(e-1) <= limit
        e = max(-1000, min(1000, e - 1));
        s = ldexp(1.0, e);
    }
    return s;
}

// class of a stripe / tile with magnitude bound b under the global bound B (see "Scale classes")
__device__ __forceinline__ int scale_class(double b, double B) {
    if (!(B > 0.0) || !isfinite(B)) return 0;
    if (b <= ldexp(B, -2 * GVB_CLASS_BITS)) return 2;
    if (b <= ldexp(B, -GVB_CLASS_BITS)) return 1;
    return 0;
}

__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
    atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

__device__ __forceinline__ int dosage_of(unsigned code) { return code == 0u ? 2 : (code == 2u ? 1 : 0); }

// ---- X . v ----------------------------------------------------------------------------------------
// pass 1 (one block per marker tile of 128): bound = max over tiles of sum_j max_c |val_j(c)|, the largest |table entry
// sum| a 32-lookup window can see; also clears the int64 accumulators of the individuals
// The three per-code values of marker j that the table walk sums over the markers (missing -> 0 in every mode):
//   mode 0: (a(c) - mu_j) sigma_j v_j          X.v, data::Ax
//   mode 1: 1                                   number of non-missing genotypes of an individual (numb_people, data.cpp:585)
//   mode 2: ((a(c) - mu_j) sigma_j)^2           sum of squared standardised genotypes of an individual (msig_people, data.cpp:586)
__device__ __forceinline__ void ax_code_values(int mode, double vj, double mu, double sg, double& v00, double& v10, double& v11) {
    if (mode == 1) {
        v00 = v10 = v11 = (sg != 0.0) ? 1.0 : 0.0;   // padded markers carry sigma = 0 and must stay out
    } else {
        const double w = (mode == 0) ? sg * vj : sg;
        v00 = (2.0 - mu) * w;
        v10 = (1.0 - mu) * w;
        v11 = -mu * w;
        if (mode == 2) { v00 *= v00; v10 *= v10; v11 *= v11; }
    }
}

__global__ void __launch_bounds__(128) ax_prep_kernel(const double* __restrict__ v, const double* __restrict__ mave, const double* __restrict__ msig,
                                                      double* __restrict__ scal, unsigned long long* __restrict__ accN, long Npad, int mode, const int* __restrict__ skip) {
    if (skip && *skip) return;
    const long j = blockIdx.x * 128l + threadIdx.x;
    double v00, v10, v11;
    ax_code_values(mode, v ? v[j] : 1.0, mave[j], msig[j], v00, v10, v11);
    double e = fmax(fmax(fabs(v00), fabs(v11)), fabs(v10));
    if (!isfinite(v00 + v10 + v11)) e = INFINITY;   // fmax drops NaNs: a non-finite input makes the bound non-finite
    __shared__ double sm[4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) atomic_max_nonneg(scal + SCAL_AX_BOUND, sm[0] + sm[1] + sm[2] + sm[3]);
    for (long i = j; i < Npad; i += (long)gridDim.x * 128) accN[i] = 0ull;
}

// pass 2 (one block per marker tile): tabv[(T*256 + e)*32 + s] = sum_q Q_{4g+q}(c_q(e)), g = 32 T + s,
// Q_j(c) = rint(val_j(c) * scale), val_j(00) = (2-mu) w, val_j(10) = (1-mu) w, val_j(11) = -mu w, val_j(missing) = 0
__global__ void __launch_bounds__(256) ax_build_kernel(const double* __restrict__ v, const double* __restrict__ mave, const double* __restrict__ msig,
                                                       double* __restrict__ scal, int* __restrict__ tabv, int mode, const int* __restrict__ skip,
                                                       int* __restrict__ shifts, int pairs, int dual_rhs) {
    if (skip && *skip) return;
    __shared__ int Qs[4][4][32];   // [q][code][slot]: a warp reads one (q, code) row -> conflict free
    __shared__ double s_scale;
    __shared__ double s_e[4];
    const long T = blockIdx.x;
    double v00 = 0.0, v10 = 0.0, v11 = 0.0;
    if (threadIdx.x < 128) {       // this tile's own bound: sum_j max_c |val_j(c)|, the same sum ax_prep_kernel maximised over the tiles
        const long j = T * 128 + threadIdx.x;
        ax_code_values(mode, v ? v[j] : 1.0, mave[j], msig[j], v00, v10, v11);
        double e = fmax(fmax(fabs(v00), fabs(v11)), fabs(v10));
        if (!isfinite(v00 + v10 + v11)) e = INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
        if ((threadIdx.x & 31) == 0) s_e[threadIdx.x >> 5] = e;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double B = scal[SCAL_AX_BOUND];
        const double sc0 = scale_for(B, 1.0, 64.0 + 1.0);
        const int k = scale_class(s_e[0] + s_e[1] + s_e[2] + s_e[3], B);
        s_scale = ldexp(sc0, GVB_CLASS_BITS * k);
        shifts[T] = GVB_CLASS_BITS * (GVB_CLASS_MAX - k);
        if (T == 0) {
            const double fine = ldexp(sc0, GVB_CLASS_BITS * GVB_CLASS_MAX);   // the scale everything is summed at
            scal[2] = fine;
            scal[3] = 1.0 / fine;
            scal[SCAL_AX_BAD] = isfinite(B) ? 0.0 : 1.0;
        }
    }
    __syncthreads();
    if (threadIdx.x < 128) {
        const int s = threadIdx.x >> 2, q = threadIdx.x & 3;
        const double sc = s_scale;
        Qs[q][0][s] = (int)rint(v00 * sc);
        Qs[q][1][s] = 0;
        Qs[q][2][s] = (int)rint(v10 * sc);
        Qs[q][3][s] = (int)rint(v11 * sc);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 4
    for (int e = warp; e < 256; e += 8)
        tabv[dual_rhs < 0 ? gvb_tab_index(T, e, lane, pairs) : (size_t)(T * 256 + e) * 64 + lane * 2 + dual_rhs] =
            Qs[0][e & 3][lane] + Qs[1][(e >> 2) & 3][lane] + Qs[2][(e >> 4) & 3][lane] + Qs[3][e >> 6][lane];
}

// out[i] = present_i ? acc_i / scale / sqrt(N) : 0   (the mask m_i of data.cpp:972; pads are never present)
__global__ void ax_finish_kernel(const unsigned long long* __restrict__ acc, double* __restrict__ scal, const uint32_t* __restrict__ maskw, long Npad,
                                 double inv_sqrt_n, double* __restrict__ out, const int* __restrict__ skip) {
    if (skip && *skip) return;
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i == 0) scal[SCAL_AX_BOUND] = 0.0;
    if (i >= Npad) return;
    bool present = (maskw[i >> 2] >> (2 * (i & 3))) & 1u;
    const double val = scal[SCAL_AX_BAD] != 0.0 ? __longlong_as_double(0x7ff8000000000000ll) : (double)(long long)acc[i] * scal[3] * inv_sqrt_n;
    out[i] = present ? val : 0.0;
}

// ---- X^T . u --------------------------------------------------------------------------------------
// pass 1: bound = max over byte positions p of 2 (|u_4p| + ... + |u_4p+3|), the largest |table entry| / scale; clears
// the accumulators of the markers and the running sum of U
__global__ void __launch_bounds__(256) atx_prep_kernel(const double* __restrict__ u, long npos, double* __restrict__ scal,
                                                       unsigned long long* __restrict__ acc, long nacc, long long* __restrict__ usum, const int* __restrict__ skip) {
    if (skip && *skip) return;
    double m = 0.0;
    const long tid = blockIdx.x * (long)blockDim.x + threadIdx.x, nth = (long)gridDim.x * blockDim.x;
    for (long p = tid; p < npos; p += nth) {
        const double2* up = reinterpret_cast<const double2*>(u + 4 * p);
        double2 a = up[0], b = up[1];
        double s4 = 2.0 * (fabs(a.x) + fabs(a.y) + fabs(b.x) + fabs(b.y));
        if (!isfinite(s4)) s4 = INFINITY;   // fmax drops NaNs: a non-finite input makes the bound non-finite
        m = fmax(m, s4);
    }
    for (long i = tid; i < nacc; i += nth) acc[i] = 0ull;
    if (tid == 0) *usum = 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0) atomic_max_nonneg(scal + SCAL_ATX_BOUND, m);
}

// pass 2 (one block per stripe): tab[(t*256 + B)*32 + l] = sum_k a(code_k(B)) U_{4p+k}, p = 32 t + l, U = rint(u * scale);
// tabm (shards with missing genotypes) holds the same sum over the MISSING codes with weight 1; accumulates sum_i U_i
__global__ void __launch_bounds__(256) atx_build_kernel(const double* __restrict__ u, double window, double* __restrict__ scal, int* __restrict__ tab,
                                                        int* __restrict__ tabm, long long* __restrict__ usum, int* __restrict__ uq, const int* __restrict__ skip,
                                                        int* __restrict__ shifts, int pairs, int dual_rhs) {
    if (skip && *skip) return;
    __shared__ int Us[4][32];   // [k][position]
    __shared__ double s_scale, s_scale0;
    __shared__ int s_shift;
    const long t = blockIdx.x;
    double ui = 0.0;
    if (threadIdx.x < 128) ui = u[t * 128 + threadIdx.x];
    if (threadIdx.x < 32) {       // this stripe's own bound: max over its 32 positions of 2 (|u_4p| + ... + |u_4p+3|), as atx_prep_kernel
        const double2* up = reinterpret_cast<const double2*>(u + t * 128 + 4 * threadIdx.x);
        const double2 a = up[0], b = up[1];
        double m = 2.0 * (fabs(a.x) + fabs(a.y) + fabs(b.x) + fabs(b.y));
        if (!isfinite(m)) m = INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) {
            const double B = scal[SCAL_ATX_BOUND];
            const double sc0 = scale_for(B, window, 4.0);
            const int k = scale_class(m, B);
            s_scale0 = sc0;
            s_scale = ldexp(sc0, GVB_CLASS_BITS * k);
            s_shift = GVB_CLASS_BITS * (GVB_CLASS_MAX - k);
            shifts[t] = s_shift;
            if (t == 0) {
                const double fine = ldexp(sc0, GVB_CLASS_BITS * GVB_CLASS_MAX);   // the scale everything is summed at
                scal[0] = fine;
                scal[1] = 1.0 / fine;
                scal[SCAL_ATX_BAD] = isfinite(B) ? 0.0 : 1.0;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 128) {
        const int Ui = (int)rint(ui * s_scale);
        Us[threadIdx.x & 3][threadIdx.x >> 2] = Ui;
        // the gather of misslist.cu sums U's of MANY stripes in one int32 window: it keeps the common class-0 scale
        if (uq) uq[t * 128 + threadIdx.x] = (int)rint(ui * s_scale0);
        long long ls = Ui;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ls += __shfl_xor_sync(0xffffffffu, ls, o);
        if ((threadIdx.x & 31) == 0 && ls != 0) atomicAdd(reinterpret_cast<unsigned long long*>(usum), (unsigned long long)(ls << s_shift));
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int U0 = Us[0][lane], U1 = Us[1][lane], U2 = Us[2][lane], U3 = Us[3][lane];
#pragma unroll 4
    for (int B = warp; B < 256; B += 8) {
        const unsigned c0 = B & 3, c1 = (B >> 2) & 3, c2 = (B >> 4) & 3, c3 = B >> 6;
        const size_t at = dual_rhs < 0 ? gvb_tab_index(t, B, lane, pairs) : (size_t)(t * 256 + B) * 64 + lane * 2 + dual_rhs;
        tab[at] = dosage_of(c0) * U0 + dosage_of(c1) * U1 + dosage_of(c2) * U2 + dosage_of(c3) * U3;
        if (tabm) tabm[at] = (c0 == 1u ? U0 : 0) + (c1 == 1u ? U1 : 0) + (c2 == 1u ? U2 : 0) + (c3 == 1u ? U3 : 0);
    }
}

// out[j] = sigma_j * (A_j - mu_j * B_j) / scale / sqrt(N),  B_j = sum_i U_i - sum_{i missing in j} U_i
__global__ void atx_finish_kernel(const unsigned long long* __restrict__ acc, const unsigned long long* __restrict__ accm, const long long* __restrict__ usum,
                                  double* __restrict__ scal, const double* __restrict__ mave, const double* __restrict__ msig, long Mpad, long M,
                                  double inv_sqrt_n, double* __restrict__ out, double* __restrict__ outB, const int* __restrict__ skip, int accm_shift) {
    if (skip && *skip) return;
    long j = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (j == 0) scal[SCAL_ATX_BOUND] = 0.0;
    if (j >= Mpad) return;
    if (scal[SCAL_ATX_BAD] != 0.0) {   // non-finite input: every marker's sum is NaN in the reference; padded markers stay 0
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        out[j] = j < M ? nan : 0.0;
        if (outB) outB[j] = j < M ? nan : 0.0;
        return;
    }
    long long A = (long long)acc[j];
    long long Bs = *usum - (accm ? ((long long)accm[j] << accm_shift) : 0ll);   // the list's gather sums at the class-0 scale
    double inv_s = scal[1];
    out[j] = msig[j] * ((double)A * inv_s - mave[j] * ((double)Bs * inv_s)) * inv_sqrt_n;
    if (outB) outB[j] = (double)Bs * inv_s;   // sum_i b_ij u_i on its own (association tests, assoc.cu)
}

int ensure_scratch(gvb_ctx* c, bool need_tab_u, bool need_tab_v, bool miss) {
    if (need_tab_u) {
        const size_t tab_ints = (size_t)gvb_roundup(c->n_stripes, 2) * 8192 * (miss ? 2 : 1);   // whole pairs of steps
        if (c->tab_u_cap < tab_ints) {
            if (c->tab_u) cudaFree(c->tab_u);
            c->tab_u = nullptr;
            GVB_CUDA(gvb_malloc(c, &c->tab_u, tab_ints * sizeof(int)));
            c->tab_u_cap = tab_ints;
        }
    }
    if (need_tab_v) {
        const size_t tab_ints = (size_t)gvb_roundup(c->Mg_pad / 32, 2) * 8192;
        if (c->tab_v_cap < tab_ints) {
            if (c->tab_v) cudaFree(c->tab_v);
            c->tab_v = nullptr;
            GVB_CUDA(gvb_malloc(c, &c->tab_v, tab_ints * sizeof(int)));
            c->tab_v_cap = tab_ints;
        }
    }
    const size_t shift_need = (size_t)std::max(c->n_stripes, c->Mg_pad / 32);
    if (c->shift_cap < shift_need) {
        if (c->shift_u) cudaFree(c->shift_u);
        if (c->shift_v) cudaFree(c->shift_v);
        c->shift_u = c->shift_v = nullptr;
        GVB_CUDA(gvb_malloc(c, &c->shift_u, shift_need * sizeof(int)));
        GVB_CUDA(gvb_malloc(c, &c->shift_v, shift_need * sizeof(int)));
        c->shift_cap = shift_need;
    }
    const size_t acc_need = 2 * (size_t)c->Mg_pad * 4 + (size_t)c->Npad + 8;
    if (c->acc_i64_cap < acc_need) {
        if (c->acc_i64) cudaFree(c->acc_i64);
        c->acc_i64 = nullptr;
        GVB_CUDA(gvb_malloc(c, &c->acc_i64, acc_need * sizeof(unsigned long long)));
        c->acc_i64_cap = acc_need;
    }
    return GVB_OK;
}

int ax_gather(gvb_ctx* c, unsigned long long* accN, const TileTune& t, long stripe0, long n) {
    if (t.tma_gather) {
        switch (t.pair_shape) {
            case 1: return t.use_mad ? launch_ax_pair<12, 2, 1, false>(c, accN, stripe0, n) : launch_ax_pair<12, 2, 0, false>(c, accN, stripe0, n);
            case 2: return t.use_mad ? launch_ax_pair<7, 3, 1, false>(c, accN, stripe0, n) : launch_ax_pair<7, 3, 0, false>(c, accN, stripe0, n);
            default: return t.use_mad ? launch_ax_pair<11, 2, 1, false>(c, accN, stripe0, n) : launch_ax_pair<11, 2, 0, false>(c, accN, stripe0, n);
        }
    }
    if (t.variant == 1) return t.use_mad ? launch_ax<12, 3, 4, false>(c, accN, stripe0, n) : launch_ax<12, 3, 0, false>(c, accN, stripe0, n);
    switch (t.use_mad) {
        case 0: return launch_ax<15, 2, 0, false>(c, accN, stripe0, n);
        case 1: return launch_ax<15, 2, 1, false>(c, accN, stripe0, n);
        case 2: return launch_ax<15, 2, 2, false>(c, accN, stripe0, n);
        case 3: return launch_ax<15, 2, 3, false>(c, accN, stripe0, n);
        default: return launch_ax<15, 2, 4, false>(c, accN, stripe0, n);
    }
}

int ax_main(gvb_ctx* c, unsigned long long* accN) {
    const TileTune t = tune();
    // X.v on the individual-major twin (twin.cu) for as many stripes as spare HBM holds one: the walk of X^T.u, no index gather;
    // the remaining stripes gather their indices from the one matrix
    const char* tw = getenv("GVB_TWIN");
    const bool want_twin = !(tw && !strcmp(tw, "0"));
    if (want_twin && c->twin_state == 0) GVB_CHECK(gvb_twin_build(c));
    long T = (want_twin && c->twin_state > 0) ? c->twin_stripes : 0;
    if (T > 0 && t.tma_walk) {
        switch (t.pair_shape) {
            case 1: GVB_CHECK((launch_ax_pair<12, 2, 0, true>(c, accN, 0, T))); break;
            case 2: GVB_CHECK((launch_ax_pair<7, 3, 0, true>(c, accN, 0, T))); break;
            default: GVB_CHECK((launch_ax_pair<11, 2, 0, true>(c, accN, 0, T))); break;
        }
    } else if (T > 0) {
        if (t.variant == 1)
            GVB_CHECK((launch_ax<12, 3, 0, true>(c, accN, 0, T)));
        else
            GVB_CHECK((t.twin_mad > 0 ? launch_ax<15, 2, 1, true>(c, accN, 0, T) : launch_ax<15, 2, 0, true>(c, accN, 0, T)));
    }
    return ax_gather(c, accN, t, T, c->n_stripes - T);
}

int atx_main(gvb_ctx* c, const int* tab, unsigned long long* acc) {
    const TileTune t = tune();
    if (t.tma_walk) {
        switch (t.pair_shape) {
            case 1: return launch_atx_pair<12, 2, false, 0>(c, tab, acc, c->shift_u);
            case 2: return launch_atx_pair<7, 3, false, 0>(c, tab, acc, c->shift_u);
            default: return launch_atx_pair<11, 2, false, 0>(c, tab, acc, c->shift_u);
        }
    }
    if (t.variant == 1) return t.use_mad >= 4 ? launch_atx<12, 3, true, 0>(c, tab, acc, c->shift_u) : launch_atx<12, 3, false, 0>(c, tab, acc, c->shift_u);
    return t.use_mad >= 4 ? launch_atx<15, 2, true, 0>(c, tab, acc, c->shift_u) : launch_atx<15, 2, false, 0>(c, tab, acc, c->shift_u);
}

}   // namespace

// X . v of the local shard (reference data::Ax, data.cpp:848-1011, before its MPI_Allreduce): 4 launches.
// mode 1 / 2 (v unused, may be NULL): the per-individual sums of compute_people_statistics, unscaled (see ax_code_values).
int gvb_ax_tile(gvb_ctx* c, const double* v, double* out, int mode) {
    GVB_CHECK(ensure_scratch(c, false, true, false));
    const long n_tiles = c->Mg_pad / 32;
    unsigned long long* accN = c->acc_i64 + 2 * (size_t)c->Mg_pad * 4;
    ax_prep_kernel<<<(unsigned)n_tiles, 128, 0, c->stream>>>(v, c->mave, c->msig, c->scal, accN, c->Npad, mode, c->skip);
    GVB_LAUNCHED(c);
    ax_build_kernel<<<(unsigned)n_tiles, 256, 0, c->stream>>>(v, c->mave, c->msig, c->scal, c->tab_v, mode, c->skip, c->shift_v, c->tab_pairs, -1);
    GVB_LAUNCHED(c);
    GVB_CHECK(ax_main(c, accN));
    ax_finish_kernel<<<(unsigned)((c->Npad + 255) / 256), 256, 0, c->stream>>>(accN, c->scal, c->maskw, c->Npad,
                                                                               mode == 0 ? 1.0 / sqrt((double)c->N) : 1.0, out, c->skip);
    GVB_LAUNCHED(c);
    return GVB_OK;
}

// out0 = X.v0 and out1 = X.v1 of the local shard from one pass over the bed (ax_dual_kernel); bit-identical to two gvb_ax_tile calls.
// The second product keeps its scales in scal[32..], its shifts in shift_v2 and its accumulators in acc_dual.
int gvb_ax_tile_dual(gvb_ctx* c, const double* v0, const double* v1, double* out0, double* out1) {
    GVB_CHECK(ensure_scratch(c, false, true, false));
    const long n_tiles = c->Mg_pad / 32;
    const size_t tab2_ints = (size_t)n_tiles * 16384;
    if (c->tab_v2_cap < tab2_ints) {   // tables and shifts of the second product: sized by the marker tiles
        if (c->tab_v2) cudaFree(c->tab_v2);
        if (c->shift_v2) cudaFree(c->shift_v2);
        c->tab_v2 = nullptr; c->shift_v2 = nullptr;
        c->tab_v2_cap = 0;
        GVB_CUDA(gvb_malloc(c, &c->tab_v2, tab2_ints * sizeof(int)));
        GVB_CUDA(gvb_malloc(c, &c->shift_v2, (size_t)n_tiles * sizeof(int)));
        c->tab_v2_cap = tab2_ints;
    }
    if (c->acc_dual_cap < (size_t)c->Npad) {   // its accumulators: sized by the individuals (a reload may change either)
        if (c->acc_dual) cudaFree(c->acc_dual);
        c->acc_dual = nullptr;
        c->acc_dual_cap = 0;
        GVB_CUDA(gvb_malloc(c, &c->acc_dual, (size_t)c->Npad * sizeof(unsigned long long)));
        c->acc_dual_cap = (size_t)c->Npad;
    }
    unsigned long long* accN0 = c->acc_i64 + 2 * (size_t)c->Mg_pad * 4;
    unsigned long long* accN1 = c->acc_dual;
    double* scal1 = c->scal + 32;
    ax_prep_kernel<<<(unsigned)n_tiles, 128, 0, c->stream>>>(v0, c->mave, c->msig, c->scal, accN0, c->Npad, 0, c->skip);
    GVB_LAUNCHED(c);
    ax_prep_kernel<<<(unsigned)n_tiles, 128, 0, c->stream>>>(v1, c->mave, c->msig, scal1, accN1, c->Npad, 0, c->skip);
    GVB_LAUNCHED(c);
    ax_build_kernel<<<(unsigned)n_tiles, 256, 0, c->stream>>>(v0, c->mave, c->msig, c->scal, c->tab_v2, 0, c->skip, c->shift_v, c->tab_pairs, 0);
    GVB_LAUNCHED(c);
    ax_build_kernel<<<(unsigned)n_tiles, 256, 0, c->stream>>>(v1, c->mave, c->msig, scal1, c->tab_v2, 0, c->skip, c->shift_v2, c->tab_pairs, 1);
    GVB_LAUNCHED(c);
    const char* tw = getenv("GVB_TWIN");
    const bool want_twin = !(tw && !strcmp(tw, "0"));
    if (want_twin && c->twin_state == 0) GVB_CHECK(gvb_twin_build(c));
    const long T = (want_twin && c->twin_state > 0) ? c->twin_stripes : 0;
    GVB_CHECK((launch_ax_dual<12, 2, true>(c, accN0, accN1, 0, T)));
    GVB_CHECK((launch_ax_dual<12, 2, false>(c, accN0, accN1, T, c->n_stripes - T)));
    const double isn = 1.0 / sqrt((double)c->N);
    ax_finish_kernel<<<(unsigned)((c->Npad + 255) / 256), 256, 0, c->stream>>>(accN0, c->scal, c->maskw, c->Npad, isn, out0, c->skip);
    GVB_LAUNCHED(c);
    ax_finish_kernel<<<(unsigned)((c->Npad + 255) / 256), 256, 0, c->stream>>>(accN1, scal1, c->maskw, c->Npad, isn, out1, c->skip);
    GVB_LAUNCHED(c);
    return GVB_OK;
}

// X^T . u of the local shard (reference dot_product + ATx, data.cpp:728-835): 4 launches, 5 for shards with missing
// genotypes: sum_{i missing in j} U_i comes from the sparse list of misslist.cu (a gather over 2 bytes per missing
// genotype), or, when that list cannot be held in HBM, from a second walk with the table of the missing codes
int gvb_atx_tile(gvb_ctx* c, const double* u, double* out, double* outB) {
    const bool miss = c->total_missing > 0;
    if (miss && c->miss_state == 0) GVB_CHECK(gvb_misslist_build(c));
    const bool list = miss && c->miss_state == 1;
    GVB_CHECK(ensure_scratch(c, true, false, miss && !list));
    const size_t Mpad = (size_t)c->Mg_pad * 4;
    const long npos = c->n_stripes * 32, total = gvb_roundup(c->n_stripes, 2) * 8192;
    unsigned long long* acc = c->acc_i64;
    unsigned long long* accm = c->acc_i64 + Mpad;
    long long* usum = reinterpret_cast<long long*>(c->acc_i64 + 2 * Mpad + c->Npad);
    int nb = (int)std::max(1l, std::min((npos + 255) / 256, 2l * c->sm_count));
    atx_prep_kernel<<<nb, 256, 0, c->stream>>>(u, npos, c->scal, acc, (long)((miss ? 2 : 1) * Mpad), usum, c->skip);
    GVB_LAUNCHED(c);
    atx_build_kernel<<<(unsigned)c->n_stripes, 256, 0, c->stream>>>(u, 32.0, c->scal, c->tab_u, (miss && !list) ? c->tab_u + total : nullptr, usum,
                                                                    list ? c->uq : nullptr, c->skip, c->shift_u, c->tab_pairs, -1);
    GVB_LAUNCHED(c);
    GVB_CHECK(atx_main(c, c->tab_u, acc));
    if (list)
        GVB_CHECK(gvb_misslist_sum(c, accm));
    else if (miss)
        GVB_CHECK(atx_main(c, c->tab_u + total, accm));
    atx_finish_kernel<<<(unsigned)((Mpad + 255) / 256), 256, 0, c->stream>>>(acc, miss ? accm : nullptr, usum, c->scal, c->mave, c->msig, (long)Mpad, c->M,
                                                                             1.0 / sqrt((double)c->N), out, outB, c->skip, list ? GVB_CLASS_BITS * GVB_CLASS_MAX : 0);
    GVB_LAUNCHED(c);
    return GVB_OK;
}

// out0 = X^T.u0 and out1 = X^T.u1 of the local shard from one pass over the bed (atx_dual_kernel); bit-identical to two gvb_atx_tile
// calls.  Shards with missing genotypes keep two single sweeps (their second term comes from one list / one quantised copy of u).
int gvb_atx_tile_dual(gvb_ctx* c, const double* u0, const double* u1, double* out0, double* out1) {
    if (c->total_missing > 0) {
        GVB_CHECK(gvb_atx_tile(c, u0, out0, nullptr));
        return gvb_atx_tile(c, u1, out1, nullptr);
    }
    GVB_CHECK(ensure_scratch(c, true, false, false));
    const size_t tab2_ints = (size_t)c->n_stripes * 16384;
    if (c->tab_u2_cap < tab2_ints) {
        if (c->tab_u2) cudaFree(c->tab_u2);
        if (c->shift_u2) cudaFree(c->shift_u2);
        c->tab_u2 = nullptr; c->shift_u2 = nullptr;
        c->tab_u2_cap = 0;
        GVB_CUDA(gvb_malloc(c, &c->tab_u2, tab2_ints * sizeof(int)));
        GVB_CUDA(gvb_malloc(c, &c->shift_u2, (size_t)c->n_stripes * sizeof(int)));
        c->tab_u2_cap = tab2_ints;
    }
    const size_t Mpad = (size_t)c->Mg_pad * 4;
    const long npos = c->n_stripes * 32;
    unsigned long long* acc0 = c->acc_i64;
    unsigned long long* acc1 = c->acc_i64 + Mpad;   // the accumulators of the missing-genotype term: unused on this shard
    long long* usum0 = reinterpret_cast<long long*>(c->acc_i64 + 2 * Mpad + c->Npad);
    long long* usum1 = usum0 + 1;
    double* scal1 = c->scal + 32;
    int nb = (int)std::max(1l, std::min((npos + 255) / 256, 2l * c->sm_count));
    atx_prep_kernel<<<nb, 256, 0, c->stream>>>(u0, npos, c->scal, acc0, (long)Mpad, usum0, c->skip);
    GVB_LAUNCHED(c);
    atx_prep_kernel<<<nb, 256, 0, c->stream>>>(u1, npos, scal1, acc1, (long)Mpad, usum1, c->skip);
    GVB_LAUNCHED(c);
    atx_build_kernel<<<(unsigned)c->n_stripes, 256, 0, c->stream>>>(u0, 32.0, c->scal, c->tab_u2, nullptr, usum0, nullptr, c->skip, c->shift_u, c->tab_pairs, 0);
    GVB_LAUNCHED(c);
    atx_build_kernel<<<(unsigned)c->n_stripes, 256, 0, c->stream>>>(u1, 32.0, scal1, c->tab_u2, nullptr, usum1, nullptr, c->skip, c->shift_u2, c->tab_pairs, 1);
    GVB_LAUNCHED(c);
    GVB_CHECK((launch_atx_dual<12, 2>(c, acc0, acc1)));
    const double isn = 1.0 / sqrt((double)c->N);
    atx_finish_kernel<<<(unsigned)((Mpad + 255) / 256), 256, 0, c->stream>>>(acc0, nullptr, usum0, c->scal, c->mave, c->msig, (long)Mpad, c->M, isn, out0, nullptr, c->skip, 0);
    GVB_LAUNCHED(c);
    atx_finish_kernel<<<(unsigned)((Mpad + 255) / 256), 256, 0, c->stream>>>(acc1, nullptr, usum1, scal1, c->mave, c->msig, (long)Mpad, c->M, isn, out1, nullptr, c->skip, 0);
    GVB_LAUNCHED(c);
    return GVB_OK;
}

// the table walk of X^T.u with packed counters (MODE 1): acc[j] = n00 | n10 << 21 | n11 << 42 over the individuals the table weights
int gvb_count_tile_main(gvb_ctx* c, const int* tab, unsigned long long* acc) {
    const TileTune t = tune();
    if (t.tma_walk) return launch_atx_pair<11, 2, false, 1>(c, tab, acc, nullptr);
    if (t.variant == 1) return launch_atx<12, 3, false, 1>(c, tab, acc, nullptr);
    return launch_atx<15, 2, false, 1>(c, tab, acc, nullptr);
}
