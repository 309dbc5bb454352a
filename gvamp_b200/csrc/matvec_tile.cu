// matvec_tile.cu -- generation-2 main kernels of the bed mat-vecs for sm_100a.
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
// A CTA = NW warps in lock step over tiles (one __syncthreads per tile), sharing the table tile of the
// step (X.v: warps = NW stripes, same marker tile; X^T.u: warps = NW marker tiles, same stripe), which
// all threads fetch with 16-byte cp.async into a double buffer interleaved at 128 B.  Bed tiles are
// per-warp, NS stages deep.  Persistent CTAs pull rectangular work items from an atomic counter; items
// are ordered so that CTAs running at the same time share their table tiles in L2.
//
// Arithmetic: fixed point, exact after quantisation, see matvec_lut.cu (scales, pre- and post-kernels).
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
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

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
    static constexpr int THREADS = NW * 32;
    static constexpr int SMEM = TAB_REGION + NW * NS * TILE_BYTES + TILE_BYTES;   // + slack for the 4 KB alignment of the bed stages
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

// all threads: table tile (32 KB, contiguous in global) -> interleaved buffer `buf` of the table region
template <int THREADS>
__device__ __forceinline__ void stage_table(uint32_t tab_sm, int buf, const int* __restrict__ src) {
    const uint32_t dst = tab_sm + buf * 128;
#pragma unroll
    for (int c0 = 0; c0 < 2048; c0 += THREADS) {
        const int c = c0 + threadIdx.x;
        if (2048 % THREADS == 0 || c < 2048) cp_async16(dst + (c >> 3) * 256 + (c & 7) * 16, src + c * 4);
    }
}

// ===================================================================================================
//  X . v : warp = stripe, lane = byte position (4 individuals), walk over the marker tiles of a chunk
// ===================================================================================================
template <int BUF, bool USE_MAD>
__device__ __forceinline__ void ax_consume(const char* __restrict__ tabc, uint32_t bb, const unsigned (&spack)[8], int one, int (&a32)[4]) {
    const char* tb = tabc + BUF * 128;
#pragma unroll
    for (int tau = 0; tau < 32; tau++) {
        const uint32_t w = lds32(bb ^ (uint32_t)(tau << 7));   // row tau ^ lane, column lane
#pragma unroll
        for (int k = 0; k < 4; k++) {
            // the 2-bit codes of individual k in the four marker bytes -> one 8-bit index in the top byte
            const unsigned prod = (w & (0x03030303u << (2 * k))) * (0x01041040u >> (2 * k));
            // address = index * 256 + slot * 4 (bytes 2,3 = sign replication of a slot byte < 128 = 0)
            const unsigned a = prmt(prod, spack[tau >> 2], 0xCC30u | (4 + (tau & 3)));
            const int val = *reinterpret_cast<const int*>(tb + a);
            a32[k] = USE_MAD ? mad_one(val, one, a32[k]) : a32[k] + val;
        }
    }
}

template <int NW, int NS, bool USE_MAD>
__global__ void __launch_bounds__(NW * 32, 1)
ax_tile_kernel(const uint32_t* __restrict__ bed, const int* __restrict__ tabv, long Mg_pad, long n_stripes, int n_sblocks, int n_gchunks,
               int tiles_per_chunk, int* __restrict__ work_counter, unsigned long long* __restrict__ acc_out, int one) {
    using Cfg = TileCfg<NW, NS>;
    extern __shared__ __align__(1024) char smem[];
    __shared__ int s_item;
    __shared__ __align__(8) unsigned long long s_bar[NW * NS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tab_sm = smem_u32(smem);
    const uint32_t bed_sm = ((tab_sm + TAB_REGION + 4095u) & ~4095u) + warp * (NS * TILE_BYTES);   // 4 KB aligned: XOR addressing
    const uint32_t bar0 = smem_u32(&s_bar[warp * NS]);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NS; s++) mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
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
        const long t = (long)sb * NW + warp;
        const bool active = t < n_stripes;
        const long tile_lo = (long)gc * tiles_per_chunk;
        const int nt = (int)min((long)tiles_per_chunk, n_tiles - tile_lo);
        const uint32_t* bsrc = bed + ((active ? t : 0) * Mg_pad + tile_lo * 32) * 32;
        const int* tsrc = tabv + tile_lo * 8192;

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
        auto issue_tab = [&](int i) {
            if (i < nt) stage_table<Cfg::THREADS>(tab_sm, i & 1, tsrc + (long)i * 8192);
            cp_async_commit();
        };

#pragma unroll
        for (int s = 0; s < NS - 1; s++) issue_bed(s);
        issue_tab(0);
        long long acc64[4] = {0, 0, 0, 0};

        for (int i0 = 0; i0 < nt; i0 += 2) {
#define AX_STEP(B)                                                                           \
    if (i0 + B < nt) {                                                                       \
        const int i = i0 + B;                                                                \
        cp_async_wait_all();                                                                 \
        __syncthreads(); /* table i visible; everyone is done with tile i-1 */               \
        issue_tab(i + 1);                                                                    \
        issue_bed(i + NS - 1);                                                               \
        if (active) {                                                                        \
            const uint32_t s = n_use % NS;                                                   \
            mbar_wait(bar0 + 8 * s, (n_use / NS) & 1);                                       \
            n_use++;                                                                         \
            int a32[4] = {0, 0, 0, 0};                                                       \
            ax_consume<B, USE_MAD>(smem, bed_sm + s * TILE_BYTES + lane_off, spack, one, a32); \
            _Pragma("unroll") for (int k = 0; k < 4; k++) acc64[k] += (long long)a32[k];     \
        }                                                                                    \
    }
            AX_STEP(0)
            AX_STEP(1)
#undef AX_STEP
        }
        cp_async_wait_all();
        if (active) {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (acc64[k] != 0) atomicAdd(acc_out + (t * 32 + lane) * 4 + k, (unsigned long long)acc64[k]);
        }
    }
}

// ===================================================================================================
//  X^T . u : warp = marker tile (32 groups), lane = marker group (4 markers), walk over the stripes of a chunk
// ===================================================================================================
template <int BUF, bool USE_MAD>
__device__ __forceinline__ void atx_consume(const char* __restrict__ tabc, uint32_t bb, const unsigned (&spack)[8], int one, int (&a32)[4]) {
    const char* tb = tabc + BUF * 128;
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

template <int NW, int NS, bool USE_MAD>
__global__ void __launch_bounds__(NW * 32, 1)
atx_tile_kernel(const uint32_t* __restrict__ bed, const int* __restrict__ tab, long Mg_pad, long n_stripes, int n_gblocks, int n_schunks,
                int stripes_per_chunk, int* __restrict__ work_counter, unsigned long long* __restrict__ acc_out, int one) {
    using Cfg = TileCfg<NW, NS>;
    extern __shared__ __align__(1024) char smem[];
    __shared__ int s_item;
    __shared__ __align__(8) unsigned long long s_bar[NW * NS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tab_sm = smem_u32(smem);
    const uint32_t bed_sm = ((tab_sm + TAB_REGION + 4095u) & ~4095u) + warp * (NS * TILE_BYTES);
    const uint32_t bar0 = smem_u32(&s_bar[warp * NS]);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NS; s++) mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
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
        const long T = (long)gb * NW + warp;
        const bool active = T < n_tiles;
        const long t_lo = (long)sc * stripes_per_chunk;
        const int ns = (int)min((long)stripes_per_chunk, n_stripes - t_lo);
        const uint32_t* bsrc = bed + (t_lo * Mg_pad + (active ? T : 0) * 32) * 32;   // + i * Mg_pad * 32 words per stripe
        const int* tsrc = tab + t_lo * 8192;

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
        auto issue_tab = [&](int i) {
            if (i < ns) stage_table<Cfg::THREADS>(tab_sm, i & 1, tsrc + (long)i * 8192);
            cp_async_commit();
        };

#pragma unroll
        for (int s = 0; s < NS - 1; s++) issue_bed(s);
        issue_tab(0);
        long long acc64[4] = {0, 0, 0, 0};

        for (int i0 = 0; i0 < ns; i0 += 2) {
#define ATX_STEP(B)                                                                            \
    if (i0 + B < ns) {                                                                         \
        const int i = i0 + B;                                                                  \
        cp_async_wait_all();                                                                   \
        __syncthreads();                                                                       \
        issue_tab(i + 1);                                                                      \
        issue_bed(i + NS - 1);                                                                 \
        if (active) {                                                                          \
            const uint32_t s = n_use % NS;                                                     \
            mbar_wait(bar0 + 8 * s, (n_use / NS) & 1);                                         \
            n_use++;                                                                           \
            int a32[4] = {0, 0, 0, 0};                                                         \
            atx_consume<B, USE_MAD>(smem, bed_sm + s * TILE_BYTES + lane_off, spack, one, a32); \
            _Pragma("unroll") for (int q = 0; q < 4; q++) acc64[q] += (long long)a32[q];       \
        }                                                                                      \
    }
            ATX_STEP(0)
            ATX_STEP(1)
#undef ATX_STEP
        }
        cp_async_wait_all();
        if (active) {
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (acc64[q] != 0) atomicAdd(acc_out + (T * 32 + lane) * 4 + q, (unsigned long long)acc64[q]);
        }
    }
}

struct TileTune {
    int variant;   // 0: 16 warps x 2 stages, 1: 12 warps x 3 stages
    bool use_mad;
    int ax_tiles_per_chunk, atx_stripes_per_chunk;
};

TileTune tune_from_env() {
    TileTune t{0, false, 64, 64};
    if (const char* e = getenv("GVB_TILE_VARIANT")) t.variant = atoi(e);
    if (const char* e = getenv("GVB_TILE_MAD")) t.use_mad = atoi(e) != 0;
    if (const char* e = getenv("GVB_AX_TPC")) t.ax_tiles_per_chunk = std::max(1, atoi(e));
    if (const char* e = getenv("GVB_ATX_SPC")) t.atx_stripes_per_chunk = std::max(1, atoi(e));
    return t;
}
TileTune tune() { return tune_from_env(); }   // read per launch: the tests vary the chunking within one process

template <int NW, int NS, bool USE_MAD>
int launch_ax(gvb_ctx* c, unsigned long long* accN) {
    using Cfg = TileCfg<NW, NS>;
    auto kern = ax_tile_kernel<NW, NS, USE_MAD>;
    static bool attr_done = false;
    if (!attr_done) {
        GVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        attr_done = true;
    }
    const int tpc = tune().ax_tiles_per_chunk;
    const long n_tiles = c->Mg_pad / 32;
    int n_sblocks = (int)((c->n_stripes + NW - 1) / NW);
    int n_gchunks = (int)((n_tiles + tpc - 1) / tpc);
    int grid = std::min(n_sblocks * n_gchunks, c->sm_count);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, c->stream>>>(c->bed, c->tab_v, c->Mg_pad, c->n_stripes, n_sblocks, n_gchunks, tpc, c->work_counter, accN, 1);
    GVB_LAUNCHED(c);
    return GVB_OK;
}

template <int NW, int NS, bool USE_MAD>
int launch_atx(gvb_ctx* c, unsigned long long* acc) {
    using Cfg = TileCfg<NW, NS>;
    auto kern = atx_tile_kernel<NW, NS, USE_MAD>;
    static bool attr_done = false;
    if (!attr_done) {
        GVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        attr_done = true;
    }
    const int spc = tune().atx_stripes_per_chunk;
    const long n_tiles = c->Mg_pad / 32;
    int n_gblocks = (int)((n_tiles + NW - 1) / NW);
    int n_schunks = (int)((c->n_stripes + spc - 1) / spc);
    int grid = std::min(n_gblocks * n_schunks, c->sm_count);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, c->stream>>>(c->bed, c->tab_u, c->Mg_pad, c->n_stripes, n_gblocks, n_schunks, spc, c->work_counter, acc, 1);
    GVB_LAUNCHED(c);
    return GVB_OK;
}

}   // namespace

// main kernel of X.v: accN[i] += sum over local markers of the table values (c->tab_v built by the caller)
int gvb_ax_tile_main(gvb_ctx* c, unsigned long long* accN) {
    const TileTune t = tune();
    if (t.variant == 1) return t.use_mad ? launch_ax<12, 3, true>(c, accN) : launch_ax<12, 3, false>(c, accN);
    return t.use_mad ? launch_ax<16, 2, true>(c, accN) : launch_ax<16, 2, false>(c, accN);
}

// main kernel of X^T.u for shards without missing genotypes: acc[j] += sum_i a_ij U_i (c->tab_u built by the caller)
int gvb_atx_tile_main(gvb_ctx* c, unsigned long long* acc) {
    const TileTune t = tune();
    if (t.variant == 1) return t.use_mad ? launch_atx<12, 3, true>(c, acc) : launch_atx<12, 3, false>(c, acc);
    return t.use_mad ? launch_atx<16, 2, true>(c, acc) : launch_atx<16, 2, false>(c, acc);
}

// int32 accumulation window of atx_tile_kernel in table entries (one stripe: 32 positions)
int gvb_atx_tile_window() { return 32; }
