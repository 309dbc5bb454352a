// matvec_lut.cu -- generation-1 bed mat-vecs for sm_100a: table-driven, fixed-point, HBM-bound.
//
// Why tables.  At 70 % of the B200's HBM rate a GPU must consume ~18e12 genotypes/s while it can
// issue ~36e12 lane-instructions/s: the budget is < 2 instructions per genotype, so the reference's
// "decode, multiply, add per genotype in FP64" (data.cpp:758-779, :944-1007) cannot be transcribed.
// Both kernels instead look up, per 32-bit word of the striped 4x4-interleaved layout, pre-summed
// contributions of FOUR genotypes at once from shared-memory tables whose bank is private to the
// lane (conflict-free by construction):
//
//   X^T.u  (reference dot_product/ATx, data.cpp:728-835).  For byte position p (4 individuals) the
//          table T_p[B] = sum_k a(code_k(B)) * U_{4p+k}, B = 0..255, is shared by ALL markers; byte q
//          of a word is the index for marker 4g+q.  Per 4 genotypes: PRMT (address) + LDS + IADD.
//   X.v    (reference Ax, data.cpp:848-1011).  For marker group g the table H_g[e] = sum_q Q_{4g+q}(c_q(e))
//          holds the fully standardised values b(a-mu)sigma v of four markers; the index of individual k
//          is gathered from the four bytes of the word with ONE multiply
//          ((w & 0x03030303<<2k) * (0x01041040>>2k))>>24.  Lanes walk the 32 groups of a tile skewed by
//          their lane id so that every lane reads a different bank.  Per 4 genotypes: LOP3 + IMAD + PRMT +
//          LDS + IADD.  Missing genotypes and the mean correction cost nothing: they are in the table.
//
// Arithmetic.  u (resp. the per-marker values) are quantised once per sweep to integers with a
// power-of-two scale chosen so that no int32 accumulation window can overflow; everything after that
// is exact integer arithmetic (int32 windows flushed to int64, int64 atomics across CTAs), hence
// bit-reproducible and independent of tiling / summation order.  The only error is the quantisation:
// ~1e-7 norm-wise for X^T.u and ~1e-8 for X.v (tests bound it by 1e-6, the north_star tolerance).
//
// Data movement.  Packed bed bytes are read from HBM exactly once per sweep (cp.async, multi-stage);
// tables are built per sweep by a small pre-kernel, live in L2 and are re-read by the CTAs that work
// on the same stripe chunk / group chunk at the same time (the work list is ordered for that).
#include <cuda_pipeline.h>

#include "gvb_internal.cuh"

// ---------------------------------------------------------------------------------------------------
// tunables
// ---------------------------------------------------------------------------------------------------
#define AT_WARPS 16
#define AT_THREADS (AT_WARPS * 32)
#define AT_CHUNK 64          // stripes per work item == int32 accumulation window of X^T.u

#define AX_WARPS 16          // stripes per CTA
#define AX_THREADS (AX_WARPS * 32)
#define AX_TILE 32           // marker groups per table tile (== GVB_GROUP_TILE)
#define AX_GCHUNK 2048       // marker groups per work item (64 tiles)

#define TAB_REGION 65536     // [256 entries][64 slots] int32, entry stride 256 B

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// prmt.b32 in its default mode: selector nibble n picks byte (n & 7) of {b,a}; bit 3 replicates that byte's sign
__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned sel) {
    unsigned d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

__device__ __forceinline__ int dosage_i(unsigned code) { return code == 0u ? 2 : (code == 2u ? 1 : 0); }

// ===================================================================================================
//                                         X^T . u
// ===================================================================================================
// pass 1: per-block maximum of B_p = 2 * (|u_4p| + ... + |u_4p+3|), the largest |table entry| / scale
__global__ void __launch_bounds__(256) atx_bound_kernel(const double* __restrict__ u, long npos, double* __restrict__ partial) {
    double m = 0.0;
    for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < npos; p += (long)gridDim.x * blockDim.x) {
        const double2* up = reinterpret_cast<const double2*>(u + 4 * p);
        double2 a = up[0], b = up[1];
        m = fmax(m, 2.0 * (fabs(a.x) + fabs(a.y) + fabs(b.x) + fabs(b.y)));
    }
    __shared__ double sm[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) m = fmax(m, sm[w]);
        partial[blockIdx.x] = m;
    }
}

// pass 2: power-of-two scale such that `window` table entries can never overflow an int32
// scal[0] = scale, scal[1] = 1/scale, usum (int64) = 0
__global__ void scale_from_bound_kernel(const double* __restrict__ partial, int nblocks, double window, double slack, double* __restrict__ scal,
                                        long long* __restrict__ usum) {
    double m = 0.0;
    for (int b = 0; b < nblocks; b++) m = fmax(m, partial[b]);
    double s = 1.0;
    if (m > 0.0 && isfinite(m)) {
        double limit = (2147483647.0 / window - slack) / m;
        int e;
        frexp(limit, &e);              // limit = f * 2^e, f in [0.5, 1)  ->  2^(e-1) <= limit
        e = max(-1000, min(1000, e - 1));
        s = ldexp(1.0, e);
    }
    scal[0] = s;
    scal[1] = 1.0 / s;
    if (usum) *usum = 0;
}

// pass 3: tables.  tab[(t*256 + B)*32 + l] = sum_k a(code_k(B)) * U_{4p+k},  p = 32 t + l,  U = rint(u * scale);
// tabm (optional) holds the same sum over the MISSING codes (weight 1).  Also accumulates sum_i U_i.
__global__ void __launch_bounds__(256) atx_build_kernel(const double* __restrict__ u, long n_stripes, const double* __restrict__ scal,
                                                        int* __restrict__ tab, int* __restrict__ tabm, long long* __restrict__ usum) {
    long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    long total = n_stripes * 256 * 32;
    long long lsum = 0;
    if (idx < total) {
        int l = (int)(idx & 31);
        unsigned B = (unsigned)((idx >> 5) & 255);
        long t = idx >> 13;
        long p = t * 32 + l;
        const double s = scal[0];
        const double2* up = reinterpret_cast<const double2*>(u + 4 * p);
        double2 a = up[0], b = up[1];
        int U[4] = {(int)rint(a.x * s), (int)rint(a.y * s), (int)rint(b.x * s), (int)rint(b.y * s)};
        int e = 0, em = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            unsigned code = (B >> (2 * k)) & 3u;
            e += dosage_i(code) * U[k];
            em += (code == 1u) ? U[k] : 0;
        }
        tab[idx] = e;
        if (tabm) tabm[idx] = em;
        if (B == 0) lsum = (long long)U[0] + U[1] + U[2] + U[3];
    }
    // sum_i U_i (exact, order independent)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if ((threadIdx.x & 31) == 0 && lsum != 0) atomicAdd(reinterpret_cast<unsigned long long*>(usum), (unsigned long long)lsum);
}

// 32 values per lane, 32 lanes -> lane i ends up with the complete sum of value index perm(i);
// reduce-scatter butterfly: 31 exchanges instead of 160 for 32 separate warp reductions
__device__ __forceinline__ long long warp_reduce_scatter32(long long (&v)[32], int lane) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; i++) {
            long long keep = upper ? v[i + half] : v[i];
            long long send = upper ? v[i] : v[i + half];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];   // lane holds the total of original index: bit-reversal-free mapping computed by the caller
}
// index of the value whose total lane `lane` holds after warp_reduce_scatter32
__device__ __forceinline__ int reduce_scatter_index(int lane) { return lane; }

template <int GPW, int STAGES, bool HAS_MISS>
struct AtxCfg {
    static constexpr int GPB = AT_WARPS * GPW;                       // groups per block
    static constexpr int BED_STAGE = AT_WARPS * GPW * 128;           // bytes
    static constexpr int TAB_BYTES = HAS_MISS ? 2 * TAB_REGION : ((STAGES + 1) / 2) * TAB_REGION;
    static constexpr int SMEM = TAB_BYTES + STAGES * BED_STAGE;
};

// One stripe of lookups for the calling warp.  BUF is the table buffer index (compile time).
template <int GPW, int BUF, bool HAS_MISS>
__device__ __forceinline__ void atx_consume(const char* __restrict__ smem, const uint32_t* __restrict__ bedw, unsigned laneoff, int (&acc)[GPW * 4],
                                            int (&accm)[HAS_MISS ? GPW * 4 : 1]) {
    // table buffer BUF lives in region BUF>>1, slot half BUF&1 (HAS_MISS: T in region 0, Tm in region 1, buffers = halves)
    const char* tb = smem + (HAS_MISS ? 0 : (BUF >> 1) * TAB_REGION) + (BUF & 1) * 128;
#pragma unroll
    for (int gi = 0; gi < GPW; gi++) {
        uint32_t w = bedw[gi * 32];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            // address = byte_q(w) * 256 + lane * 4 : one PRMT
            unsigned a = prmt(w, laneoff, 0x6500 | (q << 4) | 4);
            acc[gi * 4 + q] += *reinterpret_cast<const int*>(tb + a);
            if (HAS_MISS) accm[gi * 4 + q] += *reinterpret_cast<const int*>(tb + TAB_REGION + a);
        }
    }
}

template <int GPW, int STAGES, bool HAS_MISS>
__global__ void __launch_bounds__(AT_THREADS, 1)
atx_lut_kernel(const uint32_t* __restrict__ bed, const int* __restrict__ tab, const int* __restrict__ tabm, long Mg_pad, long n_stripes, int n_mblocks,
               int n_chunks, int* __restrict__ work_counter, unsigned long long* __restrict__ acc_out, unsigned long long* __restrict__ accm_out) {
    using Cfg = AtxCfg<GPW, STAGES, HAS_MISS>;
    extern __shared__ __align__(1024) char smem[];
    __shared__ int s_item;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned laneoff = lane * 4;
    char* bed_sm = smem + Cfg::TAB_BYTES;
    const int n_items = n_mblocks * n_chunks;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(work_counter, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= n_items) break;
        // stripe-chunk major order: CTAs running at the same time share the chunk's tables in L2
        const int chunk = item / n_mblocks, mb = item % n_mblocks;
        const long t_lo = (long)chunk * AT_CHUNK;
        const int ns = (int)min((long)AT_CHUNK, n_stripes - t_lo);
        const long g0 = (long)mb * Cfg::GPB + (long)warp * GPW;   // first group of this warp

        int acc[GPW * 4];
        int accm[HAS_MISS ? GPW * 4 : 1];
#pragma unroll
        for (int i = 0; i < GPW * 4; i++) acc[i] = 0;
#pragma unroll
        for (int i = 0; i < (HAS_MISS ? GPW * 4 : 1); i++) accm[i] = 0;

        auto issue = [&](int s) {   // loads of stripe t_lo + s into stage s % STAGES
            if (s < ns) {
                const long t = t_lo + s;
                const int st = s % STAGES;
                // table tile: 32 KB = 2048 x 16 B, 4 per thread; destination entry stride is 256 B
                const int* src = tab + t * 8192;
                char* dst = smem + (HAS_MISS ? 0 : (st >> 1) * TAB_REGION) + (st & 1) * 128;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    int c = threadIdx.x + j * AT_THREADS;
                    cp_async16(dst + (c >> 3) * 256 + (c & 7) * 16, src + c * 4);
                }
                if (HAS_MISS) {
                    const int* srcm = tabm + t * 8192;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        int c = threadIdx.x + j * AT_THREADS;
                        cp_async16(dst + TAB_REGION + (c >> 3) * 256 + (c & 7) * 16, srcm + c * 4);
                    }
                }
                // this warp's GPW units (128 B each, consecutive in memory): GPW*8 chunks of 16 B
                const char* bsrc = reinterpret_cast<const char*>(bed + (t * Mg_pad + g0) * 32);
                char* bdst = bed_sm + st * Cfg::BED_STAGE + warp * (GPW * 128);
#pragma unroll
                for (int j = 0; j < (GPW * 8 + 31) / 32; j++) {
                    int c = lane + j * 32;
                    if (c < GPW * 8) cp_async16(bdst + c * 16, bsrc + c * 16);
                }
            }
            cp_async_commit();
        };

#pragma unroll
        for (int s = 0; s < STAGES - 1; s++) issue(s);

        // main loop, unrolled by STAGES so that every table / stage offset is a compile-time constant
        for (int s0 = 0; s0 < ns; s0 += STAGES) {
#define ATX_STAGE(B)                                                                                                              \
    if constexpr (STAGES > B) {                                                                                                   \
        if (s0 + B < ns) {                                                                                                        \
            cp_async_wait<STAGES - 2>();                                                                                          \
            __syncthreads(); /* stripe s visible to all; everyone is done with stripe s-1 */                                      \
            issue(s0 + B + STAGES - 1); /* refill the stage that stripe s-1 used */                                               \
            const uint32_t* bw = reinterpret_cast<const uint32_t*>(bed_sm + B * Cfg::BED_STAGE + warp * (GPW * 128)) + lane;      \
            atx_consume<GPW, B, HAS_MISS>(smem, bw, laneoff, acc, accm);                                                          \
        }                                                                                                                         \
    }
            ATX_STAGE(0)
            ATX_STAGE(1)
            ATX_STAGE(2)
#undef ATX_STAGE
        }
        cp_async_wait<0>();

        // cross-lane reduction of the GPW*4 per-position partial sums and one int64 atomic per marker
        constexpr int NV = GPW * 4;
#pragma unroll
        for (int base = 0; base < NV; base += 32) {
            long long v[32];
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = (base + i < NV) ? (long long)acc[(base + i < NV) ? base + i : 0] : 0ll;
            long long tot = warp_reduce_scatter32(v, lane);
            // after the butterfly lane L holds value index rev: at each round the kept half is selected by the lane bit
            int idx = 0;
#pragma unroll
            for (int half = 16; half >= 1; half >>= 1) idx += (lane & half) ? half : 0;
            long j = (g0 * 4) + base + idx;
            if (base + idx < NV && j < Mg_pad * 4 && tot != 0) atomicAdd(acc_out + j, (unsigned long long)tot);
            if (HAS_MISS) {
#pragma unroll
                for (int i = 0; i < 32; i++) v[i] = (base + i < NV) ? (long long)accm[(base + i < NV) ? base + i : 0] : 0ll;
                long long totm = warp_reduce_scatter32(v, lane);
                if (base + idx < NV && j < Mg_pad * 4 && totm != 0) atomicAdd(accm_out + j, (unsigned long long)totm);
            }
        }
    }
}

// out[j] = sigma_j * (A_j - mu_j * B_j) / scale / sqrt(N),  B_j = sum_i U_i - (missing part)
__global__ void atx_finish_kernel(const unsigned long long* __restrict__ acc, const unsigned long long* __restrict__ accm, const long long* __restrict__ usum,
                                  const double* __restrict__ scal, const double* __restrict__ mave, const double* __restrict__ msig, long Mpad,
                                  double inv_sqrt_n, double* __restrict__ out) {
    long j = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (j >= Mpad) return;
    long long A = (long long)acc[j];
    long long Bs = *usum - (accm ? (long long)accm[j] : 0ll);
    double inv_s = scal[1];
    out[j] = msig[j] * ((double)A * inv_s - mave[j] * ((double)Bs * inv_s)) * inv_sqrt_n;
}

template <int GPW, int STAGES, bool HAS_MISS>
static int launch_atx(gvb_ctx* c, unsigned long long* acc, unsigned long long* accm) {
    using Cfg = AtxCfg<GPW, STAGES, HAS_MISS>;
    auto kern = atx_lut_kernel<GPW, STAGES, HAS_MISS>;
    static unsigned long long attr_done = 0;   // one bit per device
    if (!(attr_done >> (c->device & 63) & 1ull)) {
        GVB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        attr_done |= 1ull << (c->device & 63);
    }
    int n_mblocks = (int)((c->Mg_pad + Cfg::GPB - 1) / Cfg::GPB);
    int n_chunks = (int)((c->n_stripes + AT_CHUNK - 1) / AT_CHUNK);
    int grid = std::min(n_mblocks * n_chunks, c->sm_count);
    kern<<<grid, AT_THREADS, Cfg::SMEM, c->stream>>>(c->bed, c->tab_u, HAS_MISS ? c->tab_u + (size_t)c->n_stripes * 8192 : nullptr, c->Mg_pad, c->n_stripes,
                                                     n_mblocks, n_chunks, c->work_counter, acc, accm);
    GVB_LAUNCHED(c);
    return GVB_OK;
}

int gvb_atx_lut(gvb_ctx* c, const double* u, double* out) {
    const bool miss = c->total_missing > 0;
    const long npos = c->n_stripes * 32;
    const size_t tab_ints = (size_t)c->n_stripes * 8192 * (miss ? 2 : 1);
    if (c->tab_u_cap < tab_ints) {
        if (c->tab_u) cudaFree(c->tab_u);
        c->tab_u = nullptr;
        GVB_CUDA(cudaMalloc(&c->tab_u, tab_ints * sizeof(int)));
        c->tab_u_cap = tab_ints;
    }
    const size_t Mpad = (size_t)c->Mg_pad * 4;
    const size_t acc_need = 2 * Mpad + (size_t)c->Npad + 8;
    if (c->acc_i64_cap < acc_need) {
        if (c->acc_i64) cudaFree(c->acc_i64);
        c->acc_i64 = nullptr;
        GVB_CUDA(cudaMalloc(&c->acc_i64, acc_need * sizeof(unsigned long long)));
        c->acc_i64_cap = acc_need;
    }
    unsigned long long* acc = c->acc_i64;
    unsigned long long* accm = c->acc_i64 + Mpad;
    long long* usum = reinterpret_cast<long long*>(c->acc_i64 + 2 * Mpad + c->Npad);
    GVB_CUDA(cudaMemsetAsync(acc, 0, (miss ? 2 : 1) * Mpad * sizeof(unsigned long long), c->stream));
    GVB_CUDA(cudaMemsetAsync(c->work_counter, 0, sizeof(int), c->stream));

    int nb = (int)std::max(1l, std::min((npos + 255) / 256, (long)GVB_RED_BLOCKS));
    atx_bound_kernel<<<nb, 256, 0, c->stream>>>(u, npos, c->red_partial);
    GVB_LAUNCHED(c);
    scale_from_bound_kernel<<<1, 1, 0, c->stream>>>(c->red_partial, nb, (double)std::min((long)AT_CHUNK, c->n_stripes), 4.0, c->scal, usum);
    GVB_LAUNCHED(c);
    long total = c->n_stripes * 8192;
    atx_build_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(u, c->n_stripes, c->scal, c->tab_u, miss ? c->tab_u + total : nullptr, usum);
    GVB_LAUNCHED(c);
    if (miss)
        GVB_CHECK((launch_atx<6, 2, true>(c, acc, accm)));
    else
        GVB_CHECK((launch_atx<12, 3, false>(c, acc, accm)));
    atx_finish_kernel<<<(unsigned)((Mpad + 255) / 256), 256, 0, c->stream>>>(acc, miss ? accm : nullptr, usum, c->scal, c->mave, c->msig, (long)Mpad,
                                                                             1.0 / sqrt((double)c->N), out);
    GVB_LAUNCHED(c);
    return GVB_OK;
}

// ===================================================================================================
//                                           X . v
// ===================================================================================================
// pass 1: E_j = max_c |b(a-mu) sigma v| per marker, summed over each tile of 32 groups (128 markers)
__global__ void __launch_bounds__(128) ax_bound_kernel(const double* __restrict__ v, const double* __restrict__ mave, const double* __restrict__ msig,
                                                       double* __restrict__ tile_sum) {
    long j = blockIdx.x * 128l + threadIdx.x;   // blockIdx.x == tile
    double w = msig[j] * v[j], mu = mave[j];
    double e = fmax(fabs((2.0 - mu) * w), fabs(mu * w));
    e = fmax(e, fabs((1.0 - mu) * w));
    __shared__ double sm[4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = sm[0] + sm[1] + sm[2] + sm[3];
}

__global__ void __launch_bounds__(256) max_kernel(const double* __restrict__ x, long n, double* __restrict__ partial) {
    double m = 0.0;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) m = fmax(m, x[i]);
    __shared__ double sm[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) m = fmax(m, sm[w]);
        partial[blockIdx.x] = m;
    }
}

// pass 3: tabv[(T*256 + e)*32 + s] = sum_q rint(val_{4g+q}(c_q(e)) * scale),  g = 32 T + s,
// val_j(00) = (2-mu) w, val_j(10) = (1-mu) w, val_j(11) = -mu w, val_j(01 = missing) = 0,  w = sigma_j v_j
__global__ void __launch_bounds__(256) ax_build_kernel(const double* __restrict__ v, const double* __restrict__ mave, const double* __restrict__ msig,
                                                       long n_tiles, const double* __restrict__ scal, int* __restrict__ tabv) {
    long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (idx >= n_tiles * 8192) return;
    int s = (int)(idx & 31);
    unsigned e = (unsigned)((idx >> 5) & 255);
    long T = idx >> 13;
    long g = T * 32 + s;
    const double sc = scal[2];
    int sum = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        long j = g * 4 + q;
        unsigned code = (e >> (2 * q)) & 3u;
        double w = msig[j] * v[j], mu = mave[j];
        double val = code == 0u ? (2.0 - mu) * w : (code == 2u ? (1.0 - mu) * w : (code == 3u ? -mu * w : 0.0));
        sum += (int)rint(val * sc);
    }
    tabv[idx] = sum;
}

__global__ void __launch_bounds__(AX_THREADS, 1)
ax_lut_kernel(const uint32_t* __restrict__ bed, const int* __restrict__ tabv, long Mg_pad, long n_stripes, int n_sblocks, int n_gchunks,
              int* __restrict__ work_counter, unsigned long long* __restrict__ acc_out) {
    extern __shared__ __align__(1024) char smem[];
    __shared__ int s_item;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    char* bed_sm = smem + TAB_REGION + warp * 8192;       // two 4 KB buffers per warp
    const int n_items = n_sblocks * n_gchunks;

    // slot offsets of the skewed walk, packed 4 per register: byte (tau & 3) of spack[tau >> 2] = ((tau + lane) & 31) * 4
    unsigned spack[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        unsigned x = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) x |= ((unsigned)(((4 * j + b + lane) & 31) * 4)) << (8 * b);
        spack[j] = x;
    }

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(work_counter, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= n_items) break;
        // group-chunk major order: concurrent CTAs share the chunk's tables in L2
        const int gc = item / n_sblocks, sb = item % n_sblocks;
        const long t = (long)sb * AX_WARPS + warp;
        const bool active = t < n_stripes;
        const long tile_lo = (long)gc * (AX_GCHUNK / AX_TILE);
        const int nt = (int)min((long)(AX_GCHUNK / AX_TILE), Mg_pad / AX_TILE - tile_lo);

        long long acc64[4] = {0, 0, 0, 0};

        auto issue = [&](int i) {   // tile tile_lo + i -> buffer i & 1
            if (i < nt) {
                const long T = tile_lo + i;
                const int buf = i & 1;
                const int* src = tabv + T * 8192;
                char* dst = smem + buf * 128;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    int c = threadIdx.x + j * AX_THREADS;
                    cp_async16(dst + (c >> 3) * 256 + (c & 7) * 16, src + c * 4);
                }
                if (active) {
                    // unit u of the tile goes to row (u - lane) & 31: at step tau every lane reads row tau and
                    // thereby walks the groups skewed by its lane id
                    const uint32_t* bsrc = bed + (t * Mg_pad + T * 32) * 32 + lane;
                    char* bdst = bed_sm + buf * 4096 + lane * 4;
#pragma unroll
                    for (int u = 0; u < 32; u++) cp_async4(bdst + ((u - lane) & 31) * 128, bsrc + u * 32);
                }
            }
            cp_async_commit();
        };

        issue(0);
        for (int i0 = 0; i0 < nt; i0 += 2) {
#pragma unroll
            for (int b = 0; b < 2; b++) {
                const int i = i0 + b;
                if (i < nt) {
                    cp_async_wait<0>();
                    __syncthreads();
                    issue(i + 1);
                    if (active) {
                        int a32[4] = {0, 0, 0, 0};
                        const uint32_t* bw = reinterpret_cast<const uint32_t*>(bed_sm + b * 4096) + lane;
                        const char* tb = smem + b * 128;
#pragma unroll
                        for (int tau = 0; tau < 32; tau++) {
                            const uint32_t w = bw[tau * 32];
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                // codes of individual k in the four marker bytes -> one 8-bit index in the top byte
                                unsigned prod = (w & (0x03030303u << (2 * k))) * (0x01041040u >> (2 * k));
                                // address = index * 256 + slot offset of this step (sign-replicated zero in the upper bytes)
                                unsigned a = prmt(prod, spack[tau >> 2], 0xCC30 | (4 + (tau & 3)));
                                a32[k] += *reinterpret_cast<const int*>(tb + a);
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 4; k++) acc64[k] += (long long)a32[k];
                    }
                }
            }
        }
        cp_async_wait<0>();
        if (active) {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (acc64[k] != 0) atomicAdd(acc_out + (t * 32 + lane) * 4 + k, (unsigned long long)acc64[k]);
        }
    }
}

__global__ void ax_finish_kernel2(const unsigned long long* __restrict__ acc, const double* __restrict__ scal, const uint32_t* __restrict__ maskw, long Npad,
                                  double inv_sqrt_n, double* __restrict__ out) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= Npad) return;
    bool present = (maskw[i >> 2] >> (2 * (i & 3))) & 1u;
    out[i] = present ? (double)(long long)acc[i] * scal[3] * inv_sqrt_n : 0.0;
}

int gvb_ax_lut(gvb_ctx* c, const double* v, double* out) {
    const long n_tiles = c->Mg_pad / AX_TILE;
    const size_t tab_ints = (size_t)n_tiles * 8192;
    if (c->tab_v_cap < tab_ints) {
        if (c->tab_v) cudaFree(c->tab_v);
        c->tab_v = nullptr;
        GVB_CUDA(cudaMalloc(&c->tab_v, tab_ints * sizeof(int)));
        c->tab_v_cap = tab_ints;
    }
    const size_t Mpad = (size_t)c->Mg_pad * 4;
    const size_t acc_need = 2 * Mpad + (size_t)c->Npad + 8;
    if (c->acc_i64_cap < acc_need) {
        if (c->acc_i64) cudaFree(c->acc_i64);
        c->acc_i64 = nullptr;
        GVB_CUDA(cudaMalloc(&c->acc_i64, acc_need * sizeof(unsigned long long)));
        c->acc_i64_cap = acc_need;
    }
    unsigned long long* accN = c->acc_i64 + 2 * Mpad;
    GVB_CUDA(cudaMemsetAsync(accN, 0, (size_t)c->Npad * sizeof(unsigned long long), c->stream));
    GVB_CUDA(cudaMemsetAsync(c->work_counter, 0, sizeof(int), c->stream));

    // scale: the int32 window of X.v is one tile (32 lookups per individual)
    ax_bound_kernel<<<(unsigned)n_tiles, 128, 0, c->stream>>>(v, c->mave, c->msig, c->wv);
    GVB_LAUNCHED(c);
    int nb = (int)std::max(1l, std::min((n_tiles + 255) / 256, (long)GVB_RED_BLOCKS));
    max_kernel<<<nb, 256, 0, c->stream>>>(c->wv, n_tiles, c->red_partial);
    GVB_LAUNCHED(c);
    scale_from_bound_kernel<<<1, 1, 0, c->stream>>>(c->red_partial, nb, 1.0, 64.0 + 1.0, c->scal + 2, nullptr);
    GVB_LAUNCHED(c);
    long total = n_tiles * 8192;
    ax_build_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(v, c->mave, c->msig, n_tiles, c->scal, c->tab_v);
    GVB_LAUNCHED(c);

    static unsigned long long attr_done = 0;   // one bit per device
    const int smem = TAB_REGION + AX_WARPS * 8192;
    if (!(attr_done >> (c->device & 63) & 1ull)) {
        GVB_CUDA(cudaFuncSetAttribute(ax_lut_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_done |= 1ull << (c->device & 63);
    }
    int n_sblocks = (int)((c->n_stripes + AX_WARPS - 1) / AX_WARPS);
    int n_gchunks = (int)((c->Mg_pad + AX_GCHUNK - 1) / AX_GCHUNK);
    int grid = std::min(n_sblocks * n_gchunks, c->sm_count);
    ax_lut_kernel<<<grid, AX_THREADS, smem, c->stream>>>(c->bed, c->tab_v, c->Mg_pad, c->n_stripes, n_sblocks, n_gchunks, c->work_counter, accN);
    GVB_LAUNCHED(c);
    ax_finish_kernel2<<<(unsigned)((c->Npad + 255) / 256), 256, 0, c->stream>>>(accN, c->scal, c->maskw, c->Npad, 1.0 / sqrt((double)c->N), out);
    GVB_LAUNCHED(c);
    return GVB_OK;
}
