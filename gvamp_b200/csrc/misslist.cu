// misslist.cu -- sparse list of the MISSING genotypes of the shard and the correction sum it serves.
//
// X^T.u (reference dot_product / ATx, data.cpp:728-835) needs, next to A_j = sum_i a_ij u_i, the sum over the
// non-missing individuals B_j = sum_i b_ij u_i = sum_i u_i - sum_{i missing in j} u_i (the na_lut factor of
// data.cpp:766).  The table walk of matvec_tile.cu gives A_j in one pass over the bed; computing the second sum
// with a second walk doubles the HBM traffic of every X^T.u on a shard that has missing genotypes (config 5:
// 1 % missing).  Missing genotypes are rare, so their coordinates are extracted ONCE after the statistics and
// the correction becomes a sparse gather:
//
//   * individuals are cut into blocks of 32768 (256 stripes); the quantised vector U of one block is 128 KB and
//     sits in shared memory, a missing genotype is a 16-bit index into it;
//   * segment (b, j) = the missing individuals of local marker j inside block b, stored block-major / marker-minor,
//     padded to a multiple of 4 entries with a sentinel that points at a zero word, so every segment is a run of
//     8-byte groups and seg_off[] counts groups;
//   * a warp takes 32 consecutive markers of one block: their segments are contiguous in memory, lanes stream
//     the groups with 8-byte loads, gather four U's from shared memory each, and a warp reduction per segment
//     ends in one int64 RED on the marker's accumulator.
//
// Bank-aware order inside a segment.  The gather's cost is shared-memory bank conflicts: 32 lanes reading 32 random words of U land
// 3.5 deep on the busiest bank on average (round 1, ncu: 201 M wavefronts for 54 M requests).  The bank of an entry is its
// individual's index mod 32 -- known when the list is built -- so the fill kernel writes every segment in an order that deals the
// banks out evenly over the gather's instructions: a segment is padded to whole passes of 128 entries (32 lanes x 4 entries, one
// 8-byte group per lane), its entries are ranked by bank, and rank s goes to instruction row (s mod R) of lane (s div R), R = 4 x
// passes.  The 32 lanes of one LDS then hold ranks R apart in bank order, i.e. (almost always) 32 different banks; the padding
// (sentinels that all read the one zero word: a broadcast) costs ~17 % more index bytes at 1 % missing.
//
// Cost per sweep: 2 bytes per missing genotype (config 5 per GPU: 8.4 GB next to the 105 GB bed) instead of a
// second 105 GB walk.  Everything is integer, so the result is bit-identical to the two-pass form (tests).
// If the list does not fit in HBM the sweep falls back to the second walk (ctx->miss_state = -1).
#include <cub/device/device_scan.cuh>

#include "gvb_internal.cuh"

namespace {

constexpr int MISS_BLOCK_STRIPES = 256;                        // stripes per individual block
constexpr int MISS_BLOCK_IND = MISS_BLOCK_STRIPES * 128;       // 32768 individuals
constexpr unsigned MISS_SENTINEL = MISS_BLOCK_IND;             // index of the always-zero word
constexpr int MISS_SMEM = (MISS_BLOCK_IND + 32) * 4;
constexpr int MISS_THREADS = 1024;

__device__ __forceinline__ uint32_t missing_bits(uint32_t w, uint32_t valid) {
    return w & ~(w >> 1) & 0x55555555u & valid;   // bit 2k of byte q: marker q, individual k carries code 01 and i < N
}

// pass 1: groups (of 4 entries) per segment.  One warp per (block b, marker group g); lane = byte position.
__global__ void __launch_bounds__(256) miss_count_kernel(const uint32_t* __restrict__ bed, const uint32_t* __restrict__ validw, long M, long Mg,
                                                         long Mg_pad, long n_stripes, unsigned long long* __restrict__ seg_groups) {
    const int lane = threadIdx.x & 31;
    const long g = blockIdx.x * 8l + (threadIdx.x >> 5);
    const long b = blockIdx.y;
    if (g >= Mg) return;
    const long t0 = b * MISS_BLOCK_STRIPES, t1 = min(n_stripes, t0 + MISS_BLOCK_STRIPES);
    int cnt[4] = {0, 0, 0, 0};
    for (long t = t0; t < t1; t++) {
        const uint32_t m = missing_bits(bed[(t * Mg_pad + g) * 32 + lane], validw[t * 32 + lane]);
#pragma unroll
        for (int q = 0; q < 4; q++) cnt[q] += __popc(m & (0xFFu << (8 * q)));
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        int x = cnt[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        // whole passes of 32 groups (128 entries), see "Bank-aware order"
        if (lane == 0) seg_groups[b * (Mg_pad * 4) + g * 4 + q] = (g * 4 + q < M) ? (unsigned long long)(32 * ((x + 127) >> 7)) : 0ull;
    }
}

// pass 2: the indices, in the bank-aware order.  Same walk, done twice by the same warp: first the per-lane counts of every (marker q,
// individual k of the byte position) -- an entry's bank is (4 lane + k) mod 32 -- from which every lane derives the rank of its first
// entry of each kind in the segment's bank-sorted order; then the entries themselves.
__global__ void __launch_bounds__(256) miss_fill_kernel(const uint32_t* __restrict__ bed, const uint32_t* __restrict__ validw, long M, long Mg,
                                                        long Mg_pad, long n_stripes, const unsigned long long* __restrict__ seg_off,
                                                        uint16_t* __restrict__ idx) {
    const int lane = threadIdx.x & 31;
    const long g = blockIdx.x * 8l + (threadIdx.x >> 5);
    const long b = blockIdx.y;
    if (g >= Mg) return;
    const long t0 = b * MISS_BLOCK_STRIPES, t1 = min(n_stripes, t0 + MISS_BLOCK_STRIPES);
    unsigned long long seg_lo[4];
    unsigned rows[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const long seg = b * (Mg_pad * 4) + g * 4 + q;
        seg_lo[q] = seg_off[seg] * 4ull;
        const unsigned long long seg_hi = seg_off[seg + 1] * 4ull;
        rows[q] = (unsigned)((seg_hi - seg_lo[q]) >> 5);                       // entries / 32 = instruction rows of the gather
        for (unsigned long long p = seg_lo[q] + lane; p < seg_hi; p += 32) idx[p] = (uint16_t)MISS_SENTINEL;
    }
    __syncwarp();
    // ---- counts per (q, k) of this lane
    unsigned cnt[4][4];
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int k = 0; k < 4; k++) cnt[q][k] = 0;
    for (long t = t0; t < t1; t++) {
        const uint32_t m = missing_bits(bed[(t * Mg_pad + g) * 32 + lane], validw[t * 32 + lane]);
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int k = 0; k < 4; k++) cnt[q][k] += (m >> (8 * q + 2 * k)) & 1u;
    }
    // ---- rank of this lane's first (q, k) entry: banks in the order beta = 4 (lane & 7) + k; within a bank the four lanes that share
    // (lane & 7) in the order of lane >> 3
    unsigned base[4][4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        unsigned before_in_lane_group = 0;   // sum over k' < k of the bank totals of this lane's (lane & 7)
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const unsigned own = cnt[q][k];
            const unsigned a8 = __shfl_xor_sync(0xffffffffu, own, 8);
            const unsigned s1 = own + a8;
            const unsigned a16 = __shfl_xor_sync(0xffffffffu, s1, 16);
            const unsigned bank_tot = s1 + a16;                                  // all four lanes of this bank
            // entries of the same bank held by lanes with a smaller lane >> 3
            unsigned prior = 0;
            if (lane & 8) prior += a8;                                           // lane ^ 8 has the smaller lane >> 3
            if (lane & 16) prior += a16;                                         // both lanes of the lower half (lane ^ 16, lane ^ 24)
            base[q][k] = before_in_lane_group + prior;
            before_in_lane_group += bank_tot;
        }
        // exclusive prefix over the eight bank groups (lane & 7): every lane of group m adds the totals of groups m' < m
        const unsigned group_tot = before_in_lane_group;
        unsigned incl = group_tot;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const unsigned up = __shfl_up_sync(0xffffffffu, incl, o, 8);
            if ((lane & 7) >= o) incl += up;
        }
        const unsigned excl = incl - group_tot;
#pragma unroll
        for (int k = 0; k < 4; k++) base[q][k] += excl;
    }
    // ---- the entries
    for (long t = t0; t < t1; t++) {
        const uint32_t m = missing_bits(bed[(t * Mg_pad + g) * 32 + lane], validw[t * 32 + lane]);
        if (m == 0u) continue;
        const unsigned local = (unsigned)((t - t0) * 32 + lane) * 4u;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (g * 4 + q >= M) continue;   // padded markers carry the pad byte (all "missing"): not part of the shard
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if ((m >> (8 * q + 2 * k)) & 1u) {
                    const unsigned s = base[q][k]++;
                    const unsigned R = rows[q], row = s % R, ln = s / R;       // s < 32 R: the segment holds every entry
                    idx[seg_lo[q] + 128u * (row >> 2) + 4u * ln + (row & 3u)] = (uint16_t)(local + k);
                }
            }
        }
    }
}

// accm[j] += sum over the missing individuals i of marker j of U_i.  Persistent CTAs; a CTA owns a contiguous run
// of (block, 32-marker batch) items in block-major order and reloads the 128 KB U block when the block changes.
__global__ void __launch_bounds__(MISS_THREADS, 1)
miss_sum_kernel(const uint16_t* __restrict__ idx, const unsigned long long* __restrict__ seg_off, const int* __restrict__ uq, long Npad, long Mpad,
                long n_items, unsigned long long* __restrict__ accm, const int* __restrict__ skip) {
    if (skip && *skip) return;
    extern __shared__ __align__(16) int U[];   // [MISS_BLOCK_IND] + zero word(s)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long n_batches = Mpad / 32;
    const long per = (n_items + gridDim.x - 1) / gridDim.x;
    const long lo = blockIdx.x * per, hi = min(n_items, lo + per);
    long item = lo;
    while (item < hi) {
        const long b = item / n_batches;
        const long item_end = min(hi, (b + 1) * n_batches);   // items of this CTA that share block b
        __syncthreads();
        const long i0 = b * MISS_BLOCK_IND;
        for (int i = threadIdx.x * 4; i < MISS_BLOCK_IND; i += MISS_THREADS * 4) {
            int4 x = make_int4(0, 0, 0, 0);
            if (i0 + i < Npad) x = *reinterpret_cast<const int4*>(uq + i0 + i);   // Npad is a multiple of 128
            *reinterpret_cast<int4*>(U + i) = x;
        }
        if (threadIdx.x < 32) U[MISS_BLOCK_IND + threadIdx.x] = 0;
        __syncthreads();
        for (long it = item + warp; it < item_end; it += MISS_THREADS / 32) {
            const long seg0 = b * Mpad + (it - b * n_batches) * 32;
            const unsigned long long o_lo = seg_off[seg0 + lane], o_hi = seg_off[seg0 + lane + 1];
            if (!__any_sync(0xffffffffu, o_hi > o_lo)) continue;
            const uint2* groups = reinterpret_cast<const uint2*>(idx);
            // software pipeline over the 32 segments: the first four passes (512 entries: a whole segment up to 1.5 % missing) of
            // segment s + 1 are in flight while segment s is gathered -- the kernel is bound by the latency of these loads (ncu:
            // long_scoreboard 5.6 per issue with one segment in flight)
            const uint2 kSentinel = make_uint2(0x80008000u, 0x80008000u);
            uint2 nxt[4];
            unsigned long long nbeg = __shfl_sync(0xffffffffu, o_lo, 0), nend = __shfl_sync(0xffffffffu, o_hi, 0);
#pragma unroll
            for (int r = 0; r < 4; r++) nxt[r] = (nbeg + lane + 32 * r < nend) ? __ldg(groups + nbeg + lane + 32 * r) : kSentinel;
#pragma unroll 1
            for (int s = 0; s < 32; s++) {
                const unsigned long long beg = nbeg, end = nend;
                uint2 e[4];
#pragma unroll
                for (int r = 0; r < 4; r++) e[r] = nxt[r];
                if (s + 1 < 32) {
                    nbeg = __shfl_sync(0xffffffffu, o_lo, s + 1);
                    nend = __shfl_sync(0xffffffffu, o_hi, s + 1);
#pragma unroll
                    for (int r = 0; r < 4; r++) nxt[r] = (nbeg + lane + 32 * r < nend) ? __ldg(groups + nbeg + lane + 32 * r) : kSentinel;
                }
                if (beg == end) continue;
                long long acc = 0;
                {
                    // 4 groups (16 gathers) per lane; |U| < 2^24, so 16 of them fit an int32
                    int a32 = 0;
#pragma unroll
                    for (int r = 0; r < 4; r++)
                        a32 += U[e[r].x & 0xFFFFu] + U[e[r].x >> 16] + U[e[r].y & 0xFFFFu] + U[e[r].y >> 16];
                    acc += a32;
                }
                for (unsigned long long gidx = beg + lane + 128; gidx < end; gidx += 128) {   // longer segments: the rest, unpipelined
#pragma unroll
                    for (int r = 0; r < 4; r++) e[r] = (gidx + 32 * r < end) ? __ldg(groups + gidx + 32 * r) : kSentinel;
                    int a32 = 0;
#pragma unroll
                    for (int r = 0; r < 4; r++)
                        a32 += U[e[r].x & 0xFFFFu] + U[e[r].x >> 16] + U[e[r].y & 0xFFFFu] + U[e[r].y >> 16];
                    acc += a32;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0 && acc != 0) atomicAdd(accm + (it - b * n_batches) * 32 + s, (unsigned long long)acc);
            }
        }
        item = item_end;
    }
}

// The same sum with ONE LANE PER MARKER: a warp still takes 32 consecutive markers of one block, but every lane walks its own
// segment (8-byte groups, 8 in flight: two 32-byte sectors per lane, the L1 serves the three groups that follow a sector's first
// touch) and owns its marker's sum: no lane sits idle on a short segment (1 % missing of a 32768-block is 82 groups for 128 lane
// slots in the warp-per-segment form above), no shuffle tree and one RED per marker and block.  ncu of the warp-per-segment form at
// 1 % missing (profiles/r01_ncu_full_miss_sum_c4shard_1pct.txt): 2.2 TB/s, long_scoreboard 5.5 / short_scoreboard 2.9 per issue,
// i.e. latency bound; this form keeps 64 KB of index stream in flight per SM.
__global__ void __launch_bounds__(MISS_THREADS, 1)
miss_sum_lane_kernel(const uint16_t* __restrict__ idx, const unsigned long long* __restrict__ seg_off, const int* __restrict__ uq, long Npad, long Mpad,
                     long n_items, unsigned long long* __restrict__ accm, const int* __restrict__ skip) {
    if (skip && *skip) return;
    extern __shared__ __align__(16) int U[];   // [MISS_BLOCK_IND] + zero word(s)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long n_batches = Mpad / 32;
    const long per = (n_items + gridDim.x - 1) / gridDim.x;
    const long lo = blockIdx.x * per, hi = min(n_items, lo + per);
    const uint2* __restrict__ groups = reinterpret_cast<const uint2*>(idx);
    long item = lo;
    while (item < hi) {
        const long b = item / n_batches;
        const long item_end = min(hi, (b + 1) * n_batches);   // items of this CTA that share block b
        __syncthreads();
        const long i0 = b * MISS_BLOCK_IND;
        for (int i = threadIdx.x * 4; i < MISS_BLOCK_IND; i += MISS_THREADS * 4) {
            int4 x = make_int4(0, 0, 0, 0);
            if (i0 + i < Npad) x = *reinterpret_cast<const int4*>(uq + i0 + i);   // Npad is a multiple of 128
            *reinterpret_cast<int4*>(U + i) = x;
        }
        if (threadIdx.x < 32) U[MISS_BLOCK_IND + threadIdx.x] = 0;
        __syncthreads();
        for (long it = item + warp; it < item_end; it += MISS_THREADS / 32) {
            const long m0 = (it - b * n_batches) * 32;
            const long seg = b * Mpad + m0 + lane;
            unsigned long long g = seg_off[seg];
            const unsigned long long end = seg_off[seg + 1];
            long long acc = 0;
            // |U| < 2^24, so the 32 gathers of one pass fit an int32
            while (__any_sync(0xffffffffu, g < end)) {
                uint2 e[8];
#pragma unroll
                for (int r = 0; r < 8; r++) e[r] = (g + r < end) ? __ldg(groups + g + r) : make_uint2(0x80008000u, 0x80008000u);
                int a32 = 0;
#pragma unroll
                for (int r = 0; r < 8; r++) a32 += U[e[r].x & 0xFFFFu] + U[e[r].x >> 16] + U[e[r].y & 0xFFFFu] + U[e[r].y >> 16];
                acc += a32;
                g += 8;
            }
            if (acc != 0) atomicAdd(accm + m0 + lane, (unsigned long long)acc);
        }
        item = item_end;
    }
}

void free_list(gvb_ctx* c) {
    if (c->miss_idx) cudaFree(c->miss_idx);
    if (c->miss_off) cudaFree(c->miss_off);
    if (c->uq) cudaFree(c->uq);
    c->miss_idx = nullptr;
    c->miss_off = nullptr;
    c->uq = nullptr;
    c->miss_nblk = 0;
    c->miss_entries = 0;
}

}   // namespace

void gvb_misslist_reset(gvb_ctx* c) {
    free_list(c);
    c->miss_state = 0;
}

// Builds the list (once per matrix, after the statistics told us there ARE missing genotypes).  Returns GVB_OK also
// when the list cannot be held (miss_state = -1: the caller keeps the two-pass form); errors are real CUDA errors.
int gvb_misslist_build(gvb_ctx* c) {
    if (c->miss_state != 0) return GVB_OK;
    const char* mode = getenv("GVB_MISS");
    if (mode && !strcmp(mode, "twopass")) {
        c->miss_state = -1;
        return GVB_OK;
    }
    const long Mpad = c->Mg_pad * 4;
    const long nblk = (c->n_stripes + MISS_BLOCK_STRIPES - 1) / MISS_BLOCK_STRIPES;
    const size_t nseg = (size_t)nblk * Mpad;
    void* tmp = nullptr;
    auto fail_soft = [&](const char* what) {
        cudaGetLastError();
        if (tmp) cudaFree(tmp);
        free_list(c);
        c->miss_state = -1;
        if (getenv("GVB_VERBOSE")) fprintf(stderr, "[gvamp_b200] missing-genotype list not built (%s): X^T.u keeps the second walk\n", what);
        return GVB_OK;
    };
    if (cudaMalloc(&c->miss_off, (nseg + 1) * sizeof(unsigned long long)) != cudaSuccess) return fail_soft("segment offsets");
    GVB_CUDA(cudaMemsetAsync(c->miss_off, 0, (nseg + 1) * sizeof(unsigned long long), c->stream));
    dim3 grid((unsigned)((c->Mg + 7) / 8), (unsigned)nblk);
    miss_count_kernel<<<grid, 256, 0, c->stream>>>(c->bed, c->validw, c->M, c->Mg, c->Mg_pad, c->n_stripes, c->miss_off);
    GVB_LAUNCHED(c);
    // groups per segment -> exclusive prefix sums, in place; entry nseg (a zero) becomes the total
    size_t tmp_bytes = 0;
    GVB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, c->miss_off, c->miss_off, (long long)(nseg + 1), c->stream));
    if (cudaMalloc(&tmp, tmp_bytes) != cudaSuccess) return fail_soft("scan workspace");
    GVB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, c->miss_off, c->miss_off, (long long)(nseg + 1), c->stream));
    c->launches++;
    unsigned long long groups = 0;
    GVB_CUDA(cudaMemcpyAsync(&groups, c->miss_off + nseg, sizeof(groups), cudaMemcpyDeviceToHost, c->stream));
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(tmp);
    tmp = nullptr;
    const size_t entries = (size_t)groups * 4;
    // the list pays off while it is much smaller than the bed (2 bytes per missing genotype against 1/4 byte per genotype):
    // beyond 1 missing genotype in 16 it would be more than half the bed and the second walk is kept
    if (entries * 2 > c->bed_words * 4 / 2) return fail_soft("more than 1/16 of the genotypes are missing");
    if (cudaMalloc(&c->miss_idx, std::max<size_t>(entries, 4) * sizeof(uint16_t)) != cudaSuccess) return fail_soft("index list");
    if (cudaMalloc(&c->uq, (size_t)c->Npad * sizeof(int)) != cudaSuccess) return fail_soft("quantised vector");
    miss_fill_kernel<<<grid, 256, 0, c->stream>>>(c->bed, c->validw, c->M, c->Mg, c->Mg_pad, c->n_stripes, c->miss_off, c->miss_idx);
    GVB_LAUNCHED(c);
    GVB_CUDA(cudaFuncSetAttribute(miss_sum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MISS_SMEM));
    GVB_CUDA(cudaFuncSetAttribute(miss_sum_lane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MISS_SMEM));
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    c->miss_nblk = nblk;
    c->miss_entries = (long)entries;
    c->miss_state = 1;
    return GVB_OK;
}

// accm[j] (zeroed by the caller) += sum_{i missing in j} uq[i]
int gvb_misslist_sum(gvb_ctx* c, unsigned long long* accm) {
    const long Mpad = c->Mg_pad * 4;
    const long n_items = c->miss_nblk * (Mpad / 32);
    const int grid = (int)std::max(1l, std::min(n_items, (long)c->sm_count));
    // default: one warp per segment -- the form whose instruction rows the bank-aware order of the list is built for; "lane": one lane
    // per marker (measured equal to the round-1 warp form at 1 % missing; kept as a cross-check of the list's content)
    const char* form = getenv("GVB_MISS_SUM");
    if (!(form && !strcmp(form, "lane")))
        miss_sum_kernel<<<grid, MISS_THREADS, MISS_SMEM, c->stream>>>(c->miss_idx, c->miss_off, c->uq, c->Npad, Mpad, n_items, accm, c->skip);
    else
        miss_sum_lane_kernel<<<grid, MISS_THREADS, MISS_SMEM, c->stream>>>(c->miss_idx, c->miss_off, c->uq, c->Npad, Mpad, n_items, accm, c->skip);
    GVB_LAUNCHED(c);
    return GVB_OK;
}

extern "C" long gvb_missing_list_entries(gvb_ctx* c) { return (c && c->miss_state == 1) ? c->miss_entries : 0; }
