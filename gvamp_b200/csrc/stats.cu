// stats.cu -- per-marker genotype-code counts (pure popcount, bit-exact) and the marker mean /
// inverse standard deviation derived from them.
//
// Replaces data::compute_markers_statistics (reference data.cpp:392-485, scalar-branch semantics).
// The reference makes two passes over each column with a 256-entry double LUT; here one pass counts
// the masked codes n00 / n10 / n11 with popcounts and the statistics follow in closed form
// (SURVEY.md 8a3):  mu = (2 n00 + n10) / (n00 + n10 + n11),
//                   SS = n00 (2-mu)^2 + n10 (1-mu)^2 + n11 mu^2,  msig = 1/sqrt(SS/(nonas-1))^alpha.
#include "gvb_internal.cuh"

// one warp per marker group; lane l walks position l of every stripe
__global__ void __launch_bounds__(256) counts_kernel(const uint32_t* __restrict__ bed, const uint32_t* __restrict__ maskw,
                                                     const uint32_t* __restrict__ validw, long Mg, long Mg_pad, long n_stripes,
                                                     int64_t* __restrict__ counts) {
    int lane = threadIdx.x & 31;
    long g = blockIdx.x * (long)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= Mg) return;
    // cnt[q][c] masked, cnt[q][4+c] unmasked
    int cnt[4][8];
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int c = 0; c < 8; c++) cnt[q][c] = 0;
    for (long t = 0; t < n_stripes; t++) {
        uint32_t w = bed[(t * Mg_pad + g) * 32 + lane];
        uint32_t mw = maskw[t * 32 + lane];
        uint32_t vw = validw[t * 32 + lane];
        uint32_t lo = w & 0x55555555u;
        uint32_t hi = (w >> 1) & 0x55555555u;
        uint32_t cls[4];
        cls[0] = ~(lo | hi) & 0x55555555u;  // 00
        cls[1] = lo & ~hi;                  // 01 missing
        cls[2] = hi & ~lo;                  // 10
        cls[3] = hi & lo;                   // 11
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t m = cls[c] & mw, v = cls[c] & vw;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                cnt[q][c] += __popc(m & (0xFFu << (8 * q)));
                cnt[q][4 + c] += __popc(v & (0xFFu << (8 * q)));
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int c = 0; c < 8; c++) {
            int x = cnt[q][c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) counts[(g * 4 + q) * 8 + c] = (int64_t)x;
        }
}

// ---- gen 2: the counts as a table walk (matvec_tile.cu, MODE 1) -------------------------------------------------
// tab[(t*256 + B)*32 + l] = sum over the individuals k of position p = 32t+l that the weight word selects of
// 1 << (10 * class(code_k(B))), class 00 -> 0, 10 -> 1, 11 -> 2; the missing code adds nothing (n01 follows from the total)
__global__ void __launch_bounds__(256) count_table_kernel(const uint32_t* __restrict__ weightw, long n_stripes, int* __restrict__ tab, int pairs) {
    long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (idx >= n_stripes * 8192) return;
    int l = (int)(idx & 31);
    unsigned B = (unsigned)((idx >> 5) & 255);
    long t = idx >> 13;
    uint32_t mw = weightw[t * 32 + l];
    int e = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        unsigned code = (B >> (2 * k)) & 3u;
        bool present = (mw >> (2 * k)) & 1u;
        if (present && code != 1u) e += 1 << (10 * (code == 0u ? 0 : (code == 2u ? 1 : 2)));
    }
    tab[gvb_tab_index(t, (int)B, l, pairs)] = e;
}

// packed n00 | n10 << 21 | n11 << 42 -> counts[j*8 + base + {0,1,2,3}] = n00, n01, n10, n11 (n01 = total - the rest)
__global__ void unpack_counts_kernel(const unsigned long long* __restrict__ packed, long Mpad, long total, int base, bool also_other,
                                     int64_t* __restrict__ counts) {
    long j = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (j >= Mpad) return;
    unsigned long long p = packed[j];
    long n00 = (long)(p & 0x1FFFFFull), n10 = (long)((p >> 21) & 0x1FFFFFull), n11 = (long)(p >> 42);
    long n01 = total - n00 - n10 - n11;
    for (int b = base; b <= (also_other ? 4 : base); b += 4) {
        counts[j * 8 + b + 0] = n00;
        counts[j * 8 + b + 1] = n01;
        counts[j * 8 + b + 2] = n10;
        counts[j * 8 + b + 3] = n11;
    }
}

static int counts_tile_run(gvb_ctx* c) {
    const size_t tab_ints = (size_t)gvb_roundup(c->n_stripes, 2) * 8192;   // whole pairs of steps (pair mode)
    if (c->tab_u_cap < tab_ints) {
        if (c->tab_u) cudaFree(c->tab_u);
        c->tab_u = nullptr;
        GVB_CUDA(gvb_malloc(c, &c->tab_u, tab_ints * sizeof(int)));
        c->tab_u_cap = tab_ints;
    }
    const size_t Mpad = (size_t)c->Mg_pad * 4;
    const size_t acc_need = 2 * Mpad + (size_t)c->Npad + 8;
    if (c->acc_i64_cap < acc_need) {
        if (c->acc_i64) cudaFree(c->acc_i64);
        c->acc_i64 = nullptr;
        GVB_CUDA(gvb_malloc(c, &c->acc_i64, acc_need * sizeof(unsigned long long)));
        c->acc_i64_cap = acc_need;
    }
    // no phenotype NAs: the masked and the unmasked counts coincide and one walk serves both
    const bool same = c->mask_present == c->N;
    for (int pass = 0; pass < (same ? 1 : 2); pass++) {
        GVB_CUDA(cudaMemsetAsync(c->acc_i64, 0, Mpad * sizeof(unsigned long long), c->stream));
        count_table_kernel<<<(unsigned)((tab_ints + 255) / 256), 256, 0, c->stream>>>(pass == 0 ? c->maskw : c->validw, c->n_stripes, c->tab_u, c->tab_pairs);
        GVB_LAUNCHED(c);
        GVB_CHECK(gvb_count_tile_main(c, c->tab_u, c->acc_i64));
        unpack_counts_kernel<<<(unsigned)((Mpad + 255) / 256), 256, 0, c->stream>>>(c->acc_i64, (long)Mpad, pass == 0 ? c->mask_present : c->N, pass == 0 ? 0 : 4,
                                                                                     same, c->counts);
        GVB_LAUNCHED(c);
    }
    return GVB_OK;
}

__global__ void stats_from_counts_kernel(const int64_t* __restrict__ counts, long M, long Mpad, int nonas, double alpha_scale,
                                         double* __restrict__ mave, double* __restrict__ msig, unsigned long long* __restrict__ total_missing) {
    long j = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (j >= Mpad) return;
    if (j >= M) {  // padded markers: sigma = 0 keeps them out of every product
        mave[j] = 0.0;
        msig[j] = 0.0;
        return;
    }
    double n00 = (double)counts[j * 8 + 0], n10 = (double)counts[j * 8 + 2], n11 = (double)counts[j * 8 + 3];
    double suma = 2.0 * n00 + n10, sumb = n00 + n10 + n11;
    double mu = (sumb != 0.0) ? suma / sumb : 0.0;                       // data.cpp:462-465
    double d2 = 2.0 - mu, d1 = 1.0 - mu;
    double ss = n00 * (d2 * d2) + n10 * (d1 * d1) + n11 * (mu * mu);
    double sig;
    if (ss != 0.0) {                                                     // data.cpp:475-483
        double sd = sqrt(ss / ((double)nonas - 1.0));
        sig = (alpha_scale == 1.0) ? 1.0 / sd : 1.0 / pow(sd, alpha_scale);
    } else {
        sig = 1.0;
    }
    mave[j] = mu;
    msig[j] = sig;
    long miss = counts[j * 8 + 5];
    if (miss) atomicAdd(total_missing, (unsigned long long)miss);
}

int gvb_stats_run(gvb_ctx* c) {
    int warps = 8;
    long blocks = (c->Mg + warps - 1) / warps;
    GVB_CUDA(cudaMemsetAsync(c->counts, 0, (size_t)c->Mg_pad * 4 * 8 * sizeof(int64_t), c->stream));
    if (c->kernel_gen >= 2 && c->N < (1l << 21)) {   // 21-bit packed counters
        GVB_CHECK(counts_tile_run(c));
    } else {
        counts_kernel<<<(unsigned)blocks, warps * 32, 0, c->stream>>>(c->bed, c->maskw, c->validw, c->Mg, c->Mg_pad, c->n_stripes, c->counts);
        GVB_LAUNCHED(c);
    }
    unsigned long long* d_miss = (unsigned long long*)c->scal;
    GVB_CUDA(cudaMemsetAsync(d_miss, 0, sizeof(unsigned long long), c->stream));
    long Mpad = c->Mg_pad * 4;
    stats_from_counts_kernel<<<(unsigned)((Mpad + 255) / 256), 256, 0, c->stream>>>(c->counts, c->M, Mpad, c->nonas, c->alpha_scale, c->mave,
                                                                                    c->msig, d_miss);
    GVB_LAUNCHED(c);
    unsigned long long miss = 0;
    GVB_CUDA(cudaMemcpyAsync(&miss, d_miss, sizeof(miss), cudaMemcpyDeviceToHost, c->stream));
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    c->total_missing = (long)miss;
    c->have_stats = true;
    return GVB_OK;
}

extern "C" int gvb_compute_stats(gvb_ctx* c, double alpha_scale) {
    GVB_ARG(c && c->bed, "ctx / matrix not loaded");
    c->alpha_scale = alpha_scale;
    return gvb_stats_run(c);
}

extern "C" int gvb_get_stats(gvb_ctx* c, double* mave, double* msig) {
    GVB_ARG(c && c->have_stats, "statistics not computed");
    if (mave) GVB_CUDA(cudaMemcpyAsync(mave, c->mave, c->M * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (msig) GVB_CUDA(cudaMemcpyAsync(msig, c->msig, c->M * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    return GVB_OK;
}

extern "C" int gvb_get_counts(gvb_ctx* c, int64_t* counts) {
    GVB_ARG(c && c->have_stats && counts, "statistics not computed");
    GVB_CUDA(cudaMemcpyAsync(counts, c->counts, c->M * 8 * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    return GVB_OK;
}
