// layout.cu -- the HBM-resident genotype matrix: allocation, re-tiling of PLINK SNP-major bytes
// into the striped 4-marker-interleaved layout (gvb_internal.cuh), decode back to PLINK bytes,
// file / host loaders and the synthetic generator.
//
// Replaces data::read_genotype_data (reference data.cpp:201-234) and the bed_data member
// (data.hpp:45): the shard is read at file offset 3 + S*mbytes, uploaded once and stays in HBM.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>

#include "gvb_internal.cuh"

// ------------------------------------------------------------------------------------------------
__global__ void retile_kernel(const uint8_t* __restrict__ snp, long nmark, long mbytes, long g0, long M_local, long j0,
                              uint32_t* __restrict__ bed, long Mg_pad, long n_stripes) {
    // one thread per output word (t, g, l); g runs over the groups of this chunk
    long ngroups = (nmark + 3) / 4;
    long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    long total = n_stripes * ngroups * 32;
    if (idx >= total) return;
    int l = (int)(idx & 31);
    long gi = (idx >> 5) % ngroups;
    long t = (idx >> 5) / ngroups;
    long p = t * 32 + l;
    uint32_t w = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        long jj = gi * 4 + q;   // marker inside the chunk
        uint32_t b = GVB_PAD_BYTE;
        if (jj < nmark && (j0 + jj) < M_local && p < mbytes) b = snp[jj * mbytes + p];
        w |= b << (8 * q);
    }
    bed[(t * Mg_pad + (g0 + gi)) * 32 + l] = w;
}

__global__ void decode_kernel(const uint32_t* __restrict__ bed, long Mg_pad, long mbytes, long j0, long n, uint8_t* __restrict__ out) {
    long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (idx >= n * mbytes) return;
    long jj = idx / mbytes, p = idx % mbytes;
    long j = j0 + jj;
    long t = p >> 5;
    int l = (int)(p & 31);
    uint32_t w = bed[(t * Mg_pad + (j >> 2)) * 32 + l];
    out[idx] = (uint8_t)(w >> (8 * (j & 3)));
}

// ---- synthetic generator: must stay byte-identical to oracle/gvamp_oracle.c:orc_synth_bed ----
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void synth_kernel(uint32_t* __restrict__ bed, uint64_t seed, long N, long mbytes, long S, long M, long Mg, long Mg_pad,
                             long n_stripes, unsigned miss_thr16) {
    long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    long total = n_stripes * Mg * 32;
    if (idx >= total) return;
    int l = (int)(idx & 31);
    long g = (idx >> 5) % Mg;
    long t = (idx >> 5) / Mg;
    long p = t * 32 + l;
    uint64_t hs = mix64(seed);
    uint32_t w = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        long jl = g * 4 + q;
        uint32_t byte = GVB_PAD_BYTE;
        if (jl < M && p < mbytes) {
            uint64_t hj = mix64(hs ^ ((uint64_t)(S + jl + 1) * 0xD1B54A32D192ED03ull));
            // explicit round-to-nearest mul/add: no FMA contraction, same bits as the C oracle
            double pj = __dadd_rn(0.01, __dmul_rn(0.49, __dmul_rn((double)(hj >> 11), 1.0 / 9007199254740992.0)));
            unsigned thr = (unsigned)__dmul_rn(pj, 65536.0);
            uint64_t h0 = mix64(hj + 0x9E3779B97F4A7C15ull * (uint64_t)(2 * p + 1));
            uint64_t h1 = mix64(hj + 0x9E3779B97F4A7C15ull * (uint64_t)(2 * p + 2));
            uint64_t hm = miss_thr16 ? mix64((hj ^ 0xA5A5A5A5A5A5A5A5ull) + 0xC2B2AE3D27D4EB4Full * (uint64_t)(p + 1)) : 0ull;
            byte = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                long i = 4 * p + k;
                unsigned code = 0u;
                if (i < N) {
                    uint64_t h = (k < 2) ? h0 : h1;
                    int sh = (k & 1) * 32;
                    unsigned d = ((((unsigned)(h >> sh)) & 0xFFFFu) < thr) + ((((unsigned)(h >> (sh + 16))) & 0xFFFFu) < thr);
                    code = d == 2u ? 0u : (d == 1u ? 2u : 3u);
                    if (miss_thr16 && ((((unsigned)(hm >> (16 * k))) & 0xFFFFu) < miss_thr16)) code = 1u;
                }
                byte |= code << (2 * k);
            }
        }
        w |= byte << (8 * q);
    }
    bed[(t * Mg_pad + g) * 32 + l] = w;
}

// ------------------------------------------------------------------------------------------------
static void free_layout(gvb_ctx* c) {
    auto fr = [](auto*& p) { if (p) { cudaFree(p); p = nullptr; } };
    fr(c->bed); fr(c->maskw); fr(c->validw); fr(c->mave); fr(c->msig); fr(c->counts);
    fr(c->tmpN); fr(c->tmpN2); fr(c->tmpM); fr(c->tmpM2); fr(c->wv); fr(c->cv);
    fr(c->ax_partial); c->ax_partial_cap = 0;
    fr(c->tab_u); c->tab_u_cap = 0;
    fr(c->tab_v); c->tab_v_cap = 0;
    fr(c->shift_u); fr(c->shift_v); c->shift_cap = 0;
    fr(c->acc_i64); c->acc_i64_cap = 0;
    gvb_misslist_reset(c);
    gvb_twin_reset(c);
    c->total_missing = 0;
    c->have_mask = c->have_stats = false;
}

int gvb_layout_alloc(gvb_ctx* c, long N, long Mt, long S, long M) {
    GVB_ARG(N > 0 && M > 0 && Mt >= M && S >= 0 && S + M <= Mt, "N, M, Mt, S");
    GVB_CUDA(cudaSetDevice(c->device));
    free_layout(c);
    c->N = N; c->Mt = Mt; c->S = S; c->M = M;
    c->mbytes = (N + 3) / 4;
    c->n_stripes = (c->mbytes + GVB_STRIPE_POS - 1) / GVB_STRIPE_POS;
    c->Npad = c->n_stripes * GVB_STRIPE_IND;
    c->Mg = (M + 3) / 4;
    c->Mg_pad = gvb_roundup(c->Mg, GVB_GROUP_TILE);
    c->bed_words = (size_t)c->n_stripes * (size_t)c->Mg_pad * 32;
    // 64 KB of slack behind the last stripe: the X^T.u kernel's last marker block may read past Mg_pad
    const size_t slack_words = 16384;
    cudaError_t e = gvb_malloc(c, &c->bed, (c->bed_words + slack_words) * sizeof(uint32_t));
    if (e != cudaSuccess) {
        gvb_set_error("cannot allocate %.3f GB of HBM for the packed genotype matrix: %s", c->bed_words * 4.0 / 1e9, cudaGetErrorString(e));
        return GVB_ERR_NOMEM;
    }
    GVB_CUDA(cudaMemsetAsync(c->bed, 0x55, (c->bed_words + slack_words) * sizeof(uint32_t), c->stream));
    size_t npos = (size_t)c->n_stripes * 32;
    size_t mp = (size_t)c->Mg_pad * 4;
    GVB_CUDA(gvb_malloc(c, &c->maskw, npos * 4));
    GVB_CUDA(gvb_malloc(c, &c->validw, npos * 4));
    GVB_CUDA(gvb_malloc(c, &c->mave, mp * 8));
    GVB_CUDA(gvb_malloc(c, &c->msig, mp * 8));
    GVB_CUDA(gvb_malloc(c, &c->counts, mp * 8 * sizeof(int64_t)));
    GVB_CUDA(gvb_malloc(c, &c->tmpN, c->Npad * 8));
    GVB_CUDA(gvb_malloc(c, &c->tmpN2, c->Npad * 8));
    GVB_CUDA(gvb_malloc(c, &c->tmpM, mp * 8));
    GVB_CUDA(gvb_malloc(c, &c->tmpM2, mp * 8));
    GVB_CUDA(gvb_malloc(c, &c->wv, mp * 8));
    GVB_CUDA(gvb_malloc(c, &c->cv, mp * 8));
    GVB_CUDA(cudaMemsetAsync(c->mave, 0, mp * 8, c->stream));
    GVB_CUDA(cudaMemsetAsync(c->msig, 0, mp * 8, c->stream));
    GVB_CUDA(cudaMemsetAsync(c->tmpN, 0, c->Npad * 8, c->stream));
    GVB_CUDA(cudaMemsetAsync(c->tmpN2, 0, c->Npad * 8, c->stream));
    GVB_CUDA(cudaMemsetAsync(c->tmpM, 0, mp * 8, c->stream));
    GVB_CUDA(cudaMemsetAsync(c->tmpM2, 0, mp * 8, c->stream));
    // default mask: everybody present (data.cpp:86-100)
    return gvb_set_mask(c, nullptr, (int)N);
}

int gvb_layout_retile(gvb_ctx* c, const uint8_t* d_snp, long j0, long nmark) {
    long ngroups = (nmark + 3) / 4;
    long total = c->n_stripes * ngroups * 32;
    int threads = 256;
    long blocks = (total + threads - 1) / threads;
    retile_kernel<<<(unsigned)blocks, threads, 0, c->stream>>>(d_snp, nmark, c->mbytes, j0 / 4, c->M, j0, c->bed, c->Mg_pad, c->n_stripes);
    GVB_LAUNCHED(c);
    return GVB_OK;
}

// markers per upload chunk: multiple of 4, about 256 MB
static long chunk_markers(const gvb_ctx* c) {
    long m = (long)((256ll << 20) / c->mbytes);
    m = std::max(4l, m / 4 * 4);
    return std::min(m, gvb_roundup(c->M, 4));
}

extern "C" int gvb_bed_load_host(gvb_ctx* c, const uint8_t* bed, long N, long Mt, long S, long M) {
    GVB_ARG(c && bed, "ctx / bed");
    GVB_CHECK(gvb_layout_alloc(c, N, Mt, S, M));
    long cm = chunk_markers(c);
    uint8_t* d_stage = nullptr;
    GVB_CUDA(gvb_malloc(c, &d_stage, (size_t)cm * c->mbytes));
    for (long j0 = 0; j0 < M; j0 += cm) {
        long n = std::min(cm, M - j0);
        GVB_CUDA(cudaMemcpyAsync(d_stage, bed + (size_t)j0 * c->mbytes, (size_t)n * c->mbytes, cudaMemcpyHostToDevice, c->stream));
        int rc = gvb_layout_retile(c, d_stage, j0, n);
        if (rc != GVB_OK) { cudaFree(d_stage); return rc; }
    }
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_stage);
    return GVB_OK;
}

extern "C" int gvb_bed_load_file(gvb_ctx* c, const char* path, long N, long Mt, long S, long M) {
    GVB_ARG(c && path, "ctx / path");
    int fd = open(path, O_RDONLY);
    if (fd < 0) {
        gvb_set_error("could not open bed file: %s", path);
        return GVB_ERR_IO;
    }
    int rc = gvb_layout_alloc(c, N, Mt, S, M);
    if (rc != GVB_OK) { close(fd); return rc; }
    long cm = chunk_markers(c);
    size_t chunk_bytes = (size_t)cm * c->mbytes;
    uint8_t* h_stage[2] = {nullptr, nullptr};
    uint8_t* d_stage[2] = {nullptr, nullptr};
    cudaEvent_t done[2];
    for (int b = 0; b < 2; b++) {
        if (cudaMallocHost(&h_stage[b], chunk_bytes) != cudaSuccess || gvb_malloc(c, &d_stage[b], chunk_bytes) != cudaSuccess) {
            gvb_set_error("cannot allocate staging buffers for the bed upload");
            close(fd);
            return GVB_ERR_NOMEM;
        }
        cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming);
    }
    // file offset of the shard: 3 magic bytes + S columns (data.cpp:215)
    off_t base = (off_t)3 + (off_t)S * (off_t)c->mbytes;
    int buf = 0;
    rc = GVB_OK;
    for (long j0 = 0; j0 < M && rc == GVB_OK; j0 += cm, buf ^= 1) {
        long n = std::min(cm, M - j0);
        size_t want = (size_t)n * c->mbytes, got = 0;
        cudaEventSynchronize(done[buf]);   // the previous use of this staging pair has been consumed
        while (got < want) {
            ssize_t r = pread(fd, h_stage[buf] + got, want - got, base + (off_t)j0 * (off_t)c->mbytes + (off_t)got);
            if (r <= 0) break;
            got += (size_t)r;
        }
        if (got != want) {
            gvb_set_error("short read on bed file %s: wanted %zu bytes at marker %ld, got %zu", path, want, S + j0, got);
            rc = GVB_ERR_IO;
            break;
        }
        if (cudaMemcpyAsync(d_stage[buf], h_stage[buf], want, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { rc = GVB_ERR_CUDA; gvb_set_error("H2D copy failed"); break; }
        rc = gvb_layout_retile(c, d_stage[buf], j0, n);
        cudaEventRecord(done[buf], c->stream);
    }
    cudaStreamSynchronize(c->stream);
    for (int b = 0; b < 2; b++) { cudaFreeHost(h_stage[b]); cudaFree(d_stage[b]); cudaEventDestroy(done[b]); }
    close(fd);
    return rc;
}

extern "C" int gvb_bed_synth(gvb_ctx* c, uint64_t seed, long N, long Mt, long S, long M, double miss_rate) {
    GVB_ARG(c, "ctx");
    GVB_ARG(miss_rate >= 0.0 && miss_rate < 1.0, "miss_rate");
    GVB_CHECK(gvb_layout_alloc(c, N, Mt, S, M));
    unsigned thr = (unsigned)(miss_rate * 65536.0 + 0.5);
    long total = c->n_stripes * c->Mg * 32;   // one thread per 32-bit word; < 2^31 blocks even at 840 GB / 8
    int threads = 256;
    long blocks = (total + threads - 1) / threads;
    synth_kernel<<<(unsigned)blocks, threads, 0, c->stream>>>(c->bed, seed, N, c->mbytes, S, M, c->Mg, c->Mg_pad, c->n_stripes, thr);
    GVB_LAUNCHED(c);
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    return GVB_OK;
}

extern "C" int gvb_bed_decode(gvb_ctx* c, long j0, long n, uint8_t* out) {
    GVB_ARG(c && c->bed && out, "ctx / matrix not loaded / out");
    GVB_ARG(j0 >= 0 && n > 0 && j0 + n <= c->M, "marker range");
    uint8_t* d_out = nullptr;
    size_t bytes = (size_t)n * c->mbytes;
    GVB_CUDA(gvb_malloc(c, &d_out, bytes));
    int threads = 256;
    long blocks = ((long)bytes + threads - 1) / threads;
    decode_kernel<<<(unsigned)blocks, threads, 0, c->stream>>>(c->bed, c->Mg_pad, c->mbytes, j0, n, d_out);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_out);
    if (e != cudaSuccess) {
        gvb_set_error("decode failed: %s", cudaGetErrorString(e));
        return GVB_ERR_CUDA;
    }
    return GVB_OK;
}

extern "C" int gvb_set_mask(gvb_ctx* c, const uint8_t* mask4, int nonas) {
    GVB_ARG(c && c->maskw, "ctx / matrix not allocated");
    size_t npos = (size_t)c->n_stripes * 32;
    std::vector<uint32_t> mw(npos, 0u), vw(npos, 0u);
    long present = 0;
    for (long p = 0; p < c->mbytes; p++) {
        unsigned valid = 0;
        for (int k = 0; k < 4; k++)
            if (4 * p + k < c->N) valid |= 1u << k;
        unsigned m = mask4 ? (mask4[p] & 0xFu & valid) : valid;
        present += __builtin_popcount(m);
        auto spread = [](unsigned nib) { return ((nib & 1u) | ((nib & 2u) << 1) | ((nib & 4u) << 2) | ((nib & 8u) << 3)) * 0x01010101u; };
        mw[p] = spread(m);
        vw[p] = spread(valid);
    }
    GVB_CUDA(cudaMemcpyAsync(c->maskw, mw.data(), npos * 4, cudaMemcpyHostToDevice, c->stream));
    GVB_CUDA(cudaMemcpyAsync(c->validw, vw.data(), npos * 4, cudaMemcpyHostToDevice, c->stream));
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    c->nonas = nonas;
    c->mask_present = present;
    c->have_mask = true;
    c->have_stats = false;
    c->layout_gen++;   // by-products of earlier solves (A mu, A^T A mu, A^T y) describe another operator now
    c->cg_prepared = false;   // ... and so does a prepared solve or a companion that was waiting for one
    c->cg_companion = nullptr;
    return GVB_OK;
}
