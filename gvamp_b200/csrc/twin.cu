// twin.cu -- the individual-major twin of the packed genotype matrix.
//
// X^T.u wants byte q of a word to be the PLINK byte of marker 4g+q (four individuals of ONE marker): that byte is its
// table index as it stands.  X.v (reference data::Ax, data.cpp:848-1011) wants the transposed block: the codes of the four
// markers of ONE individual, which it gathers from the four bytes with a mask and a multiply per individual -- the
// instructions that make the alu pipe the limiter of ax_tile_kernel.  A 4x4 block of 2-bit codes cannot be
// byte-aligned both ways at once, but HBM can hold both orientations for every shard up to ~85 GB: the twin stores, at the
// SAME word index, the transposed block
//
//      byte k of twin word = c(marker 4g, ind k) | c(4g+1, k) << 2 | c(4g+2, k) << 4 | c(4g+3, k) << 6
//
// i.e. exactly the index the gather computes, so X.v's walk becomes PRMT + LDS per lookup like X^T.u's.  Tables, sums and
// results are bit-identical (tests/test_gpu_kernels.py::test_twin_layout_is_bit_identical); padding 0x55 maps to itself.
// The twin is built by the first X.v after the statistics when spare HBM allows it and is never required: shards that fill
// the part (config 5: 105 GB) run X.v on the one matrix.  Env GVB_TWIN=0 disables it, GVB_TWIN=1 skips the head-room rule.
#include <stdlib.h>

#include <algorithm>

#include "gvb_internal.cuh"

namespace {

__global__ void __launch_bounds__(256) twin_kernel(const uint4* __restrict__ bed, uint4* __restrict__ twin, size_t n_vec) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_vec; i += stride) {
        const uint4 in = bed[i];
        const uint32_t w[4] = {in.x, in.y, in.z, in.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t t = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) t |= (((w[j] & (0x03030303u << (2 * k))) * (0x01041040u >> (2 * k))) >> 24) << (8 * k);
            o[j] = t;
        }
        twin[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

}   // namespace

void gvb_twin_reset(gvb_ctx* c) {
    if (c->bed_twin) cudaFree(c->bed_twin);
    c->bed_twin = nullptr;
    c->twin_state = 0;
    c->twin_stripes = 0;
}

// Builds the twin once per matrix.  GVB_OK also when it is not built (twin_state = -1: X.v keeps gathering from `bed`).
int gvb_twin_build(gvb_ctx* c) {
    if (c->twin_state != 0) return GVB_OK;
    const char* mode = getenv("GVB_TWIN");
    auto skip = [&](const char* why) {
        cudaGetLastError();
        c->twin_state = -1;
        if (getenv("GVB_VERBOSE")) fprintf(stderr, "[gvamp_b200] individual-major twin not built (%s): X.v gathers from the one matrix\n", why);
        return GVB_OK;
    };
    if (mode && !strcmp(mode, "0")) return skip("GVB_TWIN=0");
    // the missing-genotype list of X^T.u (up to half the bed) has the first call on spare HBM
    if (c->total_missing > 0 && c->miss_state == 0) GVB_CHECK(gvb_misslist_build(c));
    const size_t slack_words = 16384;   // as gvb_layout_alloc
    const size_t stripe_words = (size_t)c->Mg_pad * 32;
    long stripes = c->n_stripes;
    if (const char* e = getenv("GVB_TWIN_STRIPES")) {   // tests: a partial twin of exactly this many stripes
        stripes = std::max(0l, std::min((long)atol(e), c->n_stripes));
    } else if (!(mode && !strcmp(mode, "1"))) {
        // head room kept free for what is allocated later (lookup tables, accumulators, solver vectors, a second context)
        size_t free_b = 0, total_b = 0;
        GVB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const size_t reserve = ((size_t)4 << 30) + 300 * (size_t)c->Mg_pad * 4 + 64 * (size_t)c->Npad;   // + X.v tables and ~30 M-vectors
        const size_t avail = free_b > reserve + slack_words * 4 ? free_b - reserve - slack_words * 4 : 0;
        stripes = std::min((long)(avail / (stripe_words * sizeof(uint32_t))), c->n_stripes);
        // A PARTIAL twin when the whole one does not fit: the first `stripes` stripes are walked on the twin, the rest gathers from
        // the one matrix (two launches per X.v).  Worth its second launch from ~5 % of the stripes on.
        if (stripes < c->n_stripes) stripes = stripes / 60 * 60;   // whole CTAs of 15 (or 12) stripes
        if (stripes < c->n_stripes / 20) return skip("not enough spare HBM");
    }
    if (stripes <= 0) return skip("no stripes");
    const size_t words = (size_t)stripes * stripe_words;
    const size_t bytes = (words + slack_words) * sizeof(uint32_t);
    if (cudaMalloc(&c->bed_twin, bytes) != cudaSuccess) return skip("allocation failed");
    GVB_CUDA(cudaMemsetAsync(c->bed_twin + words, 0x55, slack_words * sizeof(uint32_t), c->stream));
    const size_t n_vec = words / 4;   // a stripe is a multiple of 1024 words
    twin_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(reinterpret_cast<const uint4*>(c->bed), reinterpret_cast<uint4*>(c->bed_twin), n_vec);
    GVB_LAUNCHED(c);
    c->twin_stripes = stripes;
    c->twin_state = stripes == c->n_stripes ? 1 : 2;
    if (getenv("GVB_VERBOSE")) fprintf(stderr, "[gvamp_b200] individual-major twin for %ld of %ld stripes\n", stripes, c->n_stripes);
    return GVB_OK;
}

extern "C" int gvb_twin_state(gvb_ctx* c) { return c ? c->twin_state : 0; }
extern "C" int gvb_twin_release(gvb_ctx* c) {
    GVB_ARG(c, "ctx");
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    const bool had = c->bed_twin != nullptr;
    gvb_twin_reset(c);
    if (had) c->twin_state = -1;
    return GVB_OK;
}
extern "C" long gvb_twin_stripes(gvb_ctx* c) { return c ? c->twin_stripes : 0; }
