// matvec_simple.cu -- generation-0 bed mat-vecs: straightforward FP64 decode-multiply-accumulate on
// the striped layout.  They are exact up to FP64 summation order and serve as the on-device
// cross-check for the table kernels in matvec_lut.cu (select with GVB_KERNELS=simple).
//
//   X^T.u : reference data::dot_product + data::ATx, data.cpp:728-835
//   X.v   : reference data::Ax, data.cpp:848-1011 (scalar-branch semantics: phenotype mask applied)
#include "gvb_internal.cuh"

__device__ __forceinline__ double dosage(unsigned code) { return code == 0u ? 2.0 : (code == 2u ? 1.0 : 0.0); }

// one warp per marker group (4 markers); lane l owns byte position l of every stripe
__global__ void __launch_bounds__(256) atx_simple_kernel(const uint32_t* __restrict__ bed, const double* __restrict__ u,
                                                         const double* __restrict__ mave, const double* __restrict__ msig, long Mg, long Mg_pad,
                                                         long n_stripes, double scale, double* __restrict__ out) {
    int lane = threadIdx.x & 31;
    long g = blockIdx.x * (long)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= Mg) return;
    double sa[4] = {0, 0, 0, 0}, sb[4] = {0, 0, 0, 0};
    for (long t = 0; t < n_stripes; t++) {
        uint32_t w = bed[(t * Mg_pad + g) * 32 + lane];
        const double2* up = reinterpret_cast<const double2*>(u + (t * 32 + lane) * 4);
        double2 u01 = up[0], u23 = up[1];
        double uu[4] = {u01.x, u01.y, u23.x, u23.y};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            unsigned byte = (w >> (8 * q)) & 0xFFu;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                unsigned code = (byte >> (2 * k)) & 3u;
                sa[q] += dosage(code) * uu[k];
                sb[q] += (code == 1u ? 0.0 : 1.0) * uu[k];
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sa[q] += __shfl_xor_sync(0xffffffffu, sa[q], o);
            sb[q] += __shfl_xor_sync(0xffffffffu, sb[q], o);
        }
    }
    if (lane < 4) {
        long j = g * 4 + lane;
        double a = lane == 0 ? sa[0] : lane == 1 ? sa[1] : lane == 2 ? sa[2] : sa[3];
        double b = lane == 0 ? sb[0] : lane == 1 ? sb[1] : lane == 2 ? sb[2] : sb[3];
        out[j] = msig[j] * (a - mave[j] * b) * scale;
    }
}

__global__ void scale_v_kernel(const double* __restrict__ v, const double* __restrict__ mave, const double* __restrict__ msig, long Mpad,
                               double* __restrict__ wv, double* __restrict__ cv) {
    long j = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (j >= Mpad) return;
    double w = msig[j] * v[j];
    wv[j] = w;
    cv[j] = mave[j] * w;
}

// grid (n_stripes, n_chunks); 4 warps split the groups of a chunk, lane l owns position l
__global__ void __launch_bounds__(128) ax_simple_kernel(const uint32_t* __restrict__ bed, const double* __restrict__ wv,
                                                        const double* __restrict__ cv, long Mg, long Mg_pad, long groups_per_chunk,
                                                        long Npad, double* __restrict__ partial) {
    __shared__ double red[4][32][4];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long t = blockIdx.x, chunk = blockIdx.y;
    long g_lo = chunk * groups_per_chunk;
    long g_hi = min(g_lo + groups_per_chunk, Mg);
    double acc[4] = {0, 0, 0, 0};
    for (long g = g_lo + warp; g < g_hi; g += 4) {
        uint32_t w = bed[(t * Mg_pad + g) * 32 + lane];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            double wq = wv[g * 4 + q], cq = cv[g * 4 + q];
            double t0 = 2.0 * wq - cq, t2 = wq - cq, t3 = -cq;
            unsigned byte = (w >> (8 * q)) & 0xFFu;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                unsigned code = (byte >> (2 * k)) & 3u;
                acc[k] += code == 0u ? t0 : (code == 2u ? t2 : (code == 3u ? t3 : 0.0));
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) red[warp][lane][k] = acc[k];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            double s = red[0][lane][k] + red[1][lane][k] + red[2][lane][k] + red[3][lane][k];
            partial[chunk * Npad + (t * 32 + lane) * 4 + k] = s;
        }
    }
}

__global__ void ax_finish_kernel(const double* __restrict__ partial, int n_chunks, long Npad, const uint32_t* __restrict__ maskw, double scale,
                                 double* __restrict__ out) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= Npad) return;
    double s = 0.0;
    for (int c = 0; c < n_chunks; c++) s += partial[c * Npad + i];
    uint32_t m = maskw[i >> 2];
    bool present = (m >> (2 * (i & 3))) & 1u;
    out[i] = present ? s * scale : 0.0;
}

int gvb_atx_simple(gvb_ctx* c, const double* u, double* out) {
    int warps = 8;
    long blocks = (c->Mg + warps - 1) / warps;
    double scale = 1.0 / sqrt((double)c->N);
    atx_simple_kernel<<<(unsigned)blocks, warps * 32, 0, c->stream>>>(c->bed, u, c->mave, c->msig, c->Mg, c->Mg_pad, c->n_stripes, scale, out);
    GVB_LAUNCHED(c);
    return GVB_OK;
}

int gvb_ax_simple(gvb_ctx* c, const double* v, double* out) {
    long Mpad = c->Mg_pad * 4;
    scale_v_kernel<<<(unsigned)((Mpad + 255) / 256), 256, 0, c->stream>>>(v, c->mave, c->msig, Mpad, c->wv, c->cv);
    GVB_LAUNCHED(c);
    long target = (long)c->sm_count * 8;
    long n_chunks = (target + c->n_stripes - 1) / c->n_stripes;
    n_chunks = std::max(1l, std::min(n_chunks, (c->Mg + 15) / 16));
    long gpc = (c->Mg + n_chunks - 1) / n_chunks;
    n_chunks = (c->Mg + gpc - 1) / gpc;
    size_t need = (size_t)n_chunks * c->Npad;
    if (need > c->ax_partial_cap) {
        if (c->ax_partial) cudaFree(c->ax_partial);
        c->ax_partial = nullptr;
        GVB_CUDA(cudaMalloc(&c->ax_partial, need * sizeof(double)));
        c->ax_partial_cap = need;
    }
    dim3 grid((unsigned)c->n_stripes, (unsigned)n_chunks);
    ax_simple_kernel<<<grid, 128, 0, c->stream>>>(c->bed, c->wv, c->cv, c->Mg, c->Mg_pad, gpc, c->Npad, c->ax_partial);
    GVB_LAUNCHED(c);
    double scale = 1.0 / sqrt((double)c->N);
    ax_finish_kernel<<<(unsigned)((c->Npad + 255) / 256), 256, 0, c->stream>>>(c->ax_partial, (int)n_chunks, c->Npad, c->maskw, scale, out);
    GVB_LAUNCHED(c);
    return GVB_OK;
}
