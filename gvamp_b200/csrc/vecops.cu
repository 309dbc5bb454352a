// vecops.cu -- FP64 vector kernels of the VAMP loop: fused dot products / norms with a deterministic
// two-stage reduction and ONE small NCCL allreduce, the Gaussian-mixture denoiser (g1, g1d), the EM
// sufficient statistics of the prior update and the probit z-denoiser.
//
// Reference: utilities.cpp:190-214 (inner_prod, l2_norm2), vamp.cpp:805-869 (g1, g1d),
// vamp.cpp:929-1013 (updatePrior E-step), vamp_probit.cpp:661-726 + utilities.cpp:345-409 (probit).
#include "gvb_internal.cuh"

// ------------------------------------------------------------------------------------------------
// block-level helper: reduce K per-thread values, thread 0 writes them to partial[blockIdx.x*K + k]
// ------------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void block_reduce_store(double (&v)[K], double* __restrict__ partial) {
    __shared__ double sm[32][K];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < K; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if (lane == 0) sm[warp][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double s = 0.0;
        for (int w = 0; w < nw; w++) s += sm[w][threadIdx.x];
        partial[blockIdx.x * K + threadIdx.x] = s;
    }
    __syncthreads();
}

// one warp per scalar: lane l adds partials l, l+32, ... in order, then a fixed shuffle tree: deterministic, and a chain of
// nblocks/32 dependent FP64 additions instead of nblocks (this launch sits on the host-visible path of every CG iteration)
__global__ void __launch_bounds__(1024) reduce_final_kernel(const double* __restrict__ partial, int nblocks, int K, double* __restrict__ result) {
    const int lane = threadIdx.x & 31;
    for (int k = threadIdx.x >> 5; k < K; k += blockDim.x >> 5) {
        double s = 0.0;
        for (int b = lane; b < nblocks; b += 32) s += partial[b * K + k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) result[k] = s;
    }
}

// device part of a reduction: c->red_result[0..K) <- sum over blocks (and over ranks if sync) of the per-block partials; nothing
// returns to the host (the CG driver of cg.cu consumes the scalars in its own kernels)
int gvb_reduce_device(gvb_ctx* c, int nblocks, int K, bool sync) {
    reduce_final_kernel<<<1, 32 * std::min(K, 32), 0, c->stream>>>(c->red_partial, nblocks, K, c->red_result);
    GVB_LAUNCHED(c);
    if (sync && c->nranks > 1) {
        // replaces the per-scalar MPI_Allreduce calls (utilities.cpp:203, vamp.cpp:313,990,1012-1013)
        GVB_NCCL(ncclAllReduce(c->red_result, c->red_result, K, ncclDouble, ncclSum, c->comm, c->stream));
    }
    return GVB_OK;
}

int gvb_reduce_finish(gvb_ctx* c, int nblocks, int K, bool sync, double* res_host) {
    GVB_CHECK(gvb_reduce_device(c, nblocks, K, sync));
    GVB_CUDA(cudaMemcpyAsync(c->h_red, c->red_result, K * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    c->host_syncs++;
    for (int k = 0; k < K; k++) res_host[k] = c->h_red[k];
    return GVB_OK;
}

static inline int red_blocks(const gvb_ctx* c, long n) {
    long b = (n + 511) / 512;
    return (int)std::max(1l, std::min(b, (long)GVB_RED_BLOCKS));
}

// ------------------------------------------------------------------------------------------------
// axpby, dots, dist2
// ------------------------------------------------------------------------------------------------
// out may alias x or y (in-place damping), hence no __restrict__
__global__ void axpby_kernel(double* out, double a, const double* x, double b, const double* y, long n) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        out[i] = y ? a * x[i] + b * y[i] : a * x[i];
}

extern "C" int gvb_vec_axpby(gvb_ctx* c, gvb_vec out, double a, gvb_vec x, double b, gvb_vec y) {
    GVB_ARG(c && out && x && out->n == x->n && (!y || y->n == x->n), "vector lengths");
    long n = x->n;
    int blocks = (int)std::min((n + 255) / 256, (long)c->sm_count * 8);
    axpby_kernel<<<blocks, 256, 0, c->stream>>>(out->d, a, x->d, b, y ? y->d : nullptr, n);
    GVB_LAUNCHED(c);
    return GVB_OK;
}

__global__ void axpby_div_kernel(double* out, double a, const double* x, double b, const double* y, double div, long n) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) out[i] = (a * x[i] + b * y[i]) / div;
}

extern "C" int gvb_vec_axpby_div(gvb_ctx* c, gvb_vec out, double a, gvb_vec x, double b, gvb_vec y, double div) {
    GVB_ARG(c && out && x && y && out->n == x->n && y->n == x->n, "vector lengths");
    long n = x->n;
    int blocks = (int)std::min((n + 255) / 256, (long)c->sm_count * 8);
    axpby_div_kernel<<<blocks, 256, 0, c->stream>>>(out->d, a, x->d, b, y->d, div, n);
    GVB_LAUNCHED(c);
    return GVB_OK;
}

struct DotArgs {
    const double* x[8];
    const double* y[8];
};

template <int K>
__global__ void __launch_bounds__(256) dots_kernel(DotArgs a, long n, double* __restrict__ partial) {
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; k++) acc[k] = 0.0;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < K; k++) acc[k] += a.x[k][i] * a.y[k][i];
    }
    block_reduce_store<K>(acc, partial);
}

// launches the fused dot products; the partials land in c->red_partial, *blocks_out tells the reduction how many
static int vec_dots_launch(gvb_ctx* c, int n, const gvb_vec* x, const gvb_vec* y, int* blocks_out) {
    GVB_ARG(c && n >= 1 && n <= 8 && x, "1..8 dot products");
    DotArgs a;
    long len = x[0]->n;
    for (int k = 0; k < 8; k++) {
        int kk = k < n ? k : 0;
        GVB_ARG(x[kk] && x[kk]->n == len && (!y || !y[kk] || y[kk]->n == len), "vector lengths");
        a.x[k] = x[kk]->d;
        a.y[k] = (y && y[kk]) ? y[kk]->d : x[kk]->d;
    }
    int blocks = red_blocks(c, len);
    switch (n) {
        case 1: dots_kernel<1><<<blocks, 256, 0, c->stream>>>(a, len, c->red_partial); break;
        case 2: dots_kernel<2><<<blocks, 256, 0, c->stream>>>(a, len, c->red_partial); break;
        case 3: dots_kernel<3><<<blocks, 256, 0, c->stream>>>(a, len, c->red_partial); break;
        case 4: dots_kernel<4><<<blocks, 256, 0, c->stream>>>(a, len, c->red_partial); break;
        case 5: dots_kernel<5><<<blocks, 256, 0, c->stream>>>(a, len, c->red_partial); break;
        case 6: dots_kernel<6><<<blocks, 256, 0, c->stream>>>(a, len, c->red_partial); break;
        case 7: dots_kernel<7><<<blocks, 256, 0, c->stream>>>(a, len, c->red_partial); break;
        default: dots_kernel<8><<<blocks, 256, 0, c->stream>>>(a, len, c->red_partial); break;
    }
    GVB_LAUNCHED(c);
    *blocks_out = blocks;
    return GVB_OK;
}

extern "C" int gvb_vec_dots(gvb_ctx* c, int n, const gvb_vec* x, const gvb_vec* y, int sync, double* res) {
    GVB_ARG(res, "result pointer");
    int blocks = 0;
    GVB_CHECK(vec_dots_launch(c, n, x, y, &blocks));
    return gvb_reduce_finish(c, blocks, n, sync != 0, res);
}

// the same dot products, results left in c->red_result on the device
int gvb_vec_dots_device(gvb_ctx* c, int n, const gvb_vec* x, const gvb_vec* y, bool sync) {
    int blocks = 0;
    GVB_CHECK(vec_dots_launch(c, n, x, y, &blocks));
    return gvb_reduce_device(c, blocks, n, sync);
}

__global__ void __launch_bounds__(256) dist2_kernel(const double* __restrict__ x, const double* __restrict__ y, long n, double* __restrict__ partial) {
    double acc[1] = {0.0};
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        double d = x[i] - y[i];
        acc[0] += d * d;
    }
    block_reduce_store<1>(acc, partial);
}

extern "C" int gvb_vec_dist2(gvb_ctx* c, gvb_vec x, gvb_vec y, int sync, double* res) {
    GVB_ARG(c && x && y && x->n == y->n && res, "vector lengths");
    int blocks = red_blocks(c, x->n);
    dist2_kernel<<<blocks, 256, 0, c->stream>>>(x->d, y->d, x->n, c->red_partial);
    GVB_LAUNCHED(c);
    return gvb_reduce_finish(c, blocks, 1, sync != 0, res);
}

// ------------------------------------------------------------------------------------------------
// batched reductions: up to GVB_RED_BATCH dot products / squared norms of linear combinations over vectors of ANY lengths
// (M- and N-vectors mixed) in two launches and ONE host synchronisation.  The VAMP loop's print-only diagnostics and its
// end-of-iteration tests (vamp.cpp:388-392, 741-749, 892-927, 1232-1317: a dozen separate allreduces in the reference)
// travel as two such batches per iteration.
// ------------------------------------------------------------------------------------------------
struct BatchArgs {
    const double* x[GVB_RED_BATCH];
    const double* y[GVB_RED_BATCH];
    double a[GVB_RED_BATCH], b[GVB_RED_BATCH];
    long n[GVB_RED_BATCH];
    int kind[GVB_RED_BATCH], nblk[GVB_RED_BATCH], off[GVB_RED_BATCH];
};

__global__ void __launch_bounds__(256) batch_partial_kernel(BatchArgs A, double* __restrict__ partial) {
    const int op = blockIdx.y;
    if ((int)blockIdx.x >= A.nblk[op]) return;
    const double* __restrict__ x = A.x[op];
    const double* __restrict__ y = A.y[op];
    const double a = A.a[op], b = A.b[op];
    const long n = A.n[op], stride = (long)A.nblk[op] * 256;
    double acc[1] = {0.0};
    if (A.kind[op] == GVB_RED_DOT) {
        for (long i = blockIdx.x * 256l + threadIdx.x; i < n; i += stride) acc[0] += x[i] * y[i];
    } else {
        for (long i = blockIdx.x * 256l + threadIdx.x; i < n; i += stride) {
            const double d = y ? a * x[i] + b * y[i] : a * x[i];
            acc[0] += d * d;
        }
    }
    __shared__ double sm[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], o);
    if (lane == 0) sm[warp] = acc[0];
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += sm[w];
        partial[A.off[op] + blockIdx.x] = t;
    }
}

__global__ void batch_final_kernel(BatchArgs A, int nops, const double* __restrict__ partial, double* __restrict__ result) {
    const int lane = threadIdx.x & 31, op = threadIdx.x >> 5;
    if (op >= nops) return;
    double t = 0.0;
    for (int b = lane; b < A.nblk[op]; b += 32) t += partial[A.off[op] + b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) result[op] = t;
}

extern "C" int gvb_vec_reduce_batch(gvb_ctx* c, int nops, const gvb_red_op* ops, double* res) {
    GVB_ARG(c && ops && res && nops >= 1 && nops <= GVB_RED_BATCH, "1..GVB_RED_BATCH operations");
    BatchArgs A;
    int off = 0, maxblk = 1, nsync = 0;
    for (int k = 0; k < nops; k++) {
        const gvb_red_op& o = ops[k];
        GVB_ARG(o.x && (o.kind == GVB_RED_DOT || o.kind == GVB_RED_SQ), "operation");
        GVB_ARG(!o.y || o.y->n == o.x->n, "vector lengths");
        GVB_ARG(!(o.sync && k > nsync), "rank-summed operations come first in a batch");
        if (o.sync) nsync = k + 1;
        A.x[k] = o.x->d;
        A.y[k] = o.y ? o.y->d : (o.kind == GVB_RED_DOT ? o.x->d : nullptr);
        A.a[k] = o.a;
        A.b[k] = o.b;
        A.n[k] = o.x->n;
        A.kind[k] = o.kind;
        A.nblk[k] = red_blocks(c, o.x->n);
        A.off[k] = off;
        off += A.nblk[k];
        maxblk = std::max(maxblk, A.nblk[k]);
    }
    for (int k = nops; k < GVB_RED_BATCH; k++) { A.x[k] = A.y[k] = nullptr; A.a[k] = A.b[k] = 0; A.n[k] = 0; A.kind[k] = 0; A.nblk[k] = 0; A.off[k] = 0; }
    batch_partial_kernel<<<dim3((unsigned)maxblk, (unsigned)nops), 256, 0, c->stream>>>(A, c->red_partial);
    GVB_LAUNCHED(c);
    batch_final_kernel<<<1, 32 * nops, 0, c->stream>>>(A, nops, c->red_partial, c->red_result);
    GVB_LAUNCHED(c);
    if (nsync > 0 && c->nranks > 1) GVB_NCCL(ncclAllReduce(c->red_result, c->red_result, nsync, ncclDouble, ncclSum, c->comm, c->stream));
    GVB_CUDA(cudaMemcpyAsync(c->h_red, c->red_result, nops * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    c->host_syncs++;
    for (int k = 0; k < nops; k++) res[k] = c->h_red[k];
    return GVB_OK;
}

// ------------------------------------------------------------------------------------------------
// Gaussian-mixture denoiser, vamp.cpp:805-869
// ------------------------------------------------------------------------------------------------
struct MixArgs {
    int L;
    double sigma;     // 1/gam1
    double eta_max;   // max variance
    double probs[GVB_MAX_MIX];
    double vars[GVB_MAX_MIX];
};

__global__ void __launch_bounds__(256) denoise_kernel(const double* __restrict__ r1, long n, MixArgs a, double* __restrict__ x1,
                                                      double* __restrict__ partial) {
    double acc[2] = {0.0, 0.0};
    const bool degenerate = (a.sigma < 1e-10 && a.sigma > -1e-10);   // vamp.cpp:813, :844
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        double y = r1[i];
        double val, der;
        if (degenerate) {
            val = y;
            der = 1.0;
        } else {
            double pk = 0.0, pkd = 0.0, pkdd = 0.0;
            for (int k = 0; k < a.L; k++) {
                double vs = a.vars[k] + a.sigma;
                double expe = -0.5 * (y * y) * (a.eta_max - a.vars[k]) / vs / (a.eta_max + a.sigma);
                double ex = exp(expe);
                double z = a.probs[k] / sqrt(vs) * ex;
                pk += z;
                z = z / vs * y;
                pkd -= z;
                double z2 = z / vs * y;
                pkdd = pkdd - a.probs[k] / (vs * sqrt(vs)) * ex + z2;
            }
            double ratio = pkd / pk;
            val = y + a.sigma * ratio;
            der = 1.0 + a.sigma * (pkdd / pk - ratio * ratio);
        }
        x1[i] = val;
        acc[0] += der;
        double d = val - y;
        acc[1] += d * d;
    }
    block_reduce_store<2>(acc, partial);
}

extern "C" int gvb_denoise(gvb_ctx* c, gvb_vec r1, double gam1, const double* probs, const double* vars, int L, gvb_vec x1_hat, double* sums) {
    GVB_ARG(c && r1 && x1_hat && r1->n == x1_hat->n && probs && vars && sums, "arguments");
    GVB_ARG(L >= 1 && L <= GVB_MAX_MIX, "1 <= L <= GVB_MAX_MIX");
    MixArgs a;
    a.L = L;
    a.sigma = 1.0 / gam1;
    a.eta_max = vars[0];
    for (int k = 0; k < L; k++) {
        a.probs[k] = probs[k];
        a.vars[k] = vars[k];
        if (vars[k] > a.eta_max) a.eta_max = vars[k];
    }
    int blocks = red_blocks(c, r1->n);
    denoise_kernel<<<blocks, 256, 0, c->stream>>>(r1->d, r1->n, a, x1_hat->d, c->red_partial);
    GVB_LAUNCHED(c);
    return gvb_reduce_finish(c, blocks, 2, true, sums);
}

// ------------------------------------------------------------------------------------------------
// EM sufficient statistics, vamp.cpp:944-1013
// ------------------------------------------------------------------------------------------------
struct EmArgs {
    int L;
    double gam1, noise_var, lambda, max_sigma;
    double omegas[GVB_MAX_MIX];
    double vars[GVB_MAX_MIX];
};

__global__ void __launch_bounds__(128) em_stats_kernel(const double* __restrict__ r1, long n, EmArgs a, double* __restrict__ partial) {
    const int K = 2 * a.L - 1;
    double acc[2 * GVB_MAX_MIX];
    for (int k = 0; k < K; k++) acc[k] = 0.0;
    const double sqrt2pi = sqrt(2.0 * M_PI);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        double r = r1[i];
        double r2h = r * r / 2.0;
        double num[GVB_MAX_MIX];
        double s = 0.0;
        for (int j = 1; j < a.L; j++) {
            double vn = a.vars[j] + a.noise_var;
            double e = exp(-r2h * (a.max_sigma - a.vars[j]) / vn / (a.max_sigma + a.noise_var));
            double x = a.lambda * a.omegas[j] * e / sqrt(vn) / sqrt2pi;   // vamp.cpp:961
            num[j] = x;
            s += x;
        }
        double pin = 1.0 / (1.0 + (1.0 - a.lambda) / sqrt(2.0 * M_PI * a.noise_var) *
                                      exp(-r2h * a.max_sigma / a.noise_var / (a.noise_var + a.max_sigma)) / s);   // vamp.cpp:979
        acc[0] += pin;
        for (int j = 1; j < a.L; j++) {
            double beta = num[j] / s;
            double m = a.gam1 * r / (1.0 / a.vars[j] + a.gam1);        // vamp.cpp:963
            double v = 1.0 / (1.0 / a.vars[j] + a.gam1);               // vamp.cpp:984
            acc[j] += beta * pin;                                      // res        (:1006)
            acc[a.L - 1 + j] += beta * (m * m + v) * pin;              // res_gammas (:996,1007)
        }
    }
    // block reduction of K values
    __shared__ double sm[4][2 * GVB_MAX_MIX];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < K; k++) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sm[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < K) partial[blockIdx.x * K + threadIdx.x] = sm[0][threadIdx.x] + sm[1][threadIdx.x] + sm[2][threadIdx.x] + sm[3][threadIdx.x];
}

extern "C" int gvb_em_stats(gvb_ctx* c, gvb_vec r1, double gam1, double lambda, const double* omegas, const double* vars, int L, double* sums) {
    GVB_ARG(c && r1 && omegas && vars && sums, "arguments");
    GVB_ARG(L >= 2 && L <= GVB_MAX_MIX, "2 <= L <= GVB_MAX_MIX");
    EmArgs a;
    a.L = L;
    a.gam1 = gam1;
    a.noise_var = 1.0 / gam1;
    a.lambda = lambda;
    a.max_sigma = vars[0];
    for (int k = 0; k < L; k++) {
        a.omegas[k] = omegas[k];
        a.vars[k] = vars[k];
        if (vars[k] > a.max_sigma) a.max_sigma = vars[k];
    }
    int K = 2 * L - 1;
    int blocks = (int)std::max(1l, std::min((r1->n + 255) / 256, (long)GVB_RED_BLOCKS));
    em_stats_kernel<<<blocks, 128, 0, c->stream>>>(r1->d, r1->n, a, c->red_partial);
    GVB_LAUNCHED(c);
    return gvb_reduce_finish(c, blocks, K, true, sums);
}

// ------------------------------------------------------------------------------------------------
// probit z-denoiser, vamp_probit.cpp:661-726; erfcx after utilities.cpp:345-409
// ------------------------------------------------------------------------------------------------
__device__ double erfcx_dev(double x) {
    double a = fmax(x, 0.0 - x);
    double m = a - 4.0, p = a + 4.0;
    double r = 1.0 / p;
    double q = m * r;
    double t = fma(q + 1.0, -4.0, a);
    double e = fma(q, -a, t);
    q = fma(r, e, q);
    p = 0x1.edcad78fc8044p-31;
    p = fma(p, q, 0x1.b1548f14735d1p-30);
    p = fma(p, q, -0x1.a1ad2e6c4a7a8p-27);
    p = fma(p, q, -0x1.1985b48f08574p-26);
    p = fma(p, q, 0x1.c6a8093ac4f83p-24);
    p = fma(p, q, 0x1.31c2b2b44b731p-24);
    p = fma(p, q, -0x1.b87373facb29fp-21);
    p = fma(p, q, 0x1.3fef1358803b7p-22);
    p = fma(p, q, 0x1.7eec072bb0be3p-18);
    p = fma(p, q, -0x1.78a680a741c4ap-17);
    p = fma(p, q, -0x1.9951f39295cf4p-16);
    p = fma(p, q, 0x1.3be1255ce180bp-13);
    p = fma(p, q, -0x1.a1df71176b791p-13);
    p = fma(p, q, -0x1.8d4aaa0099bc8p-11);
    p = fma(p, q, 0x1.49c673066c831p-8);
    p = fma(p, q, -0x1.0962386ea02b7p-6);
    p = fma(p, q, 0x1.3079edf465cc3p-5);
    p = fma(p, q, -0x1.0fb06dfedc4ccp-4);
    p = fma(p, q, 0x1.7fee004e266dfp-4);
    p = fma(p, q, -0x1.9ddb23c3e14d2p-4);
    p = fma(p, q, 0x1.16ecefcfa4865p-4);
    p = fma(p, q, 0x1.f7f5df66fc349p-7);
    p = fma(p, q, -0x1.1df1ad154a27fp-3);
    p = fma(p, q, 0x1.dd2c8b74febf6p-3);
    double d = a + 0.5;
    r = 1.0 / d;
    r = r * 0.5;
    q = fma(p, r, r);
    t = q + q;
    e = (p - q) + fma(t, -a, 1.0);
    r = fma(e, r, q);
    if (a > 0x1.fffffffffffffp1023) r = 0.0;
    if (x < 0.0) {
        double s = x * x;
        d = fma(x, x, -s);
        e = exp(s);
        r = e - r;
        r = fma(e, d + d, r);
        r = r + e;
        if (e > 0x1.fffffffffffffp1023) r = e;
    }
    return r;
}

__global__ void __launch_bounds__(256) probit_denoise_kernel(const double* __restrict__ p1, const double* __restrict__ y,
                                                             const double* __restrict__ mcov, long n, double tau1, double probit_var,
                                                             double* __restrict__ z1, double* __restrict__ partial) {
    double acc[2] = {0.0, 0.0};
    const double s = sqrt(probit_var + 1.0 / tau1);
    const double k2pi = 2.0 / sqrt(2.0 * M_PI);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        double p = p1[i], yy = y[i], mc = mcov ? mcov[i] : 0.0;
        double sg = 2.0 * yy - 1.0;
        double cc = (p + mc) / s;
        double ratio = k2pi / erfcx_dev(-sg * cc / sqrt(2.0));
        double g = p + sg * ratio / tau1 / s;                                        // vamp_probit.cpp:684
        double gd = 1.0 - ratio / (1.0 + tau1 * probit_var) * (sg * cc + ratio);     // vamp_probit.cpp:720
        z1[i] = g;
        acc[0] += gd;
        double d = g - p;
        acc[1] += d * d;
    }
    block_reduce_store<2>(acc, partial);
}

extern "C" int gvb_probit_denoise(gvb_ctx* c, gvb_vec p1, gvb_vec y, gvb_vec mcov, double tau1, double probit_var, gvb_vec z1_hat, double* sums) {
    GVB_ARG(c && p1 && y && z1_hat && sums, "arguments");
    long n = c->N;   // individuals, not the padded length
    GVB_ARG(p1->n >= n && y->n >= n && z1_hat->n >= n && (!mcov || mcov->n >= n), "vector lengths");
    int blocks = red_blocks(c, n);
    probit_denoise_kernel<<<blocks, 256, 0, c->stream>>>(p1->d, y->d, mcov ? mcov->d : nullptr, n, tau1, probit_var, z1_hat->d, c->red_partial);
    GVB_LAUNCHED(c);
    return gvb_reduce_finish(c, blocks, 2, false, sums);   // N-vectors are replicated: no rank sum
}

// ------------------------------------------------------------------------------------------------
// probit covariate effects: the N x C passes of vamp::Newton_method_cov (vamp_probit.cpp:936-1067) with grad_cov
// (:814-839) and mlogL_probit (:841-858).  The reference walks the N x C covariate matrix on one core, once for the
// Hessian + Newton numerator, once for the gradient and once per line-search probe; here ONE launch returns any subset
// of {mlogL, gradient, numerator + Hessian} at a given eta.  Z stays on the device.
//   out[0]            = -1/N sum_i log Phi(s_i g_i / sqrt(pv)),  g_i = gg_i + <Z_i, eta>, s_i = 2 y_i - 1
//   out[1 .. C]       = gradient  -1/N sum_i r_i s_i Z_ij / sqrt(pv),  r_i = 2/sqrt(2 pi) / erfcx(-s_i g_i / sqrt(2 pv))
//   out[1+C .. 2C]    = sum_i Z_ij lam_i,  lam_i = s_i 2/sqrt(2 pi) / erfcx(-s_i g_i / sqrt(2))   (no pv: vamp_probit.cpp:951)
//   out[1+2C ..]      = H_jk = sum_i Z_ij Z_ik lam_i (lam_i + g_i)
// ------------------------------------------------------------------------------------------------
#define COV_CHUNK 128
#define COV_MAXC 32

__global__ void __launch_bounds__(256) probit_cov_kernel(const double* __restrict__ y, const double* __restrict__ gg, const double* __restrict__ Z,
                                                         long n, int C, const double* __restrict__ eta_d, double probit_var, int what,
                                                         double* __restrict__ partial) {
    __shared__ double Zs[COV_CHUNK][COV_MAXC + 1];
    __shared__ double s_lam[COV_CHUNK], s_w[COV_CHUNK], s_gr[COV_CHUNK], s_ll[COV_CHUNK];
    __shared__ double eta[COV_MAXC];
    const int K = 1 + 2 * C + C * C;
    const int tid = threadIdx.x;
    if (tid < C) eta[tid] = eta_d[tid];
    const double k2pi = 2.0 / sqrt(2.0 * M_PI), isd = 1.0 / sqrt(probit_var);
    // thread t owns H entries e = t, t+256, ... (at most 4 for C = 32) and, for t < C, step[t] / grad[t]; thread 0 owns mlogL
    double hacc[4] = {0.0, 0.0, 0.0, 0.0};
    double sacc = 0.0, gacc = 0.0, lacc = 0.0;
    const long n_chunks = (n + COV_CHUNK - 1) / COV_CHUNK;
    for (long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        const long i0 = ch * COV_CHUNK;
        const int cnt = (int)min((long)COV_CHUNK, n - i0);
        __syncthreads();
        for (int e = tid; e < cnt * C; e += 256) Zs[e / C][e % C] = Z[i0 * C + e];   // rows of the chunk are contiguous
        __syncthreads();
        if (tid < cnt) {
            double g = gg ? gg[i0 + tid] : 0.0;
            for (int j = 0; j < C; j++) g += Zs[tid][j] * eta[j];     // inner_prod(Z[i], eta), same order as the reference
            const double sgn = 2.0 * y[i0 + tid] - 1.0;
            const double lam = k2pi / erfcx_dev(-(sgn * g) / sqrt(2.0)) * sgn;
            s_lam[tid] = lam;
            s_w[tid] = lam * (lam + g);
            const double arg = sgn / sqrt(probit_var) * g;
            s_gr[tid] = -(k2pi / erfcx_dev(-arg / sqrt(2.0))) * sgn * isd;
            s_ll[tid] = -log(0.5 * erfc(-arg * M_SQRT1_2));
        }
        __syncthreads();
        if (what & 4) {
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int e = tid + 256 * r;
                if (e < C * C) {
                    const int j = e / C, k = e % C;
                    double a = 0.0;
                    for (int i = 0; i < cnt; i++) a += Zs[i][k] * (Zs[i][j] * s_w[i]);
                    hacc[r] += a;
                }
            }
        }
        if (tid < C) {
            double a = 0.0, b = 0.0;
            for (int i = 0; i < cnt; i++) {
                a += Zs[i][tid] * s_lam[i];
                b += s_gr[i] * Zs[i][tid];
            }
            sacc += a;
            gacc += b;
        }
        if (tid == 255) {
            double a = 0.0;
            for (int i = 0; i < cnt; i++) a += s_ll[i];
            lacc += a;
        }
    }
    double* mine = partial + (size_t)blockIdx.x * K;
    if (tid == 255) mine[0] = lacc;
    if (tid < C) {
        mine[1 + tid] = gacc;
        mine[1 + C + tid] = sacc;
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int e = tid + 256 * r;
        if (e < C * C) mine[1 + 2 * C + e] = hacc[r];
    }
}

__global__ void probit_cov_reduce_kernel(const double* __restrict__ partial, int nblocks, int K, int C, double inv_n, double* __restrict__ out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; b++) s += partial[(size_t)b * K + k];
    out[k] = (k <= C) ? s * inv_n : s;   // mlogL and the gradient are means (vamp_probit.cpp:837, 857)
}

extern "C" int gvb_probit_cov_pass(gvb_ctx* c, gvb_vec y, gvb_vec gg, gvb_vec Z, int C, const double* eta, double probit_var, int what, double* out) {
    GVB_ARG(c && y && Z && eta && out, "arguments");
    GVB_ARG(C >= 1 && C <= COV_MAXC, "1 <= C <= 32 covariates");
    const long n = c->N;
    GVB_ARG(y->n >= n && (!gg || gg->n >= n) && Z->n >= n * C, "vector lengths (Z is N x C, row-major)");
    const int K = 1 + 2 * C + C * C;
    const int blocks = (int)std::max(1l, std::min((n + COV_CHUNK - 1) / COV_CHUNK, 2l * c->sm_count));
    double* scratch = nullptr;
    GVB_CUDA(cudaMallocAsync(&scratch, ((size_t)blocks * K + K + COV_MAXC) * sizeof(double), c->stream));
    double* d_out = scratch + (size_t)blocks * K;
    double* d_eta = d_out + K;
    GVB_CUDA(cudaMemcpyAsync(d_eta, eta, C * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    probit_cov_kernel<<<blocks, 256, 0, c->stream>>>(y->d, gg ? gg->d : nullptr, Z->d, n, C, d_eta, probit_var, what, scratch);
    GVB_LAUNCHED(c);
    probit_cov_reduce_kernel<<<(K + 127) / 128, 128, 0, c->stream>>>(scratch, blocks, K, C, 1.0 / (double)n, d_out);
    GVB_LAUNCHED(c);
    GVB_CUDA(cudaMemcpyAsync(out, d_out, K * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    GVB_CUDA(cudaFreeAsync(scratch, c->stream));
    return GVB_OK;
}

// mcov[i] = <Z_i, eta> (vamp_probit.cpp:132: the covariate part of the liability)
__global__ void probit_cov_apply_kernel(const double* __restrict__ Z, long n, int C, const double* __restrict__ eta, double* __restrict__ mcov) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double g = 0.0;
    for (int j = 0; j < C; j++) g += Z[i * C + j] * eta[j];
    mcov[i] = g;
}

extern "C" int gvb_probit_cov_apply(gvb_ctx* c, gvb_vec Z, int C, const double* eta, gvb_vec mcov) {
    GVB_ARG(c && Z && eta && mcov && C >= 1 && C <= COV_MAXC, "arguments");
    const long n = c->N;
    GVB_ARG(Z->n >= n * C && mcov->n >= n, "vector lengths");
    double* d_eta = nullptr;
    GVB_CUDA(cudaMallocAsync(&d_eta, COV_MAXC * sizeof(double), c->stream));
    GVB_CUDA(cudaMemcpyAsync(d_eta, eta, C * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    probit_cov_apply_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(Z->d, n, C, d_eta, mcov->d);
    GVB_LAUNCHED(c);
    GVB_CUDA(cudaStreamSynchronize(c->stream));   // eta is a host buffer of the caller
    GVB_CUDA(cudaFreeAsync(d_eta, c->stream));
    return GVB_OK;
}
