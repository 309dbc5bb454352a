// people.cu -- per-individual ("people") statistics and the N-space LMMSE solve of the XXT form of the denoiser
// (--use-XXT-denoiser 1).
//
// Replaces data::compute_people_statistics (reference data.cpp:548-640) and vamp::lmmse_multAAT / CG_solverAAT
// (denoiserXXT.cpp:15-135).  The statistics are three sums over the MARKERS for every individual,
//     S1_i = sum_j value_ij,   numb_i = sum_j b_ij m_i,   S2_i = sum_j value_ij^2,   value_ij = (a_ij - mu_j) sigma_j b_ij m_i,
// i.e. three walks of the X.v kernel with different per-code table values (matvec_tile.cu: ax_code_values), summed over the
// shards like data.cpp:611-613, followed by mave = S1/numb, msig = sqrt((numb-1)/(S2 - numb mave^2)).  The solve is the
// preconditioned CG of denoiserXXT.cpp with the operator tau A A^T + gam2 on N-vectors (replicated on every rank, so its
// scalars need no allreduce) and the per-individual diagonal preconditioner of denoiserXXT.cpp:66-72.
#include <algorithm>
#include <vector>

#include "gvb_internal.cuh"

namespace {

// s1 and mave are the same buffer (in place): no __restrict__ on them
__global__ void people_finalize_kernel(const double* s1, const double* __restrict__ numb, const double* __restrict__ s2,
                                       const uint32_t* __restrict__ maskw, long Npad, double* mave, double* __restrict__ msig) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= Npad) return;
    const bool present = (maskw[i >> 2] >> (2 * (i & 3))) & 1u;
    double m = 0.0, s = 0.0;
    if (present) {   // data.cpp:621-624
        const double n = numb[i];
        m = s1[i] / n;
        s = sqrt((n - 1.0) / (s2[i] - n * m * m));
    }
    mave[i] = m;
    msig[i] = s;
}

// diag_i = tau ((numb_i - 1)/msig_i^2 + mave_i^2 numb_i)/N + gam2, denoiserXXT.cpp:68
__global__ void aat_diag_kernel(const double* __restrict__ mave, const double* __restrict__ msig, const double* __restrict__ numb, long n, double tau,
                                double gam2, double inv_n, double* __restrict__ diag) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    diag[i] = tau * ((numb[i] - 1.0) / msig[i] / msig[i] + mave[i] * mave[i] * numb[i]) * inv_n + gam2;
}

template <int K>
__device__ __forceinline__ void block_sums(double (&v)[K], double* __restrict__ partial) {
    __shared__ double sm[8][K];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if (lane == 0) sm[warp][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double s = 0.0;
        for (int w = 0; w < 8; w++) s += sm[w][threadIdx.x];
        partial[blockIdx.x * K + threadIdx.x] = s;
    }
}

// r = rhs - q ; z = r/diag ; p = z ; partial: <r,z>, ||rhs||^2
__global__ void __launch_bounds__(256) aat_init_kernel(double* __restrict__ r, double* __restrict__ p, const double* __restrict__ rhs,
                                                       const double* __restrict__ q, const double* __restrict__ diag, long n, double* __restrict__ partial) {
    double a[2] = {0.0, 0.0};
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const double b = rhs[i], x = b - q[i], z = x / diag[i];
        r[i] = x;
        p[i] = z;
        a[0] += x * z;
        a[1] += b * b;
    }
    block_sums<2>(a, partial);
}
// d = tau*d + gam2*p (the operator's tail, denoiserXXT.cpp:24-27) ; partial: <d,p>
__global__ void __launch_bounds__(256) aat_combine_kernel(double* __restrict__ d, const double* __restrict__ p, double tau, double gam2, long n,
                                                          double* __restrict__ partial) {
    double a[1] = {0.0};
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const double x = d[i] * tau + gam2 * p[i];
        d[i] = x;
        a[0] += x * p[i];
    }
    block_sums<1>(a, partial);
}
// mu += alpha p ; r -= alpha d ; partial: <r, r/diag>, ||r||^2, ||mu||^2, ||r/diag||^2
__global__ void __launch_bounds__(256) aat_update_kernel(double* __restrict__ mu, double* __restrict__ r, const double* __restrict__ p,
                                                         const double* __restrict__ d, const double* __restrict__ diag, double alpha, long n,
                                                         double* __restrict__ partial) {
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const double m = mu[i] + alpha * p[i];
        const double x = r[i] - d[i] * alpha;
        const double z = x / diag[i];
        mu[i] = m;
        r[i] = x;
        a[0] += x * z;
        a[1] += x * x;
        a[2] += m * m;
        a[3] += z * z;
    }
    block_sums<4>(a, partial);
}
// p = r/diag + beta p
__global__ void aat_update_p_kernel(double* __restrict__ p, const double* __restrict__ r, const double* __restrict__ diag, double beta, long n) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) p[i] = r[i] / diag[i] + beta * p[i];
}

int red_blocks_n(long n) { return (int)std::max(1l, std::min((n + 1023) / 1024, (long)GVB_RED_BLOCKS)); }

}   // namespace

extern "C" int gvb_people_stats(gvb_ctx* c, gvb_vec mave_people, gvb_vec msig_people, gvb_vec numb_people) {
    GVB_ARG(c && c->have_stats && mave_people && msig_people && numb_people, "ctx with marker statistics / vectors");
    GVB_ARG(mave_people->cap >= c->Npad && msig_people->cap >= c->Npad && numb_people->cap >= c->Npad, "N-vectors from gvb_vec_alloc_N");
    double* s2 = nullptr;
    GVB_CUDA(cudaMallocAsync(&s2, c->Npad * sizeof(double), c->stream));
    double* outs[3] = {mave_people->d, numb_people->d, s2};
    const int modes[3] = {0, 1, 2};
    int rc = GVB_OK;
    // mode 0 with v = sqrt(N) (a constant vector) undoes the 1/sqrt(N) of the X.v finish: S1 unscaled like the other two
    {
        const long Mpad = c->Mg_pad * 4;
        std::vector<double> ones((size_t)Mpad, sqrt((double)c->N));
        GVB_CUDA(cudaMemcpyAsync(c->tmpM2, ones.data(), Mpad * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        GVB_CUDA(cudaStreamSynchronize(c->stream));
    }
    for (int k = 0; k < 3 && rc == GVB_OK; k++) {
        rc = gvb_ax_tile(c, modes[k] == 0 ? c->tmpM2 : nullptr, outs[k], modes[k]);
        c->sweeps++;
        if (rc == GVB_OK && c->nranks > 1) {   // data.cpp:611-613
            ncclResult_t r = ncclAllReduce(outs[k], outs[k], (size_t)(4 * c->mbytes), ncclDouble, ncclSum, c->comm, c->stream);
            if (r != ncclSuccess) {
                gvb_set_error("NCCL error %s in people statistics", ncclGetErrorString(r));
                rc = GVB_ERR_NCCL;
            }
        }
    }
    if (rc == GVB_OK) {
        people_finalize_kernel<<<(unsigned)((c->Npad + 255) / 256), 256, 0, c->stream>>>(mave_people->d, numb_people->d, s2, c->maskw, c->Npad, mave_people->d,
                                                                                        msig_people->d);
        c->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = GVB_ERR_CUDA;
    }
    cudaFreeAsync(s2, c->stream);
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    return rc;
}

// vamp::CG_solverAAT, denoiserXXT.cpp:57-135: (tau A A^T + gam2) mu = rhs on N-vectors, mu holds the start vector on entry.
// log3[3*i + {0,1,2}] = ||r||/||rhs||, ||mu||, ||z||/||rhs|| per iteration (the reference's progress line).
extern "C" int gvb_cg_solve_aat(gvb_ctx* c, gvb_vec rhs, gvb_vec mu, double tau, double gam2, gvb_vec mave_people, gvb_vec msig_people,
                                gvb_vec numb_people, int max_iter, int* iters, double* log3) {
    GVB_ARG(c && rhs && mu && rhs != mu && mave_people && msig_people && numb_people, "vectors");
    GVB_ARG(rhs->cap >= c->Npad && mu->cap >= c->Npad, "N-vectors from gvb_vec_alloc_N");
    const long n = c->N;
    double* ws = nullptr;   // r, p, d, diag (Npad each) and the M-vector between A^T and A
    GVB_CUDA(cudaMallocAsync(&ws, (4 * (size_t)c->Npad + (size_t)c->Mg_pad * 4) * sizeof(double), c->stream));
    GVB_CUDA(cudaMemsetAsync(ws, 0, (4 * (size_t)c->Npad + (size_t)c->Mg_pad * 4) * sizeof(double), c->stream));
    double *r = ws, *p = ws + c->Npad, *d = ws + 2 * c->Npad, *diag = ws + 3 * c->Npad, *tm = ws + 4 * c->Npad;
    const int nb = red_blocks_n(n);
    const unsigned nbe = (unsigned)std::min((n + 255) / 256, 1184l);
    int rc = GVB_OK, it_done = 0;
    double s[4];
#define AATCHK(x) do { rc = (x); if (rc != GVB_OK) goto done; } while (0)
    aat_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(mave_people->d, msig_people->d, numb_people->d, n, tau, gam2, 1.0 / (double)c->N, diag);
    c->launches++;
    {
        // r = rhs - (tau A A^T + gam2) mu_start, with the reference's all-zero shortcut (denoiserXXT.cpp:18-19)
        gvb_vec xs[1] = {mu};
        double nn = 0.0;
        AATCHK(gvb_vec_dots(c, 1, xs, nullptr, 0, &nn));
        if (nn != 0.0) {
            AATCHK(gvb_atx_dev(c, mu->d, tm));
            AATCHK(gvb_ax_dev(c, tm, d, true));
            aat_combine_kernel<<<nb, 256, 0, c->stream>>>(d, mu->d, tau, gam2, n, c->red_partial);
            c->launches++;
        }
        aat_init_kernel<<<nb, 256, 0, c->stream>>>(r, p, rhs->d, d, diag, n, c->red_partial);
        c->launches++;
        AATCHK(gvb_reduce_finish(c, nb, 2, false, s));
    }
    {
        double rz = s[0];
        const double norm_v = sqrt(s[1]);
        for (int i = 0; i < max_iter; i++) {
            it_done = i + 1;
            AATCHK(gvb_atx_dev(c, p, tm));
            AATCHK(gvb_ax_dev(c, tm, d, true));
            aat_combine_kernel<<<nb, 256, 0, c->stream>>>(d, p, tau, gam2, n, c->red_partial);
            c->launches++;
            double dp = 0.0;
            AATCHK(gvb_reduce_finish(c, nb, 1, false, &dp));
            const double alpha = rz / dp;
            double beta = 1.0 / rz;
            aat_update_kernel<<<nb, 256, 0, c->stream>>>(mu->d, r, p, d, diag, alpha, n, c->red_partial);
            c->launches++;
            AATCHK(gvb_reduce_finish(c, nb, 4, false, s));
            rz = s[0];
            beta *= rz;
            aat_update_p_kernel<<<nbe, 256, 0, c->stream>>>(p, r, diag, beta, n);
            c->launches++;
            const double rel_err = sqrt(s[1]) / norm_v;
            if (log3) { log3[3 * i + 0] = rel_err; log3[3 * i + 1] = sqrt(s[2]); log3[3 * i + 2] = sqrt(s[3]) / norm_v; }
            if (rel_err < 1e-4) break;   // denoiserXXT.cpp:125-129
        }
    }
done:
#undef AATCHK
    cudaFreeAsync(ws, c->stream);
    if (rc != GVB_OK) return rc;
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    if (iters) *iters = it_done;
    return GVB_OK;
}
