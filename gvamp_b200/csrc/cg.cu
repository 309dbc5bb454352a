// cg.cu -- the LMMSE operator and the Jacobi-preconditioned conjugate-gradient solver of the VAMP loop
// (vamp::lmmse_mult, vamp::precondCG_solver, vamp.cpp:1074-1229) with DEVICE-RESIDENT scalars.
//
// The reference computes alpha, beta and both exit tests on the host between vector loops.  Here a CG iteration is a fixed
// sequence of launches that never returns to the host:
//
//     X.v sweep (4 launches) -> [NCCL allreduce of the N-vector] -> X^T.u sweep (4-5 launches)
//     -> d = tau d + gam2 p, <d,p>         (cg_combine_dot_kernel, per-block partials)
//     -> reduce [-> NCCL allreduce of 1 scalar] -> alpha = <r,z>/<d,p>                     (cg_scal_alpha_kernel, 1 thread)
//     -> [A mu += alpha A p]  mu += alpha p ; r -= alpha d ; [A^T A mu += ...] ; 4 partial sums  (cg_update_mu_r_kernel)
//     -> reduce [-> NCCL allreduce of 4 scalars] -> Onsager test, beta, <r,z>, ||r||/||rhs|| test, log (cg_scal_update_kernel)
//     -> p = r/diag + beta p
//
// Every scalar lives in c->cg_dev; the kernels read alpha / beta from there.  When an exit test fires, cg_scal_update_kernel
// raises c->cg_flags[0] and every later kernel of the solve -- including the sweep kernels, through gvb_ctx::skip -- returns at
// once.  The host therefore enqueues iteration i+1 BEFORE it knows whether iteration i was the last one, and reads the flag of
// iteration i (a 128-byte copy into a pinned ring + an event) only after that: the stream never drains inside a solve.  The
// one speculative iteration after the last real one costs ~20 empty launches.  Arithmetic: FP64, the reference's formulas in
// the reference's order (alpha = rz/dp; beta = 1/rz_old, then *= rz_new; rel_err = sqrt(rr)/sqrt(vv)), IEEE division / sqrt.
// GVB_CG_LAG=0 (and the legacy cross-check kernels, which do not know the predicate) wait for every iteration's flag instead.
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "gvb_internal.cuh"

// scalar slots of c->cg_dev
enum { CG_RZ = 0, CG_VV = 1, CG_ALPHA = 2, CG_BETA = 3, CG_PREV_ONS = 4, CG_RHS_MU = 5, CG_RR = 6, CG_ITERS = 7, CG_DONE = 8, CG_RHS_R = 9, CG_DP = 10 };

namespace {

__global__ void lmmse_combine_kernel(double* __restrict__ out, double tau, double gam2, const double* __restrict__ v, long n) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        double r = out[i] * tau;     // vamp.cpp:1113-1114: res *= tau; res += gam2*v
        out[i] = r + gam2 * v[i];
    }
}

template <int K>
__device__ __forceinline__ void block_sums(double (&a)[K], double* __restrict__ partial) {
    __shared__ double sm[8][K];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
        if (lane == 0) sm[warp][k] = a[k];
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double s = 0.0;
        for (int w = 0; w < 8; w++) s += sm[w][threadIdx.x];
        partial[blockIdx.x * K + threadIdx.x] = s;
    }
}

// r = rhs - q ; p = r/diag ; partials: <r, r/diag>, ||rhs||^2, <rhs, mu>   (q == nullptr: the start vector is zero, q = 0)
__global__ void __launch_bounds__(256) cg_init_kernel(double* __restrict__ r, double* __restrict__ p, const double* __restrict__ rhs,
                                                      const double* __restrict__ q, const double* __restrict__ mu, double diag, long n,
                                                      double* __restrict__ partial) {
    double a[3] = {0.0, 0.0, 0.0};
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const double b = rhs[i];
        const double x = q ? b - q[i] : b;
        r[i] = x;
        const double z = x / diag;
        p[i] = z;
        a[0] += x * z;
        a[1] += b * b;
        if (q) a[2] += b * mu[i];
    }
    block_sums<3>(a, partial);
}

// ata = (d - gam2*mu)/tau: A^T A mu out of d = (tau A^T A + gam2) mu;  d = tau*ata + gam2*mu: the way back
__global__ void cg_ata_from_d_kernel(double* __restrict__ ata, const double* __restrict__ d, const double* __restrict__ mu, double gam2, double tau, long n) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) ata[i] = (d[i] - gam2 * mu[i]) / tau;
}
__global__ void cg_d_from_ata_kernel(double* __restrict__ d, const double* __restrict__ ata, const double* __restrict__ mu, double gam2, double tau, long n) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) d[i] = tau * ata[i] + gam2 * mu[i];
}

// d = tau*d + gam2*p (the tail of lmmse_mult, vamp.cpp:1113-1114) fused with <d,p>
__global__ void __launch_bounds__(256) cg_combine_dot_kernel(double* __restrict__ d, double tau, double gam2, const double* __restrict__ p, long n,
                                                             double* __restrict__ partial, const int* __restrict__ stop) {
    if (*stop) return;
    double a[1] = {0.0};
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        double x = d[i] * tau;
        x = x + gam2 * p[i];
        d[i] = x;
        a[0] += x * p[i];
    }
    block_sums<1>(a, partial);
}

// ax += alpha * ap over the padded N-vector: the running A.mu by-product
__global__ void cg_axpy_n_kernel(double* __restrict__ ax, const double* __restrict__ ap, const double* __restrict__ S, long n, const int* __restrict__ stop) {
    if (*stop) return;
    const double alpha = S[CG_ALPHA];
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) ax[i] += alpha * ap[i];
}

// mu += alpha p ; r -= alpha d ; [ata += alpha (d - gam2 p)/tau] ; partials: <rhs,mu>, ||mu||^2, <r, r/diag>, ||r||^2.
// One kernel for the two updates of a CG iteration (vamp.cpp:1160-1207); the Onsager exit test, which the reference places
// between them, only reads <rhs,mu>, so testing it after both leaves mu and the decision unchanged.
__global__ void __launch_bounds__(256) cg_update_mu_r_kernel(double* __restrict__ mu, double* __restrict__ r, const double* __restrict__ p,
                                                             const double* __restrict__ d, const double* __restrict__ rhs, double* __restrict__ ata,
                                                             const double* __restrict__ S, double diag, double gam2, double tau, long n,
                                                             double* __restrict__ partial, const int* __restrict__ stop) {
    if (*stop) return;
    const double alpha = S[CG_ALPHA];
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const double pi = p[i], di = d[i];
        const double m = mu[i] + alpha * pi;
        const double x = r[i] - di * alpha;
        mu[i] = m;
        r[i] = x;
        if (ata) ata[i] += alpha * ((di - gam2 * pi) / tau);
        a[0] += rhs[i] * m;
        a[1] += m * m;
        a[2] += x * (x / diag);
        a[3] += x * x;
    }
    block_sums<4>(a, partial);
}

// p = r/diag + beta*p
__global__ void cg_update_p_kernel(double* __restrict__ p, const double* __restrict__ r, const double* __restrict__ S, double diag, long n,
                                   const int* __restrict__ stop) {
    if (*stop) return;
    const double beta = S[CG_BETA];
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) p[i] = r[i] / diag + beta * p[i];
}

// zero-start solve with a cached A^T A rhs (see cg_solve_impl, `ata_rhs`): the first search direction is p0 = rhs / diag, so
// A^T A p0 = (A^T A rhs) / diag -- d receives what the two sweeps of iteration 0 would have left there (before the combine kernel)
__global__ void cg_d_from_cached_kernel(double* __restrict__ d, const double* __restrict__ ata_rhs, double diag, long n) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) d[i] = ata_rhs[i] / diag;
}
// ... and the way in: after the sweeps of iteration 0 of a zero-start solve, d = A^T A (rhs / diag)
__global__ void cg_cache_from_d_kernel(double* __restrict__ ata_rhs, const double* __restrict__ d, double diag, long n) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) ata_rhs[i] = d[i] * diag;
}

// ---- the scalar steps, one thread each; R = c->red_result (already summed over blocks and ranks) ----
__global__ void cg_scal_init_kernel(double* __restrict__ S, int* __restrict__ flags, const double* __restrict__ R) {
    S[CG_RZ] = R[0];
    S[CG_VV] = R[1];
    S[CG_RHS_MU] = R[2];
    S[CG_ALPHA] = S[CG_BETA] = S[CG_PREV_ONS] = S[CG_RR] = S[CG_ITERS] = S[CG_DONE] = S[CG_RHS_R] = S[CG_DP] = 0.0;
    flags[0] = 0;
}
__global__ void cg_scal_alpha_kernel(double* __restrict__ S, const int* __restrict__ flags, const double* __restrict__ R) {
    if (flags[0]) return;
    S[CG_DP] = R[0];
    S[CG_ALPHA] = S[CG_RZ] / R[0];   // vamp.cpp:1165
}
// after the two updates of iteration i: the Onsager exit (vamp.cpp:1174-1193), beta (:1198,:1207), the residual exit (:1215-1223)
// and the iteration's log line {||r||/||rhs||, ||mu||, ||z||/||rhs||, Onsager relative change}.  residual_only: the successive-
// difference Onsager exit is only meaningful for a zero start (where <u,mu_k> grows monotonically); a warm-started Onsager
// solve (GVB_ONSAGER_WARM=1) stops on the residual alone.
__global__ void cg_scal_update_kernel(double* __restrict__ S, int* __restrict__ flags, const double* __restrict__ R, double* __restrict__ log, int i,
                                      int denoiser, int residual_only, double gam2, double diag) {
    if (flags[0]) return;
    const double s0 = R[0], s1 = R[1], s2 = R[2], s3 = R[3];
    const double norm_mu = sqrt(s1);
    S[CG_RHS_MU] = s0;
    S[CG_ITERS] = (double)(i + 1);
    double ons_rel = -1.0;
    if (denoiser == 0) {
        const double onsager = gam2 * s0;
        ons_rel = (onsager != 0.0) ? fabs((onsager - S[CG_PREV_ONS]) / onsager) : 1.0;
        if (!residual_only && ons_rel < 1e-8) {
            log[4 * i + 0] = -1.0; log[4 * i + 1] = norm_mu; log[4 * i + 2] = -1.0; log[4 * i + 3] = ons_rel;
            S[CG_DONE] = 1.0;
            flags[0] = 1;
            return;
        }
        S[CG_PREV_ONS] = onsager;
    }
    double beta = 1.0 / S[CG_RZ];   // vamp.cpp:1198
    S[CG_RZ] = s2;
    S[CG_RR] = s3;
    beta *= s2;                     // vamp.cpp:1207
    S[CG_BETA] = beta;
    const double norm_v = sqrt(S[CG_VV]);
    const double rel_err = sqrt(s3) / norm_v;   // vamp.cpp:1215
    const double norm_z = sqrt(s3) / diag;      // ||z|| with z = r/diag
    log[4 * i + 0] = rel_err; log[4 * i + 1] = norm_mu; log[4 * i + 2] = norm_z / norm_v; log[4 * i + 3] = ons_rel;
    if (rel_err < 1e-5) {           // vamp.cpp:1217-1223
        S[CG_DONE] = 1.0;
        flags[0] = 1;
    }
}
__global__ void cg_scal_final_kernel(double* __restrict__ S, const double* __restrict__ R) { S[CG_RHS_R] = R[0]; }

inline int cg_blocks(long n) { return (int)std::max(1l, std::min((n + 1023) / 1024, (long)GVB_RED_BLOCKS)); }

int ensure_cg_state(gvb_ctx* c, int max_iter) {
    if (!c->cg_flags) {
        GVB_CUDA(gvb_malloc(c, &c->cg_flags, 4 * sizeof(int)));
        GVB_CUDA(cudaMemsetAsync(c->cg_flags, 0, 4 * sizeof(int), c->stream));
        for (int k = 0; k < 4; k++) GVB_CUDA(cudaEventCreateWithFlags(&c->cg_ev[k], cudaEventDisableTiming));
    }
    if (!c->cg_dev || c->cg_log_cap < max_iter) {
        const int cap = std::max(64, max_iter);
        if (c->cg_dev) { GVB_CUDA(cudaStreamSynchronize(c->stream)); cudaFree(c->cg_dev); cudaFreeHost(c->cg_host); }
        c->cg_dev = nullptr;
        c->cg_host = nullptr;
        c->cg_log_cap = 0;
        GVB_CUDA(gvb_malloc(c, &c->cg_dev, (GVB_CG_NSCAL + 4 * (size_t)cap) * sizeof(double)));
        GVB_CUDA(cudaMallocHost(&c->cg_host, (5 * GVB_CG_NSCAL + 4 * (size_t)cap) * sizeof(double)));
        GVB_CUDA(cudaMemsetAsync(c->cg_dev, 0, (GVB_CG_NSCAL + 4 * (size_t)cap) * sizeof(double), c->stream));
        c->cg_log_cap = cap;
    }
    return GVB_OK;
}

struct SkipGuard {   // the sweep predicate is armed only while the solver's iterations are being enqueued
    gvb_ctx* c;
    explicit SkipGuard(gvb_ctx* c_) : c(c_) {}
    ~SkipGuard() { c->skip = nullptr; }
};

}   // namespace

static int lmmse_mult_dev(gvb_ctx* c, const gvb_vec_s* v, double tau, double gam2, gvb_vec_s* out, bool known_nonzero) {
    if (!known_nonzero) {
        // vamp.cpp:1079-1080: an all-zero input returns zeros without touching the matrix
        gvb_vec xs[1] = {const_cast<gvb_vec_s*>(v)};
        double nn = 0.0;
        GVB_CHECK(gvb_vec_dots(c, 1, xs, nullptr, 1, &nn));
        if (nn == 0.0) return gvb_vec_fill(c, out, 0.0);
    }
    GVB_CHECK(gvb_ax_dev(c, v->d, c->tmpN2, true));
    GVB_CHECK(gvb_atx_dev(c, c->tmpN2, out->d));
    lmmse_combine_kernel<<<(unsigned)std::min((v->n + 255) / 256, 1184l), 256, 0, c->stream>>>(out->d, tau, gam2, v->d, v->n);
    GVB_LAUNCHED(c);
    return GVB_OK;
}

extern "C" int gvb_lmmse_mult(gvb_ctx* c, gvb_vec v, double tau, double gam2, gvb_vec out) {
    GVB_ARG(c && v && out && v != out && v->cap >= c->Mg_pad * 4 && out->cap >= c->Mg_pad * 4, "M-vectors from gvb_vec_alloc_M");
    return lmmse_mult_dev(c, v, tau, gam2, out, false);
}

// vamp::precondCG_solver, vamp.cpp:1130-1229.  z = r/diag is never materialised (diag is a constant, :1137-1138).
//
// By-products (both optional, no extra bed sweep): every iteration already forms A p (the N-vector inside lmmse_mult), so
// ax_mu = A mu_start + sum_k alpha_k A p_k is A times the returned solution; and dots3 = {<rhs,rhs>, <rhs,mu>, <rhs,r>} with the
// residual r = rhs - Q mu of the returned mu gives <rhs, A^T A mu> = (dots3[0] - gam2 dots3[1] - dots3[2]) / tau.
// ata_mu (optional, needs ax_mu) = A^T A mu, accumulated the same way from d_k = Q p_k.
// have_start: 0 = mu holds any start vector (two sweeps for the initial residual, like the reference); 1 = ax_mu / ata_mu hold A mu /
// A^T A mu of the start vector (the by-products of the solve that produced it, whatever its tau / gam2 were): the initial residual
// needs no sweep; 2 = the caller states that the start vector is zero (it is cleared here): no sweep either -- the reference's
// zero-vector shortcut of lmmse_mult (vamp.cpp:1079-1080) without a host-visible norm.
//
// ata_rhs / ata_rhs_state (optional, zero-start solves without by-product vectors only): a cache of A^T A rhs for a right-hand side that
// the caller solves against again and again -- the Onsager probe, which is the same vector in every VAMP iteration (mt19937{seed + S},
// vamp.cpp:875-882).  With a zero start the first search direction is rhs / diag, so the operator product of iteration 0 is
// (tau A^T A rhs + gam2 rhs) / diag and needs no sweep once A^T A rhs is known.  *state == 0: iteration 0 sweeps and fills the cache
// (*state becomes 1); *state == 1: iteration 0 takes it from the cache.  Like the by-products this is an identity of exact arithmetic;
// in FP64 it differs from the sweeps by their own fixed-point error (linearity of the sweeps holds to ~1e-7).
//
// phase (the prepared solve, gvb_cg_prepare / gvb_cg_solve_prepared): 0 = the whole solve; 1 = only its start - initial residual, first
// search direction p0 - followed by ONE dual sweep that forms A p0 (kept in c->cg_ap) together with extra_out = A extra_v, a product the
// caller needs anyway (z1 = A x1_hat of the VAMP iteration): one bed read instead of two; 2 = the rest of a solve prepared that way:
// iteration 0 takes A p0 from c->cg_ap.  Same kernels in the same order on the same data as phase 0: the results are bit-identical.
static int cg_solve_impl(gvb_ctx* c, gvb_vec rhs, gvb_vec mu, double tau, double gam2, int max_iter, int denoiser, int* iters, double* log4,
                         gvb_vec ax_mu, double* dots3, gvb_vec ata_mu, int have_start, gvb_vec ata_rhs = nullptr, int* ata_rhs_state = nullptr,
                         int phase = 0, gvb_vec extra_v = nullptr, gvb_vec extra_out = nullptr) {
    GVB_ARG(c && rhs && mu && rhs != mu, "vectors");
    GVB_ARG(phase == 0 || (have_start != 0 && !ata_rhs), "a prepared solve starts from zero or from a vector with known by-products");
    GVB_ARG(phase != 1 || (extra_v && extra_out && extra_v->cap >= c->Mg_pad * 4 && extra_out->cap >= c->Npad), "the companion product needs an M- and an N-vector");
    GVB_ARG(phase != 2 || c->cg_prepared, "gvb_cg_solve_prepared without gvb_cg_prepare");
    if (phase != 2) c->cg_prepared = false;
    GVB_ARG(max_iter >= 0, "max_iter");
    GVB_ARG(rhs->cap >= c->Mg_pad * 4 && mu->cap >= c->Mg_pad * 4, "M-vectors from gvb_vec_alloc_M");
    GVB_ARG(!ax_mu || ax_mu->cap >= c->Npad, "ax_mu must be an N-vector from gvb_vec_alloc_N");
    GVB_ARG(!ata_mu || (ax_mu && ata_mu->cap >= c->Mg_pad * 4 && ata_mu != mu && ata_mu != rhs), "ata_mu needs ax_mu and must be its own M-vector");
    GVB_ARG(have_start != 1 || ata_mu, "have_start = 1 needs ax_mu and ata_mu of the start vector");
    GVB_ARG(!ata_rhs || (ata_rhs_state && have_start == 2 && !ax_mu && !ata_mu && ata_rhs->cap >= c->Mg_pad * 4 && ata_rhs != rhs && ata_rhs != mu),
            "the A^T A rhs cache serves zero-start solves without by-product vectors");
    const long n = c->M;
    for (int k = 0; k < 3; k++) {
        if (c->cg_ws[k] && c->cg_ws[k]->cap != c->Mg_pad * 4) {   // matrix was reloaded with another shape
            gvb_vec_free(c, c->cg_ws[k]);
            c->cg_ws[k] = nullptr;
        }
        if (!c->cg_ws[k]) GVB_CHECK(gvb_vec_alloc_M(c, &c->cg_ws[k]));
    }
    GVB_CHECK(ensure_cg_state(c, max_iter));
    gvb_vec r = c->cg_ws[0], p = c->cg_ws[1], d = c->cg_ws[2];
    double* S = c->cg_dev;
    double* L = c->cg_dev + GVB_CG_NSCAL;
    int* flags = c->cg_flags;
    const double diag = tau * (double)(c->N - 1) / (double)c->N + gam2;
    const int nb = cg_blocks(n);
    const unsigned nbm = (unsigned)std::min((n + 255) / 256, 1184l);
    const int nbn = (int)std::min((c->Npad + 255) / 256, 1184l);
    const char* lag_env = getenv("GVB_CG_LAG");
    const int lag = (c->kernel_gen == 2 && !(lag_env && lag_env[0] == '0')) ? 1 : 0;
    // an Onsager solve that does not start from zero stops on the residual only (see cg_scal_update_kernel)
    const int residual_only = (denoiser == 0 && have_start == 1) ? 1 : 0;

    if (phase == 1 && c->cg_ap_cap < (size_t)c->Npad) {
        if (c->cg_ap) cudaFree(c->cg_ap);
        c->cg_ap = nullptr;
        c->cg_ap_cap = 0;
        GVB_CUDA(gvb_malloc(c, &c->cg_ap, (size_t)c->Npad * sizeof(double)));
        c->cg_ap_cap = (size_t)c->Npad;
    }
    // ---- initial residual r = rhs - Q mu_start ; p = r/diag
    const double* q = d->d;
    if (phase == 2) {
        // done by gvb_cg_prepare
    } else if (have_start == 1) {
        cg_d_from_ata_kernel<<<nbm, 256, 0, c->stream>>>(d->d, ata_mu->d, mu->d, gam2, tau, n);
        GVB_LAUNCHED(c);
    } else if (have_start == 2) {
        GVB_CUDA(cudaMemsetAsync(mu->d, 0, mu->cap * sizeof(double), c->stream));
        if (ax_mu) GVB_CUDA(cudaMemsetAsync(ax_mu->d, 0, c->Npad * sizeof(double), c->stream));
        if (ata_mu) GVB_CUDA(cudaMemsetAsync(ata_mu->d, 0, ata_mu->cap * sizeof(double), c->stream));
        q = nullptr;
    } else {
        GVB_CHECK(lmmse_mult_dev(c, mu, tau, gam2, d, true));
        if (ax_mu) GVB_CUDA(cudaMemcpyAsync(ax_mu->d, c->tmpN2, c->Npad * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        if (ata_mu) {
            cg_ata_from_d_kernel<<<nbm, 256, 0, c->stream>>>(ata_mu->d, d->d, mu->d, gam2, tau, n);
            GVB_LAUNCHED(c);
        }
    }
    if (phase != 2) {
        cg_init_kernel<<<nb, 256, 0, c->stream>>>(r->d, p->d, rhs->d, q, mu->d, diag, n, c->red_partial);
        GVB_LAUNCHED(c);
        GVB_CHECK(gvb_reduce_device(c, nb, 3, true));
        cg_scal_init_kernel<<<1, 1, 0, c->stream>>>(S, flags, c->red_result);
        GVB_LAUNCHED(c);
    }
    if (phase == 1) {
        GVB_CHECK(gvb_ax2_dev(c, p->d, extra_v->d, c->cg_ap, extra_out->d));
        c->cg_prepared = true;
        return GVB_OK;
    }
    c->cg_prepared = false;

    // ---- iterations: enqueue i, then look at the flag of iteration i - lag
    int enqueued = 0;
    bool stopped = false;
    gvb_cg_companion_fn companion = (phase != 1) ? c->cg_companion : nullptr;   // serves this solve only
    void* companion_user = c->cg_companion_user;
    if (phase != 1) c->cg_companion = nullptr;
    bool last_had_companion = false, last_companion_atx = false;
    {
        SkipGuard guard(c);
        c->skip = flags;
        for (int i = 0; i < max_iter && !stopped; i++) {
            double* ap = c->tmpN2;
            if (i == 0 && ata_rhs && *ata_rhs_state == 1) {       // p0 = rhs / diag: A^T A p0 from the cache, no sweep
                cg_d_from_cached_kernel<<<nbm, 256, 0, c->stream>>>(d->d, ata_rhs->d, diag, n);
                GVB_LAUNCHED(c);
            } else {
                gvb_vec cv = nullptr, cav = nullptr, cw = nullptr;
                bool with_companion = false;
                if (phase == 2 && i == 0) {
                    ap = c->cg_ap;                                     // A p0 came with the dual sweep of gvb_cg_prepare
                } else {
                    // a companion (gvb_cg_set_companion) may ride on this iteration's A p: one dual sweep {A p, av = A v}.  That sweep is
                    // not predicated on the solver's exit flag - the companion's product must exist even when this iteration turns out
                    // to be a speculative one (A p is then simply not used)
                    if (companion && companion(companion_user, 0, i, &cv, &cav, &cw) == 1) {
                        GVB_ARG(cv && cav && cv->cap >= c->Mg_pad * 4 && cav->cap >= c->Npad, "the companion product needs an M- and an N-vector");
                        with_companion = true;
                        c->skip = nullptr;
                        int rc2 = gvb_ax2_dev(c, p->d, cv->d, ap, cav->d);
                        c->skip = flags;
                        GVB_CHECK(rc2);
                    } else {
                        GVB_CHECK(gvb_ax_dev(c, p->d, ap, true));      // d = Q p
                    }
                }
                last_had_companion = with_companion;
                last_companion_atx = false;
                if (with_companion && cw) {   // ... and its w = A^T av on this iteration's A^T (A p), the same way
                    GVB_ARG(cw->cap >= c->Mg_pad * 4, "the companion's A^T product needs an M-vector");
                    c->skip = nullptr;
                    int rc2 = gvb_atx2_dev(c, ap, cav->d, d->d, cw->d);
                    c->skip = flags;
                    GVB_CHECK(rc2);
                    last_companion_atx = true;
                } else {
                    GVB_CHECK(gvb_atx_dev(c, ap, d->d));
                }
                if (with_companion) {   // the companion finishes its own step (its sweeps and host-visible sums are not the solver's)
                    c->skip = nullptr;
                    int rc2 = companion(companion_user, 1, i, &cv, &cav, &cw);
                    c->skip = flags;
                    if (rc2 < 0) {
                        gvb_set_error("the companion of the solve failed in iteration %d", i);
                        return GVB_ERR_ARG;
                    }
                }
                if (i == 0 && ata_rhs) {
                    cg_cache_from_d_kernel<<<nbm, 256, 0, c->stream>>>(ata_rhs->d, d->d, diag, n);
                    GVB_LAUNCHED(c);
                    *ata_rhs_state = 1;
                }
            }
            cg_combine_dot_kernel<<<nb, 256, 0, c->stream>>>(d->d, tau, gam2, p->d, n, c->red_partial, flags);
            GVB_LAUNCHED(c);
            GVB_CHECK(gvb_reduce_device(c, nb, 1, true));
            cg_scal_alpha_kernel<<<1, 1, 0, c->stream>>>(S, flags, c->red_result);
            GVB_LAUNCHED(c);
            if (ax_mu) {   // ap still holds A p of this iteration
                cg_axpy_n_kernel<<<nbn, 256, 0, c->stream>>>(ax_mu->d, ap, S, c->Npad, flags);
                GVB_LAUNCHED(c);
            }
            cg_update_mu_r_kernel<<<nb, 256, 0, c->stream>>>(mu->d, r->d, p->d, d->d, rhs->d, ata_mu ? ata_mu->d : nullptr, S, diag, gam2, tau, n,
                                                             c->red_partial, flags);
            GVB_LAUNCHED(c);
            GVB_CHECK(gvb_reduce_device(c, nb, 4, true));
            cg_scal_update_kernel<<<1, 1, 0, c->stream>>>(S, flags, c->red_result, L, i, denoiser, residual_only, gam2, diag);
            GVB_LAUNCHED(c);
            cg_update_p_kernel<<<nbm, 256, 0, c->stream>>>(p->d, r->d, S, diag, n, flags);
            GVB_LAUNCHED(c);
            enqueued = i + 1;
            double* slot = c->cg_host + (i & 3) * GVB_CG_NSCAL;
            GVB_CUDA(cudaMemcpyAsync(slot, S, GVB_CG_NSCAL * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            GVB_CUDA(cudaEventRecord(c->cg_ev[i & 3], c->stream));
            const int look = i - lag;
            if (look >= 0) {
                GVB_CUDA(cudaEventSynchronize(c->cg_ev[look & 3]));
                c->host_syncs++;
                stopped = c->cg_host[(look & 3) * GVB_CG_NSCAL + CG_DONE] != 0.0;
            }
        }
    }
    // ---- the solver's final state: scalars, log and (optionally) <rhs, r>
    if (dots3) {   // r is the residual of the returned mu in every exit path (the two updates are one kernel)
        gvb_vec xs[1] = {rhs}, ys[1] = {r};
        GVB_CHECK(gvb_vec_dots_device(c, 1, xs, ys, true));
        cg_scal_final_kernel<<<1, 1, 0, c->stream>>>(S, c->red_result);
        GVB_LAUNCHED(c);
    }
    double* fin = c->cg_host + 4 * GVB_CG_NSCAL;
    const size_t fin_n = GVB_CG_NSCAL + 4 * (size_t)std::max(enqueued, 0);
    GVB_CUDA(cudaMemcpyAsync(fin, c->cg_dev, fin_n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    c->host_syncs++;
    const int it_done = (int)fin[CG_ITERS];
    // iterations enqueued after the solver had stopped ran as empty launches: they are not sweeps and carry no timing
    const int spec = enqueued - it_done;
    if (spec > 0) {
        c->sweeps -= 2l * spec - ((last_had_companion && spec >= 1) ? (last_companion_atx ? 2 : 1) : 0);   // a companion's dual sweeps ran in full
        for (int w = 0; w < 2; w++)
            if (c->profile && c->prof_used[w] >= (size_t)(2 * spec)) c->prof_used[w] -= (size_t)(2 * spec);
    }
    if (log4)
        for (int i = 0; i < 4 * it_done; i++) log4[i] = fin[GVB_CG_NSCAL + i];
    if (dots3) {
        dots3[0] = fin[CG_VV];
        dots3[1] = fin[CG_RHS_MU];
        dots3[2] = fin[CG_RHS_R];
    }
    if (iters) *iters = it_done;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        gvb_set_error("CUDA error in CG: %s", cudaGetErrorString(e));
        return GVB_ERR_CUDA;
    }
    return GVB_OK;
}

// the public entry points keep the reference's zero-vector shortcut (vamp.cpp:1079-1080) for callers that do not know
// what their start vector holds: one host-visible norm per solve decides between "sweep" and "start from zero"
static int start_mode(gvb_ctx* c, gvb_vec mu, int have_start, int* mode) {
    *mode = have_start;
    if (have_start != 0) return GVB_OK;
    GVB_ARG(c && mu, "vectors");
    gvb_vec xs[1] = {mu};
    double nn = 0.0;
    GVB_CHECK(gvb_vec_dots(c, 1, xs, nullptr, 1, &nn));
    if (nn == 0.0) *mode = 2;
    return GVB_OK;
}

extern "C" int gvb_cg_solve(gvb_ctx* c, gvb_vec rhs, gvb_vec mu, double tau, double gam2, int max_iter, int denoiser, int* iters, double* log4) {
    int mode = 0;
    GVB_CHECK(start_mode(c, mu, 0, &mode));
    return cg_solve_impl(c, rhs, mu, tau, gam2, max_iter, denoiser, iters, log4, nullptr, nullptr, nullptr, mode);
}
extern "C" int gvb_cg_solve_ex(gvb_ctx* c, gvb_vec rhs, gvb_vec mu, double tau, double gam2, int max_iter, int denoiser, int* iters, double* log4,
                               gvb_vec ax_mu, double* dots3) {
    int mode = 0;
    GVB_CHECK(start_mode(c, mu, 0, &mode));
    return cg_solve_impl(c, rhs, mu, tau, gam2, max_iter, denoiser, iters, log4, ax_mu, dots3, nullptr, mode);
}
extern "C" int gvb_cg_solve_cached(gvb_ctx* c, gvb_vec rhs, gvb_vec mu, double tau, double gam2, int max_iter, int denoiser, int* iters, double* log4,
                                   gvb_vec ata_rhs, int* ata_rhs_state, double* dots3) {
    GVB_ARG(ata_rhs && ata_rhs_state && (*ata_rhs_state == 0 || *ata_rhs_state == 1), "cache vector and its state (0 or 1)");
    return cg_solve_impl(c, rhs, mu, tau, gam2, max_iter, denoiser, iters, log4, nullptr, dots3, nullptr, 2, ata_rhs, ata_rhs_state);
}
extern "C" int gvb_cg_solve_warm(gvb_ctx* c, gvb_vec rhs, gvb_vec mu, double tau, double gam2, int max_iter, int denoiser, int* iters, double* log4,
                                 gvb_vec ax_mu, gvb_vec ata_mu, int have_start, double* dots3) {
    GVB_ARG(have_start >= 0 && have_start <= 2, "have_start is 0, 1 or 2");
    int mode = 0;
    GVB_CHECK(start_mode(c, mu, have_start, &mode));
    return cg_solve_impl(c, rhs, mu, tau, gam2, max_iter, denoiser, iters, log4, ax_mu, dots3, ata_mu, mode);
}

// The LMMSE solve of a VAMP iteration in two calls, so that its first product A p0 shares a bed read with a product the caller needs
// anyway: gvb_cg_prepare forms the initial residual and p0 (have_start 1 or 2 as in gvb_cg_solve_warm) and runs one dual sweep
// {A p0, extra_out = A extra_v}; gvb_cg_solve_prepared runs the iterations.  Anything may be enqueued between the two calls except
// another solve.  Results are bit-identical to gvb_dAx(extra_v) followed by gvb_cg_solve_warm.
extern "C" int gvb_cg_prepare(gvb_ctx* c, gvb_vec rhs, gvb_vec mu, double tau, double gam2, int max_iter, gvb_vec ax_mu, gvb_vec ata_mu, int have_start,
                              gvb_vec extra_v, gvb_vec extra_out) {
    GVB_ARG(have_start == 1 || have_start == 2, "have_start is 1 or 2");
    return cg_solve_impl(c, rhs, mu, tau, gam2, max_iter, 1, nullptr, nullptr, ax_mu, nullptr, ata_mu, have_start, nullptr, nullptr, 1, extra_v, extra_out);
}
extern "C" int gvb_cg_solve_prepared(gvb_ctx* c, gvb_vec rhs, gvb_vec mu, double tau, double gam2, int max_iter, int denoiser, int* iters, double* log4,
                                     gvb_vec ax_mu, gvb_vec ata_mu, int have_start, double* dots3) {
    GVB_ARG(have_start == 1 || have_start == 2, "have_start is 1 or 2");
    return cg_solve_impl(c, rhs, mu, tau, gam2, max_iter, denoiser, iters, log4, ax_mu, dots3, ata_mu, have_start, nullptr, nullptr, 2);
}

// A companion for the NEXT solve of this context (consumed by it; NULL clears): before the product A p of an iteration the solver calls
// fn(user, 0, i, &v, &av, &w); when that returns 1 with an M-vector v and an N-vector av, the iteration's X.v becomes one dual sweep
// {A p, av = A v}; when it also sets an M-vector w, the iteration's X^T.u becomes a dual sweep too, {A^T (A p), w = A^T av}.  After
// that the solver calls fn(user, 1, i, ...) so that the companion can finish its own step
// (it may enqueue sweeps and synchronise; a negative return aborts the solve).  vamp::infere_linear lets the Lanczos steps of the
// Onsager projection ride on the LMMSE solve of iteration 1 this way.
extern "C" int gvb_cg_set_companion(gvb_ctx* c, gvb_cg_companion_fn fn, void* user) {
    GVB_ARG(c, "ctx");
    c->cg_companion = fn;
    c->cg_companion_user = user;
    return GVB_OK;
}
