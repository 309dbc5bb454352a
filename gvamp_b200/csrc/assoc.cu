// assoc.cu -- marker association tests (leave-one-out / leave-one-chromosome-out p-values) on the device.
//
// Replaces data::pvals_calc and the per-chromosome test loop of data::pvals_calc_LOCO (reference data.cpp:1108-1180 and
// :1311-1343) together with linear_reg1d_pvals (utilities.cpp:321-334).  The reference walks every marker column once
// per test and accumulates six sums over the individuals; with value_ij = (a_ij - mu_j) sigma_j b_ij m_i (standardised
// genotype, zero where the genotype is missing or the phenotype is NA) and y_mark = yres + value * c_j,
// c_j = coef_j / sqrt(N) (the marker's own effect added back, LOO; c_j = 0 for LOCO), all six follow from ONE X^T.u
// sweep over the bed plus closed forms of the genotype counts the statistics pass already holds:
//
//   count  = n00 + n10 + n11                                   (masked counts, bit-exact)
//   sumx   = sigma ((2-mu) n00 + (1-mu) n10 - mu n11)
//   sumsqx = sigma^2 ((2-mu)^2 n00 + (1-mu)^2 n10 + mu^2 n11)
//   sxy0   = sum_i value_ij yres_i = sqrt(N) * (X^T (m.yres))_j          <- the sweep
//   B1     = sum_i b_ij m_i yres_i,  B2 = sum_i b_ij m_i yres_i^2        <- by-products of the sweep (sum over the
//            non-missing individuals); B2 needs a second sweep only on shards that have missing genotypes
//   sumxy  = sxy0 + c sumsqx,  sumy = B1 + c sumx,  sumsqy = B2 + 2 c sxy0 + c^2 sumsqx
//
// and the two-sided Student-t tail is the regularised incomplete beta function I_{nu/(nu+t^2)}(nu/2, 1/2), nu = count-2.
#include "gvb_internal.cuh"

namespace {

__global__ void assoc_prep_kernel(const double* __restrict__ yres, const uint32_t* __restrict__ maskw, long Npad, double* __restrict__ w1,
                                  double* __restrict__ w2) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= Npad) return;
    const bool present = (maskw[i >> 2] >> (2 * (i & 3))) & 1u;
    const double y = present ? yres[i] : 0.0;
    w1[i] = y;
    w2[i] = y * y;
}

// continued fraction of the incomplete beta function (modified Lentz)
__device__ double betacf(double a, double b, double x) {
    const double eps = 1e-16, fpmin = 1e-300;
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < fpmin) d = fpmin;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 5000; m++) {   // O(sqrt(max(a, b))) terms: a = count/2 reaches 2e5 at biobank scale
        const double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < fpmin) d = fpmin;
        c = 1.0 + aa / c;
        if (fabs(c) < fpmin) c = fpmin;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < fpmin) d = fpmin;
        c = 1.0 + aa / c;
        if (fabs(c) < fpmin) c = fpmin;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < eps) break;
    }
    return h;
}

__device__ double ibeta(double a, double b, double x) {
    if (!(x > 0.0)) return x == x ? 0.0 : x;
    if (x >= 1.0) return 1.0;
    const double bt = exp(lgamma(a + b) - lgamma(a) - lgamma(b) + a * log(x) + b * log1p(-x));
    if (x < (a + 1.0) / (a + b + 2.0)) return bt * betacf(a, b, x) / a;
    return 1.0 - bt * betacf(b, a, 1.0 - x) / b;
}

// linear_reg1d_pvals, utilities.cpp:321-334
__device__ double reg1d_pvalue(double sumx, double sumsqx, double sumxy, double sumy, double sumsqy, long n) {
    const double dn = (double)n;
    const double s2y = (sumsqy - sumy * sumy / dn) / (dn - 1.0);
    const double s2x = (sumsqx - sumx * sumx / dn) / (dn - 1.0);
    const double sxy = (sumxy - sumx * sumy / dn) / (dn - 1.0);
    const double rxy = sxy / sqrt(s2x * s2y);
    const double t = rxy * sqrt((dn - 2.0) / (1.0 - rxy * rxy));
    const double nu = dn - 2.0;
    return ibeta(0.5 * nu, 0.5, nu / (nu + t * t));   // 2 * upper tail at |t|
}

__global__ void __launch_bounds__(128) assoc_final_kernel(const int64_t* __restrict__ counts, const double* __restrict__ mave, const double* __restrict__ msig,
                                                          const double* __restrict__ t1, const double* __restrict__ B1, const double* __restrict__ B2,
                                                          const double* __restrict__ coef, const double* __restrict__ select, long M, double sqrt_n,
                                                          double* __restrict__ out) {
    long j = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (j >= M) return;
    if (select && select[j] == 0.0) return;
    const double n00 = (double)counts[j * 8 + 0], n10 = (double)counts[j * 8 + 2], n11 = (double)counts[j * 8 + 3];
    const long cnt = counts[j * 8 + 0] + counts[j * 8 + 2] + counts[j * 8 + 3];
    const double mu = mave[j], sg = msig[j];
    const double d2 = 2.0 - mu, d1 = 1.0 - mu;
    const double sumx = sg * (d2 * n00 + d1 * n10 - mu * n11);
    const double sumsqx = sg * sg * (d2 * d2 * n00 + d1 * d1 * n10 + mu * mu * n11);
    const double sxy0 = t1[j] * sqrt_n;
    const double c = coef ? coef[j] / sqrt_n : 0.0;
    const double sumxy = sxy0 + c * sumsqx;
    const double sumy = B1[j] + c * sumx;
    const double sumsqy = B2[j] + 2.0 * c * sxy0 + c * c * sumsqx;
    out[j] = reg1d_pvalue(sumx, sumsqx, sumxy, sumy, sumsqy, cnt);
}

}   // namespace

extern "C" int gvb_assoc_pvals(gvb_ctx* c, gvb_vec yres, gvb_vec coef, gvb_vec select, gvb_vec pvals) {
    GVB_ARG(c && c->have_stats && yres && pvals, "ctx with statistics / vectors");
    const long Mpad = c->Mg_pad * 4;
    GVB_ARG(yres->cap >= c->Npad && pvals->cap >= Mpad && (!coef || coef->cap >= Mpad) && (!select || select->cap >= Mpad),
            "an N-vector and M-vectors from gvb_vec_alloc_N/_M");
    GVB_ARG(c->N < (1l << 40), "N");
    // scratch: w1 = m.yres, w2 = m.yres^2 (N-vectors); t1, B1, B2 (M-vectors)
    double *w1 = c->tmpN, *w2 = c->tmpN2, *t1 = c->tmpM, *B1 = c->tmpM2;
    double* B2 = nullptr;
    GVB_CUDA(cudaMallocAsync(&B2, Mpad * sizeof(double), c->stream));
    assoc_prep_kernel<<<(unsigned)((c->Npad + 255) / 256), 256, 0, c->stream>>>(yres->d, c->maskw, c->Npad, w1, w2);
    GVB_LAUNCHED(c);
    int rc = gvb_atx_tile(c, w1, t1, B1);
    c->sweeps++;
    if (rc == GVB_OK) {
        // shards without missing genotypes: B2 is the same number for every marker, but one more sweep keeps a single code
        // path and costs 5 ms at biobank scale; the association pass runs once per fit
        rc = gvb_atx_tile(c, w2, c->wv, B2);
        c->sweeps++;
    }
    if (rc == GVB_OK) {
        assoc_final_kernel<<<(unsigned)((c->M + 127) / 128), 128, 0, c->stream>>>(c->counts, c->mave, c->msig, t1, B1, B2, coef ? coef->d : nullptr,
                                                                                 select ? select->d : nullptr, c->M, sqrt((double)c->N), pvals->d);
        c->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = GVB_ERR_CUDA;
    }
    cudaFreeAsync(B2, c->stream);
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    return rc;
}
