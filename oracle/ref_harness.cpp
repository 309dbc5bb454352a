// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, not product code.
//
// A flat C interface over the UNMODIFIED gVAMP reference classes (compiled from /root/reference with
// the single-rank MPI / Boost shims in oracle/shims) so that tests and golden-vector scripts can call
// data::Ax / data::ATx / compute_markers_statistics / vamp::g1 / g1d / updatePrior /
// precondCG_solver directly.  Built into oracle/_ref/libgvamp_ref.so by oracle/Makefile.
// The "private -> public" define only opens the reference's class members for inspection; no
// reference source is modified or copied.
#include <mpi.h>
#include <cstdio>
#include <iostream>
#include <fcntl.h>
#include <unistd.h>
#include <cstring>
#include <string>
#include <vector>

#define private public
#include "data.hpp"
#include "vamp.hpp"
#undef private
#include "options.hpp"
#include "utilities.hpp"

namespace {
struct StdoutSilencer {
    // the reference prints progress lines from rank 0; keep test logs readable
    int saved = -1;
    explicit StdoutSilencer(bool on) {
        if (!on) return;
        fflush(stdout);
        saved = dup(1);
        int dn = open("/dev/null", O_WRONLY);
        dup2(dn, 1);
        close(dn);
    }
    ~StdoutSilencer() {
        if (saved < 0) return;
        fflush(stdout);
        std::cout.flush();
        dup2(saved, 1);
        close(saved);
    }
};
bool g_quiet = true;
}  // namespace

extern "C" {

void ref_set_quiet(int q) { g_quiet = q != 0; }

// ---------------------------------------------------------------- data (data.hpp:93-94)
void* ref_data_create(const char* phen_path, const char* bed_path, int N, int M, int Mt, int S, double alpha_scale) {
    StdoutSilencer s(g_quiet);
    return new data(std::string(phen_path), std::string(bed_path), N, M, Mt, S, 0, "bed", alpha_scale, "");
}
void* ref_data_create_y(const double* y, const char* bed_path, int N, int M, int Mt, int S, double alpha_scale) {
    StdoutSilencer s(g_quiet);
    std::vector<double> yy(y, y + N);
    return new data(yy, std::string(bed_path), N, M, Mt, S, 0, "bed", alpha_scale, "");
}
void* ref_data_create_bim(const char* phen_path, const char* bed_path, int N, int M, int Mt, int S, double alpha_scale, const char* bimfp) {
    StdoutSilencer s(g_quiet);
    return new data(std::string(phen_path), std::string(bed_path), N, M, Mt, S, 0, "bed", alpha_scale, std::string(bimfp));
}
void ref_data_destroy(void* h) { delete static_cast<data*>(h); }
// data::compute_people_statistics data.cpp:548-640; each output has 4*mbytes doubles
void ref_data_people_stats(void* h, double* mave, double* msig, double* numb) {
    StdoutSilencer s(g_quiet);
    data* d = static_cast<data*>(h);
    d->compute_people_statistics();
    std::vector<double> a = d->get_mave_people(), b = d->get_msig_people(), c = d->get_numb_people();
    memcpy(mave, a.data(), a.size() * sizeof(double));
    memcpy(msig, b.data(), b.size() * sizeof(double));
    memcpy(numb, c.data(), c.size() * sizeof(double));
}
// data::pvals_calc data.cpp:1108-1180 / data::pvals_calc_LOCO data.cpp:1220-1353, one estimator; z1 and y have N entries
void ref_data_pvals(void* h, int loco, const double* z1, const double* y, const double* x1_hat, const char* out_path, double* out) {
    StdoutSilencer s(g_quiet);
    data* d = static_cast<data*>(h);
    std::vector<std::vector<double>> z{std::vector<double>(z1, z1 + d->N)}, x{std::vector<double>(x1_hat, x1_hat + d->M)};
    std::vector<double> yy(y, y + d->N);
    std::vector<std::string> fp{std::string(out_path)};
    std::vector<std::vector<double>> r = loco ? d->pvals_calc_LOCO(z, yy, x, fp) : d->pvals_calc(z, yy, x, fp);
    memcpy(out, r[0].data(), sizeof(double) * d->M);
}
long ref_data_mbytes(void* h) { return (long)static_cast<data*>(h)->get_mbytes(); }
int ref_data_nonas(void* h) { return static_cast<data*>(h)->get_nonas(); }
double ref_data_intercept(void* h) { return static_cast<data*>(h)->get_intercept(); }
double ref_data_scale(void* h) { return static_cast<data*>(h)->get_scale(); }
void ref_data_mask4(void* h, unsigned char* out) {
    std::vector<unsigned char>& m = static_cast<data*>(h)->get_mask4();
    memcpy(out, m.data(), m.size());
}
void ref_data_phen(void* h, double* out) {
    std::vector<double> p = static_cast<data*>(h)->get_phen();
    memcpy(out, p.data(), p.size() * sizeof(double));
}
void ref_data_filter_pheno(void* h, double* out) {
    std::vector<double> p = static_cast<data*>(h)->filter_pheno();
    memcpy(out, p.data(), p.size() * sizeof(double));
}
void ref_data_stats(void* h, double* mave, double* msig) {
    data* d = static_cast<data*>(h);
    memcpy(mave, d->get_mave(), sizeof(double) * d->M);
    memcpy(msig, d->get_msig(), sizeof(double) * d->M);
}
// data::Ax(double*, SB, LB) data.cpp:852 ; out has 4*LB doubles
void ref_data_Ax(void* h, const double* v, int SB, int LB, double* out) {
    data* d = static_cast<data*>(h);
    std::vector<double> vv(v, v + d->M);
    std::vector<double> r = d->Ax(vv.data(), SB, LB);
    memcpy(out, r.data(), r.size() * sizeof(double));
}
// data::ATx(double*, SB, LB) data.cpp:814 ; u has 4*LB doubles, out has M doubles
void ref_data_ATx(void* h, const double* u, int SB, int LB, double* out) {
    data* d = static_cast<data*>(h);
    std::vector<double> uu(u, u + 4 * (size_t)LB);
    std::vector<double> r = d->ATx(uu.data(), SB, LB);
    memcpy(out, r.data(), r.size() * sizeof(double));
}
void ref_data_read_covariates(void* h, const char* covfp, int C) {
    StdoutSilencer s(g_quiet);
    static_cast<data*>(h)->read_covariates(std::string(covfp), C);
}

// ---------------------------------------------------------------- vamp (vamp.hpp:78)
// Built through the "all parameters manual" constructor with a default Options, then the knobs the
// tests need are set on the (opened) members.
void* ref_vamp_create(int N, int M, int Mt, double gam1, double gamw, int max_iter, double rho,
                      const double* vars, const double* probs, int L, const char* out_dir, const char* out_name,
                      const char* model, int EM_max_iter, double EM_err_thr, int CG_max_iter, int learn_vars,
                      unsigned long seed, double stop_thr) {
    StdoutSilencer s(g_quiet);
    std::vector<double> vv(vars, vars + L), pp(probs, probs + L);
    vamp* v = new vamp(N, M, Mt, gam1, gamw, max_iter, rho, vv, pp, std::vector<double>(M, 0.0), 0,
                       std::string(out_dir), std::string(out_name), std::string(model));
    v->EM_max_iter = EM_max_iter;
    v->EM_err_thr = EM_err_thr;
    v->CG_max_iter = CG_max_iter;
    v->learn_vars = learn_vars;
    v->seed = seed;
    v->stop_criteria_thr = stop_thr;
    v->store_pvals = 0;
    v->reverse = 0;
    v->use_lmmse_damp = 0;
    v->use_freeze = 0;
    v->init_est = 0;
    v->redglob = 0;
    v->gam1_init = -1;
    v->C = 0;
    v->probit_var = 1;
    return v;
}
void ref_vamp_destroy(void* h) { delete static_cast<vamp*>(h); }
void ref_vamp_set_prior(void* h, const double* vars, const double* probs, int L) {
    vamp* v = static_cast<vamp*>(h);
    v->vars.assign(vars, vars + L);
    v->probs.assign(probs, probs + L);
}
int ref_vamp_get_prior(void* h, double* vars, double* probs) {
    vamp* v = static_cast<vamp*>(h);
    for (size_t i = 0; i < v->vars.size(); i++) { vars[i] = v->vars[i]; probs[i] = v->probs[i]; }
    return (int)v->vars.size();
}
void ref_vamp_set_state(void* h, double gam1, double gam2, double gamw, double probit_var) {
    vamp* v = static_cast<vamp*>(h);
    v->gam1 = gam1; v->gam2 = gam2; v->gamw = gamw; v->probit_var = probit_var;
}
// vamp::g1 / g1d, vamp.cpp:805-869, applied elementwise
void ref_vamp_g1(void* h, const double* r, int n, double gam1, double* out, double* outd) {
    vamp* v = static_cast<vamp*>(h);
    for (int i = 0; i < n; i++) { out[i] = v->g1(r[i], gam1); outd[i] = v->g1d(r[i], gam1); }
}
// vamp::updatePrior, vamp.cpp:929-1072 (uses members r1, gam1, probs, vars)
void ref_vamp_update_prior(void* h, const double* r1, int M, double gam1) {
    StdoutSilencer s(g_quiet);
    vamp* v = static_cast<vamp*>(h);
    v->r1.assign(r1, r1 + M);
    v->gam1 = gam1;
    v->updatePrior(0);
}
// vamp::lmmse_mult vamp.cpp:1074
void ref_vamp_lmmse_mult(void* h, void* dh, const double* vin, int M, double tau, double* out) {
    vamp* v = static_cast<vamp*>(h);
    std::vector<double> r = v->lmmse_mult(std::vector<double>(vin, vin + M), tau, static_cast<data*>(dh), 0);
    memcpy(out, r.data(), sizeof(double) * M);
}
// vamp::precondCG_solver vamp.cpp:1130 (gam2 must have been set through ref_vamp_set_state)
void ref_vamp_cg(void* h, void* dh, const double* rhs, const double* mu0, int M, double tau, int denoiser, double* out) {
    StdoutSilencer s(g_quiet);
    vamp* v = static_cast<vamp*>(h);
    std::vector<double> r = v->precondCG_solver(std::vector<double>(rhs, rhs + M), std::vector<double>(mu0, mu0 + M), tau,
                                                denoiser, static_cast<data*>(dh), 0);
    memcpy(out, r.data(), sizeof(double) * M);
}
// vamp::g2d_onsager vamp.cpp:871 ; also returns the probe and Q^-1 probe
double ref_vamp_onsager(void* h, void* dh, double gam2, double tau, int M, double* bern, double* invq) {
    StdoutSilencer s(g_quiet);
    vamp* v = static_cast<vamp*>(h);
    v->gam2 = gam2;
    double a = v->g2d_onsager(gam2, tau, static_cast<data*>(dh));
    memcpy(bern, v->bern_vec.data(), sizeof(double) * M);
    memcpy(invq, v->invQ_bern_vec.data(), sizeof(double) * M);
    return a;
}
// full inference; returns x1_hat / sqrt(N) like vamp::infere (vamp.cpp:149)
void ref_vamp_infere(void* h, void* dh, int M, double* out) {
    StdoutSilencer s(g_quiet);
    vamp* v = static_cast<vamp*>(h);
    std::vector<double> r = v->infere(static_cast<data*>(dh));
    memcpy(out, r.data(), sizeof(double) * M);
}
// vamp::CG_solverAAT denoiserXXT.cpp:57 (people statistics must have been computed on the data object; gam2 via set_state)
void ref_vamp_cg_aat(void* h, void* dh, const double* rhs, const double* mu0, int n4, double tau, double* out) {
    StdoutSilencer s(g_quiet);
    vamp* v = static_cast<vamp*>(h);
    std::vector<double> r = v->CG_solverAAT(std::vector<double>(rhs, rhs + v->N), std::vector<double>(mu0, mu0 + n4), tau, 0, static_cast<data*>(dh));
    memcpy(out, r.data(), sizeof(double) * n4);
}
double ref_vamp_gamw(void* h) { return static_cast<vamp*>(h)->gamw; }
double ref_vamp_gam1(void* h) { return static_cast<vamp*>(h)->gam1; }

// probit denoiser pieces vamp_probit.cpp:661-726, utilities.cpp:345
void ref_vamp_g1_bin_class(void* h, const double* p, const double* y, const double* mcov, int n, double tau1, double* out, double* outd) {
    vamp* v = static_cast<vamp*>(h);
    for (int i = 0; i < n; i++) {
        out[i] = v->g1_bin_class(p[i], tau1, y[i], mcov[i]);
        outd[i] = v->g1d_bin_class(p[i], tau1, y[i], mcov[i]);
    }
}
double ref_erfcx(double x) { return erfcx(x); }

// utilities.cpp:77 simulate (per-element reseeded mt19937), :259 divide_work is MPI-size bound (1 rank)
void ref_simulate(int M, const double* eta, const double* pi, int K, unsigned long seed, double* out) {
    std::vector<double> r = simulate(M, std::vector<double>(eta, eta + K), std::vector<double>(pi, pi + K), seed);
    memcpy(out, r.data(), sizeof(double) * M);
}

}  // extern "C"
