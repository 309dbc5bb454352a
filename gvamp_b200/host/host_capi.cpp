// host_capi.cpp -- flat C entry points over the C++ host classes (data, vamp) for bench.py and the
// Python tests (ctypes).  The product drivers are the executables built from main_real*.cpp; this shim
// only lets a Python harness construct the same objects and step them.
#include <cstring>
#include <string>
#include <vector>

#include "comm.hpp"
#include "data.hpp"
#include "options.hpp"
#include "utilities.hpp"
#include "vamp.hpp"

extern "C" {

// Options from an argv-style array (same parser as the executables)
void* gvbh_options_create(int argc, char** argv) { return new Options(argc, argv); }
void gvbh_options_destroy(void* o) { delete static_cast<Options*>(o); }

void* gvbh_data_create_file(const char* phen, const char* bed, int N, int M, int Mt, int S, double alpha_scale) {
    return new data(std::string(phen), std::string(bed), N, M, Mt, S, gvb_host::world().rank, "bed", alpha_scale, "");
}
void* gvbh_data_create_resident(gvb_ctx* ctx, const double* y, int N, int M, int Mt, int S, double alpha_scale) {
    return new data(ctx, std::vector<double>(y, y + N), N, M, Mt, S, gvb_host::world().rank, alpha_scale);
}
void gvbh_data_destroy(void* d) { delete static_cast<data*>(d); }
// covariates from a row-major N x C host array (the executables read them with data::read_covariates)
void gvbh_data_set_covs(void* d, const double* Z, int N, int C) {
    std::vector<std::vector<double>> rows(N, std::vector<double>(C));
    for (int i = 0; i < N; i++)
        for (int j = 0; j < C; j++) rows[i][j] = Z[(size_t)i * C + j];
    static_cast<data*>(d)->set_covs(std::move(rows));
}
gvb_ctx* gvbh_data_ctx(void* d) { return static_cast<data*>(d)->device(); }
void gvbh_data_stats(void* d, double* mave, double* msig) {
    data* p = static_cast<data*>(d);
    memcpy(mave, p->get_mave(), sizeof(double) * p->get_M());
    memcpy(msig, p->get_msig(), sizeof(double) * p->get_M());
}
void gvbh_data_Ax(void* d, double* v, double* out) {
    std::vector<double> r = static_cast<data*>(d)->Ax(v);
    memcpy(out, r.data(), r.size() * sizeof(double));
}
void gvbh_data_ATx(void* d, double* u, double* out) {
    std::vector<double> r = static_cast<data*>(d)->ATx(u);
    memcpy(out, r.data(), r.size() * sizeof(double));
}

void* gvbh_vamp_create(void* opt, int M, double gam1, double gamw) {
    Options* o = static_cast<Options*>(opt);
    return new vamp(M, gam1, gamw, std::vector<double>(M, 0.0), gvb_host::world().rank, *o);
}
void gvbh_vamp_destroy(void* v) { delete static_cast<vamp*>(v); }
void gvbh_vamp_infere(void* v, void* d, double* out) {
    vamp* p = static_cast<vamp*>(v);
    std::vector<double> r = p->infere(static_cast<data*>(d));
    memcpy(out, r.data(), r.size() * sizeof(double));
}
void gvbh_vamp_linear_begin(void* v, void* d) {
    vamp* p = static_cast<vamp*>(v);
    p->prepare(static_cast<data*>(d));
    p->linear_begin(static_cast<data*>(d));
}
int gvbh_vamp_linear_iteration(void* v, void* d, int it, const double* y_host, const double* r1_host) {
    vamp* p = static_cast<vamp*>(v);
    if (y_host || r1_host) p->upload_iteration_inputs(y_host, r1_host);
    return p->linear_iteration(static_cast<data*>(d), it) ? 1 : 0;
}
void gvbh_vamp_linear_end(void* v, double* out, int M) {
    std::vector<double> r = static_cast<vamp*>(v)->linear_end();
    if (out) memcpy(out, r.data(), sizeof(double) * M);
}
void gvbh_vamp_cg_iters(void* v, int* out2) {
    vamp* p = static_cast<vamp*>(v);
    out2[0] = p->last_cg_iters[0];
    out2[1] = p->last_cg_iters[1];
}
double gvbh_vamp_gamw(void* v) { return static_cast<vamp*>(v)->get_gamw(); }
int gvbh_vamp_prior(void* v, double* probs, double* vars) {
    vamp* p = static_cast<vamp*>(v);
    std::vector<double> pp = p->get_probs(), vv = p->get_vars();
    for (size_t i = 0; i < pp.size(); i++) { probs[i] = pp[i]; vars[i] = vv[i]; }
    return (int)pp.size();
}
long gvbh_vamp_sweeps(void* v) { return static_cast<vamp*>(v)->total_sweeps; }

}  // extern "C"
