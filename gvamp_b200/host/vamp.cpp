// vamp.cpp -- the gVAMP loop on a device-resident state.  Control flow, scalar formulas, exit tests,
// log lines and output files follow the reference (vamp.cpp:149-1374 for the linear model,
// vamp_probit.cpp:20-726 for the probit model); every vector operation is a kernel behind the C ABI.
#include "vamp.hpp"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <future>
#include <iomanip>
#include <iostream>
#include <numeric>
#include <random>
#include <stdexcept>

#include <nvtx3/nvToolsExt.h>

#include "comm.hpp"
#include "utilities.hpp"

using gvb_host::wtime;

namespace {
[[noreturn]] void device_fatal(const char* what) {
    std::cout << "FATAL: " << what << ": " << gvb_last_error() << std::endl;
    exit(EXIT_FAILURE);
}
#define DEV(call)                                  \
    do {                                           \
        if ((call) != GVB_OK) device_fatal(#call); \
    } while (0)

inline double clampd(double x, double lo, double hi) { return std::min(std::max(x, lo), hi); }

// NVTX range over one phase of an iteration (denoise / z1 / LMMSE CG / Onsager / noise precision / outputs): what a timeline
// tool shows as the iteration's structure; a no-op without an attached tool (nvtx3 is header-only and loads its injection lazily)
struct Phase {
    explicit Phase(const char* name) { nvtxRangePushA(name); }
    ~Phase() { nvtxRangePop(); }
};

bool files_enabled() {
    const char* e = getenv("GVB_NO_FILES");
    return !(e && e[0] == '1');
}
}  // namespace

// ---------------------------------------------------------------------------------------------------
// constructors (vamp.cpp:32-139 of the reference: same defaults, same Options plumbing)
// ---------------------------------------------------------------------------------------------------
vamp::vamp(int N, int M, int Mt, double gam1, double gamw, int max_iter, double rho, std::vector<double> vars, std::vector<double> probs,
           std::vector<double> true_signal, int rank, std::string out_dir, std::string out_name, std::string model, Options opt)
    : N(N), M(M), Mt(Mt), C(opt.get_C()), max_iter(max_iter), rank(rank), gam1(gam1), gam2(0), eta1(0), eta2(0), rho(rho), gamw(gamw),
      true_signal(true_signal), probs(probs), vars(vars), learn_vars(opt.get_learn_vars()), init_est(opt.get_init_est()), seed(opt.get_seed()),
      gamma_damp(opt.get_gamma_damp()), model(model), out_dir(out_dir), out_name(out_name), use_freeze(opt.get_use_freeze()),
      redglob(opt.get_redglob()), estimate_file(opt.get_estimate_file()), freeze_index_file(opt.get_freeze_index_file()) {
    EM_max_iter = opt.get_EM_max_iter();
    EM_err_thr = opt.get_EM_err_thr();
    CG_max_iter = opt.get_CG_max_iter();
    reverse = opt.get_use_XXT_denoiser();
    use_lmmse_damp = opt.get_use_lmmse_damp();
    stop_criteria_thr = opt.get_stop_criteria_thr();
    probit_var = opt.get_probit_var();
    gam1_init = opt.get_gam1_init();
    gamw_init = opt.get_gamw_init();
    r1_init_file = opt.get_estimate_file();
    initialize_prior(this->probs, this->vars, N, Mt, rank);
    nranks = gvb_host::world().nranks;
    const char* d = getenv("GVB_DIAG");
    extra_diagnostics = d && d[0] == '1';
    const char* rs = getenv("GVB_REFERENCE_SWEEPS");
    reference_sweeps = rs && rs[0] == '1';
    const char* ow = getenv("GVB_ONSAGER_WARM");
    onsager_warm = (ow && ow[0] == '1') && !reference_sweeps;   // opt-in: the default is the reference's zero start
    const char* ao = getenv("GVB_ASYNC_OUT");
    async_outputs = !(ao && ao[0] == '0');
    const char* ds = getenv("GVB_DUAL_SWEEP");
    dual_sweep = !(ds && ds[0] == '0');
    const char* ol = getenv("GVB_ONSAGER_LANCZOS");
    onsager_lanczos = !(ol && ol[0] == '0') && !reference_sweeps && !onsager_warm;
}

vamp::vamp(int M, double gam1, double gamw, std::vector<double> true_signal, int rank, Options opt)
    : M(M), C(opt.get_C()), rank(rank), gam1(gam1), gam2(0), eta1(0), eta2(0), rho(opt.get_rho()), gamw(gamw), true_signal(true_signal),
      probs(opt.get_probs()), learn_vars(opt.get_learn_vars()), init_est(opt.get_init_est()), seed(opt.get_seed()),
      gamma_damp(opt.get_gamma_damp()), model(opt.get_model()), out_dir(opt.get_out_dir()), out_name(opt.get_out_name()),
      store_pvals(opt.get_store_pvals()), reverse(opt.get_use_XXT_denoiser()), use_lmmse_damp(opt.get_use_lmmse_damp()),
      use_freeze(opt.get_use_freeze()), redglob(opt.get_redglob()), estimate_file(opt.get_estimate_file()),
      freeze_index_file(opt.get_freeze_index_file()) {
    N = opt.get_N();
    Mt = opt.get_Mt();
    max_iter = opt.get_iterations();
    EM_max_iter = opt.get_EM_max_iter();
    EM_err_thr = opt.get_EM_err_thr();
    CG_max_iter = opt.get_CG_max_iter();
    stop_criteria_thr = opt.get_stop_criteria_thr();
    vars = opt.get_vars();
    probit_var = opt.get_probit_var();
    gam1_init = opt.get_gam1_init();
    gamw_init = opt.get_gamw_init();
    r1_init_file = opt.get_estimate_file();
    initialize_prior(this->probs, this->vars, N, Mt, rank);
    nranks = gvb_host::world().nranks;
    const char* d = getenv("GVB_DIAG");
    extra_diagnostics = d && d[0] == '1';
    const char* rs = getenv("GVB_REFERENCE_SWEEPS");
    reference_sweeps = rs && rs[0] == '1';
    const char* ow = getenv("GVB_ONSAGER_WARM");
    onsager_warm = (ow && ow[0] == '1') && !reference_sweeps;   // opt-in: the default is the reference's zero start
    const char* ao = getenv("GVB_ASYNC_OUT");
    async_outputs = !(ao && ao[0] == '0');
    const char* ds = getenv("GVB_DUAL_SWEEP");
    dual_sweep = !(ds && ds[0] == '0');
    const char* ol = getenv("GVB_ONSAGER_LANCZOS");
    onsager_lanczos = !(ol && ol[0] == '0') && !reference_sweeps && !onsager_warm;
}

vamp::~vamp() {
    wait_writes();
    dev_close();
}

// ---------------------------------------------------------------------------------------------------
// device state
// ---------------------------------------------------------------------------------------------------
void vamp::dev_open(data* dataset) {
    if (dev.ctx == dataset->device() && dev.r1 && dev.layout_gen == gvb_layout_generation(dataset->device())) return;
    dev_close();
    dev.ctx = dataset->device();
    dev.layout_gen = gvb_layout_generation(dev.ctx);
    gvb_vec* mvecs[] = {&dev.r1, &dev.r2, &dev.r2_prev, &dev.x1, &dev.x1_prev, &dev.x2, &dev.mu_last, &dev.rhs, &dev.bern, &dev.invq, &dev.tmpM, &dev.truth, &dev.aty, &dev.ata_x2, &dev.ata_invq, &dev.ata_bern, &dev.lz_prev, &dev.lz_cur, &dev.lz_w};
    for (gvb_vec* v : mvecs) DEV(gvb_vec_alloc_M(dev.ctx, v));
    gvb_vec* nvecs[] = {&dev.y, &dev.z1, &dev.tmpN, &dev.tmpN2, &dev.ax_invq};
    for (gvb_vec* v : nvecs) DEV(gvb_vec_alloc_N(dev.ctx, v));
    if ((int)true_signal.size() == M) DEV(gvb_vec_upload(dev.ctx, dev.truth, true_signal.data(), M));
}

void vamp::dev_close() {
    wait_writes();   // a batch in flight reads the pinned snapshot buffers of the context that is being left
    lz_started = lz_exhausted = false;   // the Lanczos vectors go with the device state
    lz_a.clear();
    lz_b.clear();
    if (!dev.ctx) return;
    gvb_vec all[] = {dev.r1, dev.r2, dev.r2_prev, dev.x1, dev.x1_prev, dev.x2, dev.mu_last, dev.rhs, dev.bern, dev.invq, dev.tmpM, dev.truth, dev.aty, dev.ata_x2, dev.ata_invq, dev.ata_bern, dev.lz_prev, dev.lz_cur, dev.lz_w, dev.ax_invq,
                     dev.y, dev.z1, dev.tmpN, dev.tmpN2, dev.p1, dev.p2, dev.z1h, dev.z2h, dev.mcov, dev.p1_prev};
    for (gvb_vec v : all)
        if (v) gvb_vec_free(dev.ctx, v);
    dev = Dev();
    for (bool& open : snap_open) open = false;   // snapshots belong to the context that is being left
}

void vamp::sync_host(gvb_vec v, std::vector<double>& h, size_t n) {
    h.resize(n);
    DEV(gvb_vec_download(dev.ctx, v, h.data(), (long)n));
}

// <out><name>: raw doubles of (v / div) at byte offset S*8, like mpi_store_vec_to_file (utilities.cpp:293-301)
void vamp::store_scaled(gvb_vec v, const std::string& path, double div, int S) {
    std::vector<double> h;
    sync_host(v, h, M);
    if (!files_enabled()) return;
    for (double& x : h) x = x / div;
    store_doubles_at(path, h.data(), S, M);
}

// One output vector of the iteration (the z1 csv of vamp.cpp:435-436 and the x1_hat / r1 / r2 / x2_hat stores of
// vamp.cpp:453,462,542,612).  Asynchronous form: a snapshot starts here (device -> pinned host memory on the copy stream, under the
// LMMSE sweeps) and flush_outputs() hands the landed buffers to a background task that scales and writes them.
void vamp::emit_output(int which, gvb_vec v, size_t n, const std::string& path, double scale, int S, double mul) {
    snap_path[which] = path;
    if (async_outputs) {
        DEV(gvb_snapshot_begin(dev.ctx, v, (long)n, which + SNAP_COUNT * snap_set));
        snap_open[which] = true;
        return;
    }
    std::vector<double> h;
    sync_host(v, h, n);
    finish_output(which, h.data(), n, path, scale, S, mul);
}

// host mirror + file of one landed output.  Runs on the background task in the asynchronous form: until wait_writes() nothing else
// reads z1 / x1_hat / x1_hat_stored (their readers - the association tests, linear_end() - come after the loop).
void vamp::finish_output(int which, const double* h, size_t n, const std::string& path, double scale, int S, double mul) {
    if (which == SNAP_Z1) {
        z1.assign(h, h + n);
        if (!files_enabled() || rank != 0) return;
        // one value per line in the stream's default format (%g, 6 significant digits), like `file << v << std::endl` of the
        // reference (vamp.cpp:435-436); std::to_chars(general, 6) is that format at a fraction of the iostream cost
        // (400k values per iteration at biobank scale: tens of milliseconds of host time)
        std::string buf;
        buf.reserve(n * 14);
        char tmp[40];
        for (size_t i = 0; i < n; i++) {
            auto r = std::to_chars(tmp, tmp + sizeof(tmp), h[i], std::chars_format::general, 6);
            buf.append(tmp, r.ptr);
            buf.push_back('\n');
        }
        std::ofstream f(path, std::ios::binary);
        f.write(buf.data(), (std::streamsize)buf.size());
        return;
    }
    if (which == SNAP_X1) {
        x1_hat.assign(h, h + n);
        for (size_t i = 0; i < n; i++) x1_hat_stored[i] = h[i] / scale * mul;
        if (files_enabled()) store_doubles_at(path, x1_hat_stored.data(), S, M);
        return;
    }
    if (!files_enabled()) return;
    out_scratch.resize(n);
    for (size_t i = 0; i < n; i++) out_scratch[i] = h[i] / scale * mul;
    store_doubles_at(path, out_scratch.data(), S, M);
}

// At most one batch of outputs is in flight: it reads one of the two sets of pinned snapshot buffers while the next iteration's
// snapshots land in the other set, and wait_writes() is called before anything reads the files or the host mirrors, or the object
// goes away.  GVB_ASYNC_OUT=0: everything in place, on the calling thread.
void vamp::wait_writes() {
    if (writer.valid()) writer.get();
}

void vamp::flush_outputs(double scale, int S, double mul) {
    struct Landed {
        int which;
        const double* h;
        size_t n;
        std::string path;
    };
    std::vector<Landed> landed;
    for (int which = 0; which < SNAP_COUNT; which++) {
        if (!snap_open[which]) continue;
        const double* h = nullptr;
        long n = 0;
        DEV(gvb_snapshot_wait(dev.ctx, which + SNAP_COUNT * snap_set, &h, &n));
        snap_open[which] = false;
        landed.push_back({which, h, (size_t)n, snap_path[which]});
    }
    if (landed.empty()) return;
    wait_writes();     // the batch before this one: its buffer set is the one the next iteration snapshots into
    snap_set ^= 1;
    writer = std::async(std::launch::async, [this, landed = std::move(landed), scale, S, mul]() {
        for (const Landed& l : landed) finish_output(l.which, l.h, l.n, l.path, scale, S, mul);
    });
}

void vamp::dev_denoise(double g1_prec, double* sum_d, double* dist2) {
    double sums[2];
    DEV(gvb_denoise(dev.ctx, dev.r1, g1_prec, probs.data(), vars.data(), (int)probs.size(), dev.x1, sums));
    *sum_d = sums[0];
    *dist2 = sums[1];
}

// the solver's log in the reference's words (vamp.cpp:1185-1191, 1217-1220)
void vamp::print_cg_log(const std::vector<double>& log, int iters, int denoiser) {
    if (rank != 0) return;
    for (int i = 0; i < iters; i++) {
        if (denoiser == 0 && log[4 * i + 3] >= 0 && log[4 * i + 0] >= 0)
            std::cout << "[CG onsager] it = " << i << ": relative error for onsager is " << std::setprecision(10) << log[4 * i + 3] << std::endl;
        if (log[4 * i + 0] >= 0)
            std::cout << "[CG] it = " << i << ": ||r_it|| / ||RHS|| = " << std::setprecision(10) << log[4 * i + 0] << ", ||x_it|| = " << log[4 * i + 1]
                      << ", ||z|| / ||RHS|| = " << log[4 * i + 2] << std::endl;
    }
}

int vamp::dev_cg(gvb_vec rhs, gvb_vec mu, double tau, int denoiser, gvb_vec ax_mu, double* dots3, gvb_vec ata_mu, int have_start, gvb_vec ata_rhs,
                 int* ata_rhs_state) {
    std::vector<double> log(4 * (size_t)CG_max_iter, 0.0);
    int iters = 0;
    if (ata_rhs)
        DEV(gvb_cg_solve_cached(dev.ctx, rhs, mu, tau, gam2, CG_max_iter, denoiser, &iters, log.data(), ata_rhs, ata_rhs_state, dots3));
    else
        DEV(gvb_cg_solve_warm(dev.ctx, rhs, mu, tau, gam2, CG_max_iter, denoiser, &iters, log.data(), ax_mu, ata_mu, have_start, dots3));
    print_cg_log(log, iters, denoiser);
    return iters;
}

// ---------------------------------------------------------------------------------------------------
// entry point (vamp.cpp:149-183)
// ---------------------------------------------------------------------------------------------------
std::vector<double> vamp::infere(data* dataset) {
    prepare(dataset);
    if (!strcmp(model.c_str(), "linear")) return infere_linear(dataset);
    if (!strcmp(model.c_str(), "bin_class")) return infere_bin_class(dataset);
    throw "invalid model specification!";
}

void vamp::prepare(data* dataset) {
    y = dataset->get_phen();
    // the design matrix is scaled by 1/sqrt(N), so the effect variances are scaled by N
    for (double& v : vars) v *= N;
    if (reverse == 1) dataset->compute_people_statistics();   // vamp.cpp:168-170
    if (redglob != 0 || use_freeze != 0) {
        std::cout << "FATAL: --red / --use-freeze are not supported by the B200 build" << std::endl;
        exit(EXIT_FAILURE);
    }
    if (strcmp(model.c_str(), "linear") && strcmp(model.c_str(), "bin_class")) throw "invalid model specification!";
}

// ---------------------------------------------------------------------------------------------------
// linear model (vamp.cpp:190-803)
// ---------------------------------------------------------------------------------------------------
std::vector<double> vamp::infere_linear(data* dataset) {
    linear_begin(dataset);
    for (int it = 1; it <= max_iter; it++)
        if (linear_iteration(dataset, it)) break;
    wait_writes();
    if (store_pvals == 1) {   // association tests on the final estimate (vamp.cpp:761-777)
        std::vector<double> yf = dataset->filter_pheno();
        std::string filepath_out_pvals = out_dir + out_name + "_pvals.bin";
        dataset->pvals_calc(std::vector<std::vector<double>>{z1}, yf, std::vector<std::vector<double>>{x1_hat}, std::vector<std::string>{filepath_out_pvals});
        if (rank == 0) std::cout << "filepath_out_pvals = " << filepath_out_pvals << std::endl;
        if (dataset->get_bimfp() != "") {
            std::string filepath_out_pvals_LOCO = out_dir + out_name;
            dataset->pvals_calc_LOCO(std::vector<std::vector<double>>{z1}, yf, std::vector<std::vector<double>>{x1_hat},
                                     std::vector<std::string>{filepath_out_pvals_LOCO});
            if (rank == 0) std::cout << "filepath_out_pvals_LOCO = " << filepath_out_pvals_LOCO << std::endl;
        }
    }
    return linear_end();
}

void vamp::upload_iteration_inputs(const double* y_host, const double* r1_host) {
    if (y_host) {
        // A^T y is kept while the y that arrives is bit for bit the resident one (compared on the device; tmpN is free between
        // iterations); GVB_ATY_CACHE=0: every upload invalidates it
        const char* e = getenv("GVB_ATY_CACHE");
        const bool keep = !(e && e[0] == '0');
        int changed = 1;
        if (keep && dev.aty_valid)
            DEV(gvb_vec_upload_changed(dev.ctx, dev.y, dev.tmpN, y_host, N, &changed));
        else
            DEV(gvb_vec_upload(dev.ctx, dev.y, y_host, N));
        if (changed) dev.aty_valid = false;
    }
    if (r1_host) DEV(gvb_vec_upload(dev.ctx, dev.r1, r1_host, M));
}

void vamp::linear_begin(data* dataset) {
    dev_open(dataset);
    gvb_ctx* ctx = dev.ctx;
    const int S = dataset->get_S();
    shard_S = S;
    alpha1 = 0;
    dev.warm_age = -1;
    dev.onsager_age = -1;

    std::vector<double> yf = dataset->filter_pheno();   // NA phenotypes -> 0
    yf.resize(N, 0.0);
    DEV(gvb_vec_fill(ctx, dev.y, 0.0));
    DEV(gvb_vec_upload(ctx, dev.y, yf.data(), N));
    dev.aty_valid = false;
    DEV(gvb_vec_fill(ctx, dev.r1, 0.0));
    DEV(gvb_vec_fill(ctx, dev.x1, 0.0));
    DEV(gvb_vec_fill(ctx, dev.r2, 0.0));

    if (gam1_init != -1) {   // restart (vamp.cpp:226-233)
        gam1 = gam1_init;
        gamw = gamw_init;
        std::vector<double> r1_init = mpi_read_vec_from_file(r1_init_file, M, S);
        for (double& v : r1_init) v /= sqrt((double)N);
        DEV(gvb_vec_upload(ctx, dev.r1, r1_init.data(), M));
    }
    if (init_est == 1) {     // start from a supplied estimate (vamp.cpp:244-258)
        std::vector<double> x_est;
        std::string ext = estimate_file.substr(estimate_file.find(".") + 1);
        x_est = (ext == "bin") ? mpi_read_vec_from_file(estimate_file, M, S) : read_vec_from_file(estimate_file, M, S);
        for (double& v : x_est) v *= sqrt((double)N);
        x_est.resize(M, 0.0);
        DEV(gvb_vec_upload(ctx, dev.x1, x_est.data(), M));
        DEV(gvb_vec_upload(ctx, dev.r1, x_est.data(), M));
    }

    x1_hat_stored.assign(M, 0.0);
}

bool vamp::linear_iteration(data* dataset, int it) {
    gvb_ctx* ctx = dev.ctx;
    const int S = shard_S;
    const double scale = sqrt((double)N);
    {
        // ================= denoising step =================
        double start_denoising = wtime();
        const long sweeps0 = gvb_sweep_count(ctx);
        if (rank == 0)
            std::cout << std::endl << "********************" << std::endl << "iteration = " << it << std::endl << "********************" << std::endl
                      << "->DENOISING" << std::endl;
        nvtxRangePushA("vamp.iteration");
        nvtxRangePushA("vamp.denoise+EM");
        DEV(gvb_vec_copy(ctx, dev.x1_prev, dev.x1));
        probs_before = probs;
        vars_before = vars;
        double alpha1_prev = alpha1;
        double gam1_reEst_prev;
        int it_revar = 1;
        for (; it_revar <= auto_var_max_iter; it_revar++) {
            double sum_d = 0, dist2 = 0;
            dev_denoise(gam1, &sum_d, &dist2);
            if (it == 1 && init_est == 1) {
                DEV(gvb_vec_copy(ctx, dev.x1, dev.r1));
                dist2 = 0;
            }
            alpha1 = sum_d / Mt;
            eta1 = gam1 / alpha1;
            if (it <= 1) break;
            gam1_reEst_prev = gam1;
            gam1 = clampd(1.0 / (1.0 / eta1 + dist2 / Mt), gamma_min, gamma_max);
            updatePrior(0);
            if (rank == 0) std::cout << "[old] it_revar = " << it_revar << ": gam1 = " << gam1 << std::endl;
            if (std::fabs(gam1 - gam1_reEst_prev) < 1e-3) break;
        }
        gam1s.push_back(gam1);
        if (rank == 0) std::cout << "A total of " << std::max(it_revar - 1, 1) << " variance and prior tuning iterations were performed" << std::endl;

        if (it > 1) {   // damping (vamp.cpp:348-414)
            DEV(gvb_vec_axpby(ctx, dev.x1, rho, dev.x1, 1 - rho, dev.x1_prev));
            alpha1 = rho * alpha1 + (1 - rho) * alpha1_prev;
        }

        nvtxRangePop();
        double start_z1 = wtime();
        // The LMMSE solve's inputs depend on nothing that happens between here and the solve, so they are formed now and the solve's first
        // product A p0 shares ONE bed read with z1 = A x1_hat (gvb_cg_prepare: a dual sweep); the iterations run at the reference's place
        // (gvb_cg_solve_prepared).  Not when the solve re-seeds its by-products from real sweeps, in the N-space form, or with
        // GVB_DUAL_SWEEP=0 / GVB_REFERENCE_SWEEPS=1: then z1 is a sweep of its own.
        gam_before = gam2;
        gam2 = clampd(eta1 - gam1, gamma_min, gamma_max);
        const double gam2_print = gam2;
        DEV(gvb_vec_copy(ctx, dev.r2_prev, dev.r2));
        DEV(gvb_vec_axpby_div(ctx, dev.r2, eta1, dev.x1, -gam1, dev.r1, gam2));   // r2 = (eta1 x1 - gam1 r1)/gam2
        if (use_lmmse_damp == 1) {
            double xi = std::min(2 * rho, 1.0);
            if (it > 1) gam2 = 1.0 / pow(xi / sqrt(gam2) + (1 - xi) / sqrt(gam_before), 2);
        }
        const int warm = (it == 1) ? 2 : ((dev.warm_age >= 0 && dev.warm_age < 8) ? 1 : 0);
        const bool prepared = dual_sweep && !reference_sweeps && reverse != 1 && CG_max_iter > 0 && warm != 0;
        if (prepared) {
            Phase ph("vamp.z1 = X.x1 + first product of the LMMSE solve (dual sweep)");
            if (!dev.aty_valid) {
                DEV(gvb_dATx(ctx, dev.y, dev.aty));
                dev.aty_valid = true;
            }
            DEV(gvb_vec_axpby(ctx, dev.rhs, gamw, dev.aty, gam2, dev.r2));   // v = gamw * A^T y + gam2 * r2
            if (it == 1)
                DEV(gvb_vec_fill(ctx, dev.x2, 0.0));
            else
                DEV(gvb_vec_copy(ctx, dev.x2, dev.mu_last));
            DEV(gvb_cg_prepare(ctx, dev.rhs, dev.x2, gamw, gam2, CG_max_iter, dev.tmpN2, dev.ata_x2, warm, dev.x1, dev.z1));
        } else {
            Phase ph("vamp.z1 = X.x1");
            DEV(gvb_dAx(ctx, dev.x1, dev.z1));
        }
        std::string filepath_out_z1 = out_dir + out_name + "_z1_it_" + std::to_string(it) + ".csv";
        emit_output(SNAP_Z1, dev.z1, 4 * dataset->get_mbytes(), filepath_out_z1, scale, S);
        double end_z1 = wtime();
        if (rank == 0) std::cout << "time needed to calculate z1 = " << end_z1 - start_z1 << " seconds" << std::endl;
        if (rank == 0) std::cout << "filepath_out_z1 = " << filepath_out_z1 << std::endl;
        if (rank == 0) std::cout << "rho = " << rho << std::endl;

        double start_saving = wtime();
        std::string filepath_out = out_dir + out_name + "_it_" + std::to_string(it) + ".bin";
        emit_output(SNAP_X1, dev.x1, M, filepath_out, scale, S);
        if (rank == 0) std::cout << "x1_hat filepath_out is " << filepath_out << std::endl;
        std::string filepath_out_r1 = out_dir + out_name + "_r1_it_" + std::to_string(it) + ".bin";
        emit_output(SNAP_R1, dev.r1, M, filepath_out_r1, scale, S);
        if (rank == 0) std::cout << "r1 filepath_out is " << filepath_out_r1 << std::endl;
        if (rank == 0) std::cout << "time needed to save beta1 to an external file = " << wtime() - start_saving << " seconds" << std::endl;

        if (rank == 0) {
            std::cout << "eta1 = " << eta1 << std::endl;
            std::cout << "gam2 = " << gam2_print << std::endl;
        }
        // larger damping factors if the Onsager terms allow it (vamp.cpp:501-502)
        double xi = std::min(2 * std::min(alpha1, alpha2), 1.0);
        rho = std::max(rho, xi);
        {
            // one batch for "true gam2" and everything err_measures(1) prints (vamp.cpp:505-510, 1232-1317)
            const double isn = sqrt(1.0 / (double)N);
            gvb_red_op ops[7] = {{dev.r2, dev.truth, 1.0, -scale, GVB_RED_SQ, 1},   {dev.x1, dev.truth, 0, 0, GVB_RED_DOT, 1},
                                 {dev.x1, nullptr, 0, 0, GVB_RED_DOT, 1},          {dev.truth, nullptr, 0, 0, GVB_RED_DOT, 1},
                                 {dev.x1, dev.truth, isn, -1.0, GVB_RED_SQ, 1},    {dev.y, dev.z1, 1.0, -1.0, GVB_RED_SQ, 0},
                                 {dev.y, nullptr, 0, 0, GVB_RED_DOT, 0}};
            double res[7];
            DEV(gvb_vec_reduce_batch(ctx, 7, ops, res));
            if (rank == 0) std::cout << "true gam2 = " << Mt / res[0] << std::endl;
            for (int k = 0; k < 6; k++) diag_sums[k] = res[1 + k];
            diag_valid = true;
        }
        double start_prior_up = wtime();
        if (auto_var_max_iter == 0 || it <= 1) updatePrior(1);
        if (rank == 0) std::cout << "time needed to calculate conditional expectation = " << wtime() - start_prior_up << " seconds" << std::endl;
        err_measures(dataset, 1);
        double end_denoising = wtime();
        if (rank == 0) std::cout << "denoising step took " << end_denoising - start_denoising << " seconds." << std::endl;

        // ================= LMMSE step =================
        std::string filepath_out_r2 = out_dir + out_name + "_r2_it_" + std::to_string(it) + ".bin";
        emit_output(SNAP_R2, dev.r2, M, filepath_out_r2, scale, S);
        if (rank == 0) std::cout << "r2 filepath_out is " << filepath_out_r2 << std::endl;
        double start_lmmse_step = wtime();
        if (rank == 0) std::cout << "______________________" << std::endl << "->LMMSE" << std::endl;

        double start_CG = wtime();
        nvtxRangePushA("vamp.LMMSE CG");
        if (reverse == 1) {   // XXT form: solve in N-space (vamp.cpp:599-606)
            if (it == 1) mu_CG_last = std::vector<double>(4 * dataset->get_mbytes(), 0.0);
            std::vector<double> r2_h;
            sync_host(dev.r2, r2_h, M);
            x2_hat = lmmse_denoiserAAT(r2_h, mu_CG_last, dataset);
            DEV(gvb_vec_upload(ctx, dev.x2, x2_hat.data(), M));
            DEV(gvb_vec_copy(ctx, dev.mu_last, dev.x2));
            dev.ax_x2_valid = false;
            last_cg_iters[0] = 0;
        } else {
        // v = gamw * A^T y + gam2 * r2
        if (!prepared) {
            if (!dev.aty_valid) {
                DEV(gvb_dATx(ctx, dev.y, dev.aty));
                dev.aty_valid = true;
            }
            DEV(gvb_vec_axpby(ctx, dev.rhs, gamw, dev.aty, gam2, dev.r2));
            if (it == 1)
                DEV(gvb_vec_fill(ctx, dev.x2, 0.0));
            else
                DEV(gvb_vec_copy(ctx, dev.x2, dev.mu_last));   // warm start from the previous LMMSE estimate
        }
        // A x2_hat falls out of the CG (sum of alpha_k A p_k): updateNoisePrec and err_measures(2) need no sweep of their own.
        // A^T A x2_hat falls out the same way, and since the next solve starts from this x2_hat, its initial residual needs no
        // sweep either; every 8th solve re-seeds both from real sweeps so that rounding cannot accumulate over a long run.
        if (reference_sweeps) {
            last_cg_iters[0] = dev_cg(dev.rhs, dev.x2, gamw, 1, nullptr, nullptr, nullptr, it == 1 ? 2 : 0);
        } else {
            // warm = 2: zero start (iteration 1), 1: by-products of the previous solve describe the start vector, 0: re-seed by real sweeps
            lanczos_ride(dataset);
            if (prepared) {
                std::vector<double> log(4 * (size_t)CG_max_iter, 0.0);
                int iters = 0;
                DEV(gvb_cg_solve_prepared(ctx, dev.rhs, dev.x2, gamw, gam2, CG_max_iter, 1, &iters, log.data(), dev.tmpN2, dev.ata_x2, warm, nullptr));
                print_cg_log(log, iters, 1);
                last_cg_iters[0] = iters;
            } else {
                last_cg_iters[0] = dev_cg(dev.rhs, dev.x2, gamw, 1, dev.tmpN2, nullptr, dev.ata_x2, warm);
            }
            dev.warm_age = warm == 1 ? dev.warm_age + 1 : 0;
        }
        DEV(gvb_vec_copy(ctx, dev.mu_last, dev.x2));
        dev.ax_x2_valid = !reference_sweeps;
        }
        nvtxRangePop();
        std::string filepath_out_x2 = out_dir + out_name + "_it_" + std::to_string(it) + "_x2_hat.bin";
        emit_output(SNAP_X2, dev.x2, M, filepath_out_x2, scale, S);
        if (rank == 0) std::cout << "x2_hat filepath_out is " << filepath_out_x2 << std::endl;
        if (rank == 0) std::cout << "CG took " << wtime() - start_CG << " seconds." << std::endl;

        double start_onsager = wtime();
        {
            Phase ph("vamp.Onsager CG");
            alpha2 = g2d_onsager(gam2, gamw, dataset);
        }
        if (rank == 0) std::cout << "onsager took " << wtime() - start_onsager << " seconds." << std::endl;
        if (rank == 0) std::cout << "alpha2 = " << alpha2 << std::endl;

        if (it > 1 && extra_diagnostics) {   // print-only diagnostics of the reference (vamp.cpp:646-681)
            DEV(gvb_vec_axpby(ctx, dev.tmpM, 1.0, dev.r2, -1.0, dev.r2_prev));
            gvb_vec xs[2] = {dev.x2, dev.r2}, ys[2] = {dev.tmpM, dev.tmpM};
            double d2[2];
            DEV(gvb_vec_dots(ctx, 2, xs, ys, 1, d2));
            if (rank == 0) std::cout << "onsager approx = " << d2[0] / d2[1] << std::endl;
        }
        eta2 = gam2 / alpha2;

        if (auto_var_max_iter >= 1 && it > 2) {   // re-estimation of gam2 (vamp.cpp:691-693)
            double dist2 = 0;
            DEV(gvb_vec_dist2(ctx, dev.x2, dev.r2, 1, &dist2));
            gam2 = clampd(1 / (1 / eta2 + dist2 / Mt), gamma_min, gamma_max);
        }
        if (rank == 0) std::cout << "gam2 re-est = " << gam2 << std::endl;
        gam2s.push_back(gam2);

        gam1 = clampd(eta2 - gam2, gamma_min, gamma_max);
        DEV(gvb_vec_axpby_div(ctx, dev.r1, eta2, dev.x2, -gam2, dev.r2, gam1));   // r1 = (eta2 x2 - gam2 r2)/gam1
        if (rank == 0) std::cout << "gam1 = " << gam1 << std::endl;
        double stop_dd = 0, stop_nn = 0;
        {
            // one batch for "true gam1", the residual of updateNoisePrec, everything err_measures(2) prints and the stopping rule
            // (vamp.cpp:709-713, 892-927, 1232-1317, 741-749)
            if (!dev.ax_x2_valid) {
                DEV(gvb_dAx(ctx, dev.x2, dev.tmpN2));
                dev.ax_x2_valid = true;
            }
            const double isn = sqrt(1.0 / (double)N);
            gvb_red_op ops[9] = {{dev.r1, dev.truth, 1.0, -scale, GVB_RED_SQ, 1},  {dev.x2, dev.truth, 0, 0, GVB_RED_DOT, 1},
                                 {dev.x2, nullptr, 0, 0, GVB_RED_DOT, 1},         {dev.truth, nullptr, 0, 0, GVB_RED_DOT, 1},
                                 {dev.x2, dev.truth, isn, -1.0, GVB_RED_SQ, 1},   {dev.x1_prev, dev.x1, 1.0, -1.0, GVB_RED_SQ, 1},
                                 {dev.x1_prev, nullptr, 0, 0, GVB_RED_DOT, 1},    {dev.tmpN2, dev.y, 1.0, -1.0, GVB_RED_SQ, 0},
                                 {dev.y, nullptr, 0, 0, GVB_RED_DOT, 0}};
            double res[9];
            DEV(gvb_vec_reduce_batch(ctx, 9, ops, res));
            if (rank == 0) std::cout << "true gam1 = " << Mt / res[0] << std::endl;
            diag_sums[0] = res[1]; diag_sums[1] = res[2]; diag_sums[2] = res[3]; diag_sums[3] = res[4];
            diag_sums[4] = res[7]; diag_sums[5] = res[8];
            diag_valid = true;
            noise_res_norm2 = res[7];
            noise_res_valid = true;
            stop_dd = res[5];
            stop_nn = res[6];
        }

        {
            Phase ph("vamp.noise precision + outputs");
            updateNoisePrec(dataset);
            err_measures(dataset, 2);
            flush_outputs(scale, S);
        }
        nvtxRangePop();   // vamp.iteration   // the snapshots of this iteration's outputs have landed long ago: scale and write them

        double end_lmmse_step = wtime();
        if (rank == 0) std::cout << "lmmse step took " << end_lmmse_step - start_lmmse_step << " seconds." << std::endl;
        total_sweeps += gvb_sweep_count(ctx) - sweeps0;
        if (rank == 0 && (extra_diagnostics || getenv("GVB_LOG_SWEEPS"))) std::cout << "bed sweeps this iteration = " << gvb_sweep_count(ctx) - sweeps0 << std::endl;

        // stopping rule (vamp.cpp:741-749); the two sums came with the end-of-iteration batch
        const double dd = stop_dd, nn = stop_nn;
        if (it > 1 && sqrt(dd / nn) < stop_criteria_thr) {
            if (rank == 0) std::cout << "VAMP stopping criteria fulfilled with threshold = " << stop_criteria_thr << "." << std::endl;
            return true;
        }
        if (rank == 0) std::cout << "total iteration time = " << end_denoising - start_denoising + end_lmmse_step - start_lmmse_step << std::endl;
        total_comp_time += end_denoising - start_denoising + end_lmmse_step - start_lmmse_step;
        if (rank == 0) std::cout << "total computation time so far = " << total_comp_time << std::endl;
        if (rank == 0) std::cout << std::endl << std::endl;
    }
    return false;
}

std::vector<double> vamp::linear_end() {
    wait_writes();
    if (files_enabled() && rank == 0) {
        store_vec_to_file(out_dir + out_name + "_gam1s.csv", gam1s);
        store_vec_to_file(out_dir + out_name + "_gam2s.csv", gam2s);
        store_vec_to_file(out_dir + out_name + "_R2trains.csv", R2trains);
    }
    if (rank == 0) {
        std::cout << "gam1s filepath_out is " << out_dir + out_name + "_gam1s.csv" << std::endl;
        std::cout << "gam2s filepath_out is " << out_dir + out_name + "_gam2s.csv" << std::endl;
        std::cout << "R2trains filepath_out is " << out_dir + out_name + "_R2trains.csv" << std::endl;
        int best = (int)(std::max_element(R2trains.begin(), R2trains.end()) - R2trains.begin());
        std::cout << "index of max train R2 iteration = " << best << std::endl;
    }
    return x1_hat_stored;
}

// ---------------------------------------------------------------------------------------------------
// scalar denoisers (vamp.cpp:805-869): host evaluations for API parity; the loop uses gvb_denoise
// ---------------------------------------------------------------------------------------------------
double vamp::g1(double x, double gam1) {
    const double sigma = 1 / gam1;
    if (sigma < 1e-10 && sigma > -1e-10) return x;
    const double eta_max = *std::max_element(vars.begin(), vars.end());
    double pk = 0, pkd = 0;
    for (size_t i = 0; i < probs.size(); i++) {
        double vs = vars[i] + sigma;
        double z = probs[i] / sqrt(vs) * exp(-0.5 * x * x * (eta_max - vars[i]) / vs / (eta_max + sigma));
        pk += z;
        pkd -= z / vs * x;
    }
    return x + sigma * pkd / pk;
}

double vamp::g1d(double x, double gam1) {
    const double sigma = 1 / gam1;
    if (sigma < 1e-10 && sigma > -1e-10) return 1;
    const double eta_max = *std::max_element(vars.begin(), vars.end());
    double pk = 0, pkd = 0, pkdd = 0;
    for (size_t i = 0; i < probs.size(); i++) {
        double vs = vars[i] + sigma;
        double ex = exp(-0.5 * x * x * (eta_max - vars[i]) / vs / (eta_max + sigma));
        double z = probs[i] / sqrt(vs) * ex;
        pk += z;
        z = z / vs * x;
        pkd -= z;
        pkdd = pkdd - probs[i] / pow(vs, 1.5) * ex + z / vs * x;
    }
    return 1 + sigma * (pkdd / pk - pow(pkd / pk, 2));
}

// ---------------------------------------------------------------------------------------------------
// Onsager correction by a Hutchinson probe (vamp.cpp:871-889)
// ---------------------------------------------------------------------------------------------------
void vamp::onsager_probe(data* dataset) {
    // Rademacher probe +-1/sqrt(Mt) from mt19937{seed + S}: identical stream to the reference on every shard
    // (the reference re-draws it every iteration from the same seed, vamp.cpp:875-882: drawn and uploaded once here)
    const long bern_key = (long)seed + (long)dataset->get_S();
    if (!dev.bern_valid || dev.bern_key != bern_key || (int)bern_vec.size() != M) {
        std::mt19937 rd{seed + (long unsigned int)dataset->get_S()};
        std::bernoulli_distribution bern(0.5);
        bern_vec.assign(M, 0.0);
        for (int i = 0; i < M; i++) bern_vec[i] = (2 * bern(rd) - 1) / sqrt(Mt);
        DEV(gvb_vec_upload(dev.ctx, dev.bern, bern_vec.data(), M));
        dev.bern_valid = true;
        dev.bern_key = bern_key;
        dev.ata_bern_state = 0;   // another probe: its A^T A product is not cached yet
    }
    if (onsager_lanczos && lz_key != bern_key) {   // another probe (or another shard): its Krylov space starts over
        lz_started = lz_exhausted = false;
        lz_a.clear();
        lz_b.clear();
        lz_key = bern_key;
    }
}

double vamp::g2d_onsager(double gam2, double tau, data* dataset) {
    dev_open(dataset);
    onsager_probe(dataset);
    this->gam2 = gam2;
    double d3[3] = {0, 0, 0};
    if (onsager_lanczos) {
        last_cg_iters[1] = onsager_projected(gam2, tau, d3);
    } else if (onsager_warm) {   // GVB_ONSAGER_WARM=1: start from the previous iteration's Q^-1 u (same probe); the solve then stops on the residual only
        const int warm = (dev.onsager_age >= 0 && dev.onsager_age < 8) ? 1 : 2;
        last_cg_iters[1] = dev_cg(dev.bern, dev.invq, tau, 0, dev.ax_invq, d3, dev.ata_invq, warm);
        dev.onsager_age = warm == 1 ? dev.onsager_age + 1 : 0;
    } else if (reference_sweeps) {   // the reference: from zero (vamp.cpp:884, 1120-1127), every operator product by its own sweeps
        last_cg_iters[1] = dev_cg(dev.bern, dev.invq, tau, 0, nullptr, d3, nullptr, 2);
    } else {              // from zero; the probe recurs, so A^T A probe is cached and CG iteration 0 needs no sweep (gvb_cg_solve_cached)
        last_cg_iters[1] = dev_cg(dev.bern, dev.invq, tau, 0, nullptr, d3, nullptr, 2, dev.ata_bern, &dev.ata_bern_state);
    }
    // <u, A^T A Q^-1 u> from the solver's own residual: Q mu = u - r  =>  tau A^T A mu = u - r - gam2 mu
    onsager_u_AtA_invq = (d3[0] - gam2 * d3[1] - d3[2]) / tau;
    onsager_valid = true;
    return gam2 * d3[1];   // gam2 * <u, Q^-1 u> = gam2 * Tr[(tau X^T X + gam2 I)^-1] / Mt; <u, mu> is the solver's own last sum
}

// ---------------------------------------------------------------------------------------------------
// Onsager solve on the Lanczos projection (see vamp.hpp).  Lanczos on B = A^T A from v_0 = u / ||u||:
//   w = B v_K ; a_K = <w, v_K> ; w -= a_K v_K + b_{K-1} v_{K-1} ; b_K = ||w|| ; v_{K+1} = w / b_K          (two bed sweeps per step)
// All dots are rank sums, so every rank holds the same T.
// ---------------------------------------------------------------------------------------------------
void vamp::lanczos_begin() {
    if (lz_started) return;
    gvb_ctx* ctx = dev.ctx;
    gvb_vec xs[1] = {dev.bern};
    double uu = 0;
    DEV(gvb_vec_dots(ctx, 1, xs, nullptr, 1, &uu));
    lz_unorm = sqrt(uu);
    DEV(gvb_vec_axpby(ctx, dev.lz_cur, 1.0 / lz_unorm, dev.bern, 0.0, nullptr));
    DEV(gvb_vec_fill(ctx, dev.lz_prev, 0.0));
    lz_started = true;
}

void vamp::lanczos_finish_step(bool have_w) {
    gvb_ctx* ctx = dev.ctx;
    const int K = (int)lz_a.size();
    if (!have_w) DEV(gvb_dATx(ctx, dev.tmpN, dev.lz_w));
    gvb_vec xs[1] = {dev.lz_w}, ys[1] = {dev.lz_cur};
    double a = 0;
    DEV(gvb_vec_dots(ctx, 1, xs, ys, 1, &a));
    DEV(gvb_vec_axpby(ctx, dev.lz_w, 1.0, dev.lz_w, -a, dev.lz_cur));
    if (K > 0) DEV(gvb_vec_axpby(ctx, dev.lz_w, 1.0, dev.lz_w, -lz_b[K - 1], dev.lz_prev));
    double ww = 0;
    DEV(gvb_vec_dots(ctx, 1, xs, nullptr, 1, &ww));
    const double b = sqrt(ww);
    lz_a.push_back(a);
    lz_b.push_back(b);
    if (!(b > 1e-13 * std::fabs(a))) {   // the Krylov space is exhausted (tiny shards): T is complete
        lz_b.back() = 0.0;
        lz_exhausted = true;
        return;
    }
    std::swap(dev.lz_prev, dev.lz_cur);                                               // v_{K-1} <- v_K
    DEV(gvb_vec_axpby(ctx, dev.lz_cur, 1.0 / b, dev.lz_w, 0.0, nullptr));     // v_K <- w / b
}

void vamp::lanczos_extend(int K_needed) {
    lanczos_begin();
    while ((int)lz_a.size() < K_needed && !lz_exhausted) {
        DEV(gvb_dAx(dev.ctx, dev.lz_cur, dev.tmpN));
        lanczos_finish_step();
    }
}

// stage 0: does the Krylov space want another step?  then its A v_K rides on this iteration's A p; stage 1: the step's tail
int vamp::lanczos_companion(void* self, int stage, int iteration, gvb_vec* v, gvb_vec* av, gvb_vec* w) {
    (void)iteration;
    vamp* me = static_cast<vamp*>(self);
    if (stage == 0) {
        if (me->lz_exhausted || (int)me->lz_a.size() >= me->CG_max_iter + 1) return 0;   // onsager_projected never asks for more
        *v = me->dev.lz_cur;
        *av = me->dev.tmpN;
        *w = me->dev.lz_w;        // both halves of the step's A^T A v_K come with the solver's sweeps
        return 1;
    }
    me->lanczos_finish_step(*w != nullptr);
    return 0;
}

void vamp::lanczos_ride(data* dataset) {
    if (!onsager_lanczos || !dual_sweep) return;
    onsager_probe(dataset);
    if (!lz_a.empty() || lz_exhausted) return;   // only the solve that meets an empty Krylov space carries its construction
    lanczos_begin();
    DEV(gvb_cg_set_companion(dev.ctx, &vamp::lanczos_companion, this));
}

// vamp::precondCG_solver(u, 0, tau, denoiser = 0) (vamp.cpp:1130-1229) in the coordinates of the Lanczos basis: u = ||u|| e_0, the operator is
// tau T + gam2 I, inner products are Euclidean; iteration i touches coordinates <= i + 1.  Returns the iterations run and
// d3 = {<u,u>, <u,mu>, <u,r>} with r the residual of the returned mu.
int vamp::onsager_projected(double gam2, double tau, double* d3) {
    const int n = CG_max_iter + 2;
    std::vector<double> mu(n, 0.0), r(n, 0.0), z(n, 0.0), p(n, 0.0), d(n, 0.0), log(4 * (size_t)std::max(CG_max_iter, 1), 0.0);
    const double diag = tau * (double)(N - 1) / (double)N + gam2;
    lanczos_extend(2);
    r[0] = lz_unorm;
    auto dot = [&](const std::vector<double>& x, const std::vector<double>& y, int m) {
        double s = 0;
        for (int j = 0; j < m; j++) s += x[j] * y[j];
        return s;
    };
    const double norm_v = lz_unorm;
    for (int j = 0; j < n; j++) p[j] = z[j] = r[j] / diag;
    double prev_onsager = 0;
    int iters = 0;
    for (int i = 0; i < CG_max_iter; i++) {
        iters = i + 1;
        lanczos_extend(i + 2);
        const int K = (int)lz_a.size();
        const int m = std::min(n, K);                                  // coordinates in play
        for (int j = 0; j < m; j++) {                                   // d = (tau T + gam2) p   (lmmse_mult, vamp.cpp:1074-1118)
            double t = lz_a[j] * p[j];
            if (j > 0) t += lz_b[j - 1] * p[j - 1];
            if (j + 1 < m) t += lz_b[j] * p[j + 1];
            d[j] = t * tau + gam2 * p[j];
        }
        const double rz = dot(r, z, m);
        const double alpha = rz / dot(d, p, m);
        for (int j = 0; j < m; j++) mu[j] += alpha * p[j];
        const double norm_mu = sqrt(dot(mu, mu, m));
        const double onsager = gam2 * (lz_unorm * mu[0]);             // gam2 <u, mu>
        const double ons_rel = (onsager != 0.0) ? std::fabs((onsager - prev_onsager) / onsager) : 1.0;
        for (int j = 0; j < m; j++) r[j] -= d[j] * alpha;              // the residual of the returned mu in every exit path
        if (ons_rel < 1e-8) {                                          // vamp.cpp:1174-1193
            log[4 * i + 0] = -1.0; log[4 * i + 1] = norm_mu; log[4 * i + 2] = -1.0; log[4 * i + 3] = ons_rel;
            break;
        }
        prev_onsager = onsager;
        double beta = 1.0 / rz;
        for (int j = 0; j < m; j++) z[j] = r[j] / diag;
        beta *= dot(r, z, m);
        for (int j = 0; j < m; j++) p[j] = z[j] + beta * p[j];
        const double rr = dot(r, r, m);
        const double rel_err = sqrt(rr) / norm_v;
        log[4 * i + 0] = rel_err; log[4 * i + 1] = norm_mu; log[4 * i + 2] = sqrt(rr) / diag / norm_v; log[4 * i + 3] = ons_rel;
        if (rel_err < 1e-5) break;                                     // vamp.cpp:1217-1223
    }
    print_cg_log(log, iters, 0);
    d3[0] = lz_unorm * lz_unorm;
    d3[1] = lz_unorm * mu[0];
    d3[2] = lz_unorm * r[0];
    return iters;
}

// ---------------------------------------------------------------------------------------------------
// XXT form of the LMMSE step (denoiserXXT.cpp:15-55): x2 = r2 + gamw A^T (gamw A A^T + gam2)^-1 (y - A r2)
// ---------------------------------------------------------------------------------------------------
std::vector<double> vamp::lmmse_multAAT(std::vector<double> u, double tau, data* dataset) {
    const size_t phen_size = 4 * dataset->get_mbytes();
    u.resize(phen_size, 0.0);
    bool zero = true;
    for (int i = 0; i < N && zero; i++) zero = (u[i] == 0.0);
    if (zero) return std::vector<double>(phen_size, 0.0);
    std::vector<double> t = dataset->ATx(u.data());
    std::vector<double> res = dataset->Ax(t.data());
    for (int i = 0; i < N; i++) res[i] = res[i] * tau + gam2 * u[i];
    return res;
}

std::vector<double> vamp::CG_solverAAT(std::vector<double> v, std::vector<double> mu_start, double tau, int save, data* dataset) {
    dev_open(dataset);
    gvb_ctx* ctx = dev.ctx;
    const long n4 = 4 * (long)dataset->get_mbytes();
    gvb_vec dv = nullptr, dmu = nullptr;
    DEV(gvb_vec_alloc_N(ctx, &dv));
    DEV(gvb_vec_alloc_N(ctx, &dmu));
    v.resize(n4, 0.0);
    mu_start.resize(n4, 0.0);
    for (long i = N; i < n4; i++) v[i] = mu_start[i] = 0.0;
    DEV(gvb_vec_upload(ctx, dv, v.data(), n4));
    DEV(gvb_vec_upload(ctx, dmu, mu_start.data(), n4));
    gvb_vec* pe = dataset->device_people();
    if (!pe[0]) {
        std::cout << "FATAL: compute_people_statistics() must run before the XXT solver" << std::endl;
        exit(EXIT_FAILURE);
    }
    if (rank == 0) {   // the first entries of the preconditioner, as the reference prints them (denoiserXXT.cpp:68-78)
        double a[3] = {0, 0, 0}, sg[3] = {1, 1, 1}, nb[3] = {0, 0, 0};
        const long k = std::min<long>(3, N);
        DEV(gvb_vec_download(ctx, pe[0], a, k));
        DEV(gvb_vec_download(ctx, pe[1], sg, k));
        DEV(gvb_vec_download(ctx, pe[2], nb, k));
        for (long i = 0; i < k; i++)
            std::cout << "diag[" << i << "] = " << tau * ((nb[i] - 1) / sg[i] / sg[i] + a[i] * a[i] * nb[i]) / N + gam2 << std::endl;
    }
    std::vector<double> log(3 * (size_t)CG_max_iter, 0.0);
    int iters = 0;
    DEV(gvb_cg_solve_aat(ctx, dv, dmu, tau, gam2, pe[0], pe[1], pe[2], CG_max_iter, &iters, log.data()));
    if (rank == 0)
        for (int i = 0; i < iters; i++)
            std::cout << "[CG] it = " << i << ": ||r_it|| / ||RHS|| = " << log[3 * i] << ", ||x_it|| = " << log[3 * i + 1] << ", ||z|| / ||RHS|| = " << log[3 * i + 2]
                      << std::endl;
    std::vector<double> mu(n4, 0.0);
    DEV(gvb_vec_download(ctx, dmu, mu.data(), n4));
    gvb_vec_free(ctx, dv);
    gvb_vec_free(ctx, dmu);
    if (save == 1) mu_CG_last = mu;
    return mu;
}

std::vector<double> vamp::lmmse_denoiserAAT(std::vector<double> r2, std::vector<double> mu_CG_AAT_last, data* dataset) {
    double norm_r2 = sqrt(l2_norm2(r2, 1));
    if (rank == 0) std::cout << "||r2|| = " << norm_r2 << std::endl;
    std::vector<double> z2 = dataset->Ax(r2.data());
    if (rank == 0) std::cout << "||z2|| = " << sqrt(l2_norm2(z2, 0)) << std::endl;
    // the reference's member y is the NA-filtered phenotype by the time this runs (err_measures(1) has replaced it, vamp.cpp:1292)
    std::vector<double> yf = dataset->filter_pheno();
    yf.resize(N, 0.0);
    if (rank == 0) std::cout << "||y|| = " << sqrt(l2_norm2(yf, 0)) << std::endl;
    std::vector<double> v(N, 0.0);
    for (int i = 0; i < N; i++) v[i] = yf[i] - z2[i];
    if (rank == 0) std::cout << "||v|| = " << sqrt(l2_norm2(v, 0)) << std::endl;
    std::vector<double> u = CG_solverAAT(v, mu_CG_AAT_last, gamw, 1, dataset);
    std::vector<double> res = dataset->ATx(u.data());
    for (int i = 0; i < M; i++) res[i] = gamw * res[i] + r2[i];
    return res;
}

// ---------------------------------------------------------------------------------------------------
// noise precision (vamp.cpp:892-927)
// ---------------------------------------------------------------------------------------------------
void vamp::updateNoisePrec(data* dataset) {
    gvb_ctx* ctx = dev.ctx;
    if (!dev.ax_x2_valid) {
        DEV(gvb_dAx(ctx, dev.x2, dev.tmpN2));   // also reused by err_measures(2)
        dev.ax_x2_valid = true;
    }
    double temp_norm2 = 0;
    if (noise_res_valid)
        temp_norm2 = noise_res_norm2;   // ||A x2_hat - y||^2 arrived with the iteration's batch of reductions
    else
        DEV(gvb_vec_dist2(ctx, dev.tmpN2, dev.y, 0, &temp_norm2));
    noise_res_valid = false;
    double dot = 0;
    if (onsager_valid && !reference_sweeps) {
        dot = onsager_u_AtA_invq;   // by-product of the Onsager CG (g2d_onsager), no sweep
    } else {
        DEV(gvb_dAx(ctx, dev.invq, dev.tmpN));
        DEV(gvb_dATx(ctx, dev.tmpN, dev.tmpM));
        gvb_vec xs[1] = {dev.bern}, ys[1] = {dev.tmpM};
        DEV(gvb_vec_dots(ctx, 1, xs, ys, 1, &dot));
    }
    onsager_valid = false;
    double trace_corr = dot * Mt;
    if (rank == 0) {
        std::cout << "l2_norm2(temp) / N = " << temp_norm2 / N << std::endl;
        std::cout << "trace_correction / N = " << trace_corr / N << std::endl;
    }
    gamw = (double)N / (temp_norm2 + trace_corr);
    (void)dataset;
}

// ---------------------------------------------------------------------------------------------------
// EM update of the prior (vamp.cpp:929-1072): the E-step sums come from one kernel + one allreduce
// ---------------------------------------------------------------------------------------------------
void vamp::updatePrior(int verbose) {
    double lambda = 1 - probs[0];
    std::vector<double> omegas = probs;
    for (size_t j = 1; j < omegas.size(); j++) omegas[j] /= lambda;
    const int L = (int)probs.size();
    int it;
    for (it = 0; it < EM_max_iter; it++) {
        std::vector<double> probs_prev = probs, vars_prev = vars;
        std::vector<double> sums(2 * L - 1, 0.0);
        if (L >= 2) DEV(gvb_em_stats(dev.ctx, dev.r1, gam1, lambda, omegas.data(), vars.data(), L, sums.data()));
        const double sum_of_pin = sums[0];
        lambda = sum_of_pin / Mt;
        for (int j = 0; j < L - 1; j++) {
            double res = sums[1 + j], res_gammas = sums[L + j];
            if (learn_vars == 1) vars[j + 1] = res_gammas / res;
            omegas[j + 1] = res / sum_of_pin;
            probs[j + 1] = lambda * omegas[j + 1];
        }
        probs[0] = 1 - lambda;
        double dp = 0, np = 0, dv = 0, nv = 0;
        for (int j = 0; j < L; j++) {
            dp += (probs[j] - probs_prev[j]) * (probs[j] - probs_prev[j]);
            np += probs[j] * probs[j];
            dv += (vars[j] - vars_prev[j]) * (vars[j] - vars_prev[j]);
            nv += vars[j] * vars[j];
        }
        double dist_probs = sqrt(dp / np), dist_vars = sqrt(dv / nv);
        if (verbose == 1 && rank == 0) std::cout << "it = " << it << ": dist_probs = " << dist_probs << " & dist_vars = " << dist_vars << std::endl;
        if (dist_probs < EM_err_thr && dist_vars < EM_err_thr) break;
    }
    if (verbose == 1 && rank == 0) std::cout << "Final number of prior EM iterations = " << std::min(it + 1, EM_max_iter) << " / " << EM_max_iter << std::endl;

    // merge components whose variances are within 50 % of each other (vamp.cpp:1054-1071)
    for (size_t j = 0; j < vars.size(); j++) {
        for (size_t k = j + 1; k < vars.size(); k++) {
            double denom = (vars[j] != 0) ? std::min(vars[j], vars[k]) : 1e-7;
            if (std::fabs(vars[j] - vars[k]) / denom < 5e-1) {
                probs[j] += probs[k];
                vars.erase(vars.begin() + k);
                probs.erase(probs.begin() + k);
                k--;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// host-vector wrappers with the reference's signatures (vamp.cpp:1074-1229)
// ---------------------------------------------------------------------------------------------------
std::vector<double> vamp::lmmse_mult(std::vector<double> v, double tau, data* dataset, int red) {
    if (red != 0) {
        std::cout << "FATAL: reduced (--red) LMMSE products are not supported by the B200 build" << std::endl;
        exit(EXIT_FAILURE);
    }
    dev_open(dataset);
    DEV(gvb_vec_upload(dev.ctx, dev.tmpM, v.data(), M));
    gvb_vec out = nullptr;
    DEV(gvb_vec_alloc_M(dev.ctx, &out));
    DEV(gvb_lmmse_mult(dev.ctx, dev.tmpM, tau, gam2, out));
    std::vector<double> res(M);
    DEV(gvb_vec_download(dev.ctx, out, res.data(), M));
    gvb_vec_free(dev.ctx, out);
    return res;
}

std::vector<double> vamp::precondCG_solver(std::vector<double> v, double tau, int denoiser, data* dataset, int red) {
    return precondCG_solver(v, std::vector<double>(M, 0.0), tau, denoiser, dataset, red);
}

std::vector<double> vamp::precondCG_solver(std::vector<double> v, std::vector<double> mu_start, double tau, int denoiser, data* dataset, int red) {
    if (red != 0) {
        std::cout << "FATAL: reduced (--red) CG solves are not supported by the B200 build" << std::endl;
        exit(EXIT_FAILURE);
    }
    dev_open(dataset);
    gvb_vec rhs = nullptr, mu = nullptr;
    DEV(gvb_vec_alloc_M(dev.ctx, &rhs));
    DEV(gvb_vec_alloc_M(dev.ctx, &mu));
    DEV(gvb_vec_upload(dev.ctx, rhs, v.data(), M));
    DEV(gvb_vec_upload(dev.ctx, mu, mu_start.data(), M));
    dev_cg(rhs, mu, tau, denoiser);
    std::vector<double> res(M);
    DEV(gvb_vec_download(dev.ctx, mu, res.data(), M));
    if (denoiser == 1) mu_CG_last = res;
    gvb_vec_free(dev.ctx, rhs);
    gvb_vec_free(dev.ctx, mu);
    return res;
}

// ---------------------------------------------------------------------------------------------------
// error measures (vamp.cpp:1232-1374)
// ---------------------------------------------------------------------------------------------------
void vamp::err_measures(data* dataset, int ind) {
    gvb_ctx* ctx = dev.ctx;
    gvb_vec est = (ind == 1) ? dev.x1 : dev.x2;
    if (!diag_valid) {   // stand-alone call: the same sums as the iteration's batches deliver
        if (ind == 2 && !dev.ax_x2_valid) {
            DEV(gvb_dAx(ctx, dev.x2, dev.tmpN2));
            dev.ax_x2_valid = true;
        }
        const double isn = sqrt(1.0 / (double)N);
        gvb_vec ax = (ind == 1) ? dev.z1 : dev.tmpN2;
        gvb_red_op ops[6] = {{est, dev.truth, 0, 0, GVB_RED_DOT, 1},     {est, nullptr, 0, 0, GVB_RED_DOT, 1}, {dev.truth, nullptr, 0, 0, GVB_RED_DOT, 1},
                             {est, dev.truth, isn, -1.0, GVB_RED_SQ, 1}, {dev.y, ax, 1.0, -1.0, GVB_RED_SQ, 0}, {dev.y, nullptr, 0, 0, GVB_RED_DOT, 0}};
        DEV(gvb_vec_reduce_batch(ctx, 6, ops, diag_sums));
    }
    diag_valid = false;
    const double* d = diag_sums;   // <est,truth>, ||est||^2, ||truth||^2, ||est/sqrt(N) - truth||^2, ||y - A est||^2, ||y||^2
    double corr = d[0] / sqrt(d[1] * d[2]);
    if (rank == 0) std::cout << "correlation " << (ind == 1 ? "x1_hat" : "x2_hat") << " = " << corr << std::endl;
    if (rank == 0) std::cout << "l2 signal error = " << sqrt(d[3] / d[2]) << std::endl;
    // R2 = 1 - ||y - A est||^2 / ||y||^2 over the N individuals (vamp.cpp:1303-1317)
    double l2_pred_err = sqrt(d[4] / d[5]);
    if (rank == 0) std::cout << "l2 prediction error = " << l2_pred_err << std::endl;
    double R2 = 1 - l2_pred_err * l2_pred_err;
    R2trains.push_back(R2);
    if (rank == 0) {
        std::cout << "R2 = " << R2 << std::endl;
        std::cout << "prior variances = ";
        for (double v : vars) std::cout << v << ' ';
        std::cout << std::endl << "prior probabilities = ";
        for (double p : probs) std::cout << p << ' ';
        std::cout << std::endl << "gamw = " << gamw << std::endl;
    }
    (void)dataset;
}

#include "vamp_probit.inc"
