// vamp.hpp -- class vamp: the gVAMP inference loop (linear and probit models).
//
// Public surface follows the reference's class vamp (vamp.hpp:78-148): two constructors, infere(),
// the scalar denoisers g1/g1d/g1_bin_class/g1d_bin_class, updatePrior, g2d_onsager, updateNoisePrec,
// lmmse_mult and precondCG_solver with std::vector arguments.  Internally every M- and N-vector of
// the loop is resident in HBM and all vector math runs through the C ABI (include/gvamp_b200.h);
// the host keeps only scalars, the L-component prior and the file output.
// Out of scope: --model robust, --use-XXT-denoiser, cross-validation, state evolution.
#pragma once
#include <future>
#include <string>
#include <tuple>
#include <vector>

#include "data.hpp"
#include "options.hpp"

class vamp {
private:
    int N, M, Mt, C, max_iter, rank, nranks;
    double gam1, gam2, gam_before, eta1, eta2;
    std::vector<double> gam1s, gam2s, R2trains;
    double tau1, tau2;
    double alpha1 = 0, alpha2 = 0;
    double rho;
    double gamw;

    std::vector<double> x1_hat, x2_hat, true_signal;   // host mirrors, refreshed when files are written
    std::vector<double> z1_hat, z2_hat;
    std::vector<double> y;
    std::vector<double> z1;
    std::vector<double> r1, r2;
    std::vector<double> p1, p2;
    std::vector<double> cov_eff;
    std::vector<double> mu_CG_last;

    std::vector<double> probs, probs_before;
    std::vector<double> vars, vars_before;

    double gamma_min = 1e-11;
    double gamma_max = 1e11;
    double probit_var;
    int EM_max_iter;
    double EM_err_thr;
    int CG_max_iter;
    int auto_var_max_iter = 5;
    int learn_vars;
    int init_est;
    long unsigned int seed;
    double stop_criteria_thr;
    double gamma_damp;

    std::string model;
    std::string out_dir;
    std::string out_name;
    std::vector<double> bern_vec;
    std::vector<double> invQ_bern_vec;

    int store_pvals = 1;
    double total_comp_time = 0;
    int reverse = 1;
    int use_lmmse_damp = 0;
    int use_freeze = 0;
    int SBglob = 0, LBglob = 0, redglob = 0;

    double gam1_init;
    double gamw_init;
    std::string r1_init_file;
    std::string estimate_file;
    std::string freeze_index_file;
    std::vector<double> x1_hat_stored;
    int shard_S = 0;
    // GVB_REFERENCE_SWEEPS=1: A x2_hat and <u, A^T A Q^-1 u> by their own bed sweeps like the reference (vamp.cpp:897-915) instead
    // of as by-products of the two CG solves (3 sweeps per iteration less; same values to rounding, tests compare both)
    bool reference_sweeps = false;
    // GVB_ONSAGER_WARM=1 (opt-in; the default is the reference's zero start, vamp.cpp:884): the Onsager solve starts from the previous
    // iteration's Q^-1 u instead of zero -- the probe u is the same every iteration (mt19937{seed+S}, vamp.cpp:875), so this is the warm
    // start the reference gives its LMMSE solve (vamp.cpp:591-600), applied to its second solve.  It changes the solver's path: the
    // successive-difference Onsager exit presumes a zero start, so such a solve stops on ||r||/||rhs|| < 1e-5 only (cg.cu).
    bool onsager_warm = false;
    // The Onsager solve on the Lanczos projection of A^T A (default; GVB_ONSAGER_LANCZOS=0 or GVB_REFERENCE_SWEEPS=1: by bed sweeps).
    // The probe u is the same vector in every VAMP iteration and Krylov spaces are shift-invariant, K(tau B + gam2 I, u) = K(B, u), so ALL
    // of the reference's Onsager solves live in one Krylov space of B = A^T A.  Its Lanczos tridiagonal T (lz_a, lz_b) is built once, two
    // sweeps per step, and extended only when a solve needs more steps than any before; every solve is then the reference's
    // preconditioned CG recurrences, exit tests and log lines run literally on the (K x K) projected operator tau T + gam2 I in host
    // memory: no bed sweep.  Only scalars leave the solve (alpha2 and the trace term of updateNoisePrec), exactly what the loop consumes.
    bool onsager_lanczos = true;
    std::vector<double> lz_a, lz_b;   // T = tridiag(lz_b, lz_a, lz_b), K = lz_a.size()
    double lz_unorm = 0;              // ||u||
    bool lz_started = false, lz_exhausted = false;
    long lz_key = -1;
    void lanczos_extend(int K_needed);
    void onsager_probe(data* dataset);   // draws / uploads the probe of (seed, shard) once and restarts the Krylov space when it changed
    void lanczos_begin();                // v_0 = u / ||u||
    void lanczos_finish_step(bool have_w = false);   // the tail of a step whose A v_K sits in dev.tmpN: w = A^T (A v_K) unless it came with a dual sweep, the recurrence, v_{K+1}
    // In the first LMMSE solve after the Krylov space was started the Lanczos steps ride on the solver's iterations: each iteration's A p
    // and the step's A v_K come from one dual sweep (gvb_cg_set_companion); GVB_DUAL_SWEEP=0: every step sweeps for itself
    static int lanczos_companion(void* self, int stage, int iteration, gvb_vec* v, gvb_vec* av, gvb_vec* w);
    void lanczos_ride(data* dataset);    // registers the companion for the next solve when there is a Krylov space to build
    int onsager_projected(double gam2, double tau, double* d3);
    void print_cg_log(const std::vector<double>& log, int iters, int denoiser);
    bool onsager_valid = false;
    double onsager_u_AtA_invq = 0;
    // sums of the iteration's batched reductions (gvb_vec_reduce_batch), consumed by err_measures / updateNoisePrec:
    // <est,truth>, ||est||^2, ||truth||^2, ||est/sqrt(N) - truth||^2, ||y - A est||^2, ||y||^2
    double diag_sums[6] = {0, 0, 0, 0, 0, 0};
    bool diag_valid = false;
    double noise_res_norm2 = 0;   // ||A x2_hat - y||^2
    bool noise_res_valid = false;
    bool extra_diagnostics = false;   // GVB_DIAG=1: the reference's "onsager approx"/polynomial prints (3 extra sweeps)

    // ---- device-resident state (HBM) ----
    struct Dev {
        gvb_ctx* ctx = nullptr;
        gvb_vec r1 = nullptr, r2 = nullptr, r2_prev = nullptr, x1 = nullptr, x1_prev = nullptr, x2 = nullptr, mu_last = nullptr;
        gvb_vec rhs = nullptr, bern = nullptr, invq = nullptr, tmpM = nullptr, truth = nullptr;
        gvb_vec y = nullptr, z1 = nullptr, tmpN = nullptr, tmpN2 = nullptr;
        gvb_vec p1 = nullptr, p2 = nullptr, z1h = nullptr, z2h = nullptr, mcov = nullptr, p1_prev = nullptr;
        long layout_gen = -1;       // gvb_layout_generation() the vectors and cached by-products were made for
        bool ax_x2_valid = false;   // tmpN2 holds Ax(x2) of the current iteration
        gvb_vec ata_x2 = nullptr;   // A^T A x2_hat, by-product of the LMMSE solve; with tmpN2 it warm-starts the next solve sweep-free
        int warm_age = -1;          // solves since tmpN2 / ata_x2 were last seeded by real sweeps (-1: not seeded)
        gvb_vec ax_invq = nullptr, ata_invq = nullptr;   // the same by-products for the Onsager solve (GVB_ONSAGER_WARM=1)
        int onsager_age = -1;
        gvb_vec aty = nullptr;      // A^T y: y is constant over the linear model's iterations, so the reference's per-iteration
        bool aty_valid = false;     // sweep (vamp.cpp:588) is done once and reused until y is uploaded again
        gvb_vec lz_prev = nullptr, lz_cur = nullptr, lz_w = nullptr;   // Lanczos vectors v_{K-1}, v_K and scratch (onsager_projected)
        gvb_vec ata_bern = nullptr; // A^T A bern: the probe is the same vector in every iteration, so the operator product of the first CG
        int ata_bern_state = 0;     // iteration of the zero-started Onsager solve comes from this cache (gvb_cg_solve_cached): 2 sweeps less
        bool bern_valid = false;    // dev.bern holds the Onsager probe of (seed, shard): the reference re-draws the SAME probe every
        long bern_key = -1;         // iteration (mt19937{seed + S}, vamp.cpp:875-882), so it is drawn and uploaded once
    } dev;
    // the iteration's output vectors leave the device as asynchronous snapshots (gvb_snapshot_begin) while the LMMSE sweeps run;
    // flush_outputs() at the end of the iteration hands the landed pinned buffers to a background task that scales and writes
    // them under the next iteration's kernels (two sets of snapshot slots alternate); GVB_ASYNC_OUT=0: synchronous, in place
    enum { SNAP_Z1 = 0, SNAP_X1, SNAP_R1, SNAP_R2, SNAP_X2, SNAP_COUNT };
    bool async_outputs = true;
    bool dual_sweep = true;   // z1 = A x1_hat shares a bed read with the first product of the LMMSE solve (GVB_DUAL_SWEEP=0: two sweeps)
    int snap_set = 0;
    std::string snap_path[SNAP_COUNT];
    bool snap_open[SNAP_COUNT] = {false, false, false, false, false};
    std::vector<double> out_scratch;   // the scaled copy of r1 / r2 / x2_hat on its way to a file
    // a stored value is h / scale * mul: (sqrt(N), 1) in the linear model, (1, 1/sqrt(N)) in the probit model - the reference's own arithmetic
    void emit_output(int which, gvb_vec v, size_t n, const std::string& path, double scale, int S, double mul = 1.0);
    void finish_output(int which, const double* h, size_t n, const std::string& path, double scale, int S, double mul);
    void flush_outputs(double scale, int S, double mul = 1.0);
    std::future<void> writer;   // the previous iteration's outputs, scaled and written while this iteration's kernels run
    void wait_writes();
    void dev_open(data* dataset);
    void dev_close();
    void dev_denoise(double g1_prec, double* sum_d, double* dist2);
    int dev_cg(gvb_vec rhs, gvb_vec mu, double tau, int denoiser, gvb_vec ax_mu = nullptr, double* dots3 = nullptr, gvb_vec ata_mu = nullptr,
               int have_start = 0, gvb_vec ata_rhs = nullptr, int* ata_rhs_state = nullptr);
    void sync_host(gvb_vec v, std::vector<double>& h, size_t n);
    void store_scaled(gvb_vec v, const std::string& path, double div, int S);

public:
    vamp(int N, int M, int Mt, double gam1, double gamw, int max_iter, double rho, std::vector<double> vars, std::vector<double> probs,
         std::vector<double> true_signal, int rank, std::string out_dir, std::string out_name, std::string model, Options opt = Options());
    vamp(int M, double gam1, double gamw, std::vector<double> true_signal, int rank, Options opt);
    ~vamp();

    std::vector<double> infere(data* dataset);
    std::vector<double> infere_linear(data* dataset);
    std::vector<double> infere_bin_class(data* dataset);
    // the linear loop in three pieces so that a harness can time single iterations: infere_linear() is
    // linear_begin(); for (it) if (linear_iteration(it)) break; linear_end()
    void linear_begin(data* dataset);
    bool linear_iteration(data* dataset, int it);   // true when the stopping rule fired
    std::vector<double> linear_end();
    void prepare(data* dataset);                    // the scaling / checks infere() does before the loop
    // refresh the iteration's inputs from host memory (the reference keeps y and r1 in host RAM)
    void upload_iteration_inputs(const double* y_host, const double* r1_host);

    double g1(double x, double gam1);
    double g1_bin_class(double p, double tau1, double y, double m_cov);
    double g1d(double x, double gam1);
    double g1d_bin_class(double p, double tau1, double y, double m_cov);
    double g2d_onsager(double gam2, double tau, data* dataset);

    void updatePrior(int verbose);
    void updateNoisePrec(data* dataset);
    // XXT form of the LMMSE step (denoiserXXT.cpp:15-135), --use-XXT-denoiser 1
    std::vector<double> lmmse_multAAT(std::vector<double> u, double tau, data* dataset);
    std::vector<double> lmmse_denoiserAAT(std::vector<double> r2, std::vector<double> mu_CG_AAT_last, data* dataset);
    std::vector<double> CG_solverAAT(std::vector<double> v, std::vector<double> mu_start, double tau, int save, data* dataset);

    std::vector<double> lmmse_mult(std::vector<double> v, double tau, data* dataset, int red = 0);
    std::vector<double> precondCG_solver(std::vector<double> v, double tau, int denoiser, data* dataset, int red = 0);
    std::vector<double> precondCG_solver(std::vector<double> v, std::vector<double> mu_start, double tau, int denoiser, data* dataset, int red = 0);

    void err_measures(data* dataset, int ind);
    void probit_err_measures(data* dataset, int sync, std::vector<double> true_signal, std::vector<double> est, std::string var_name);

    std::vector<double> grad_cov(std::vector<double> y, std::vector<double> gg, double probit_var, std::vector<std::vector<double>> Z,
                                 std::vector<double> eta);
    double mlogL_probit(std::vector<double> y, std::vector<double> gg, double probit_var, std::vector<std::vector<double>> Z,
                        std::vector<double> eta);
    std::vector<double> newton_cov(const std::vector<double>& y, const std::vector<double>& gg, const std::vector<std::vector<double>>& Z, std::vector<double> eta);
    std::vector<double> Newton_method_cov(std::vector<double> y, std::vector<double> gg, std::vector<std::vector<double>> Z,
                                          std::vector<double> eta);

    void set_SBglob(int SB) { SBglob = SB; }
    void set_LBglob(int LB) { LBglob = LB; }
    void set_gam2(double gam) { gam2 = gam; }
    std::vector<double> get_cov_eff() const { return cov_eff; }

    // state accessors used by the tests / the C shim (host_capi.cpp)
    std::vector<double> get_probs() const { return probs; }
    std::vector<double> get_vars() const { return vars; }
    double get_gamw() const { return gamw; }
    double get_gam1() const { return gam1; }
    double get_gam2() const { return gam2; }
    double get_alpha2() const { return alpha2; }
    const std::vector<double>& get_gam1s() const { return gam1s; }
    const std::vector<double>& get_R2trains() const { return R2trains; }
    int last_cg_iters[2] = {0, 0};   // main solve / Onsager solve of the last iteration
    long total_sweeps = 0;
};
