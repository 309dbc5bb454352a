// options.hpp -- command-line options of the gVAMP drivers.  Same flags, defaults, getters and error
// behaviour as the reference's Options class (options.hpp:5-155, options.cpp:18-429) so existing
// launch scripts keep working; the parser itself is table driven.
#pragma once
#include <string>
#include <vector>

class Options {
public:
    Options() = default;
    Options(int argc, char** argv) {
        read_command_line_options(argc, argv);
        check_options();
    }

    void read_command_line_options(int argc, char** argv);

    std::string get_bed_file() const { return bed_file; }
    std::string get_bed_file_test() const { return bed_file_test; }
    std::string get_bim_file() const { return bim_file; }
    std::string get_estimate_file() const { return estimate_file; }
    std::string get_cov_estimate_file() const { return cov_estimate_file; }
    std::string get_cov_file() const { return cov_file; }
    std::string get_freeze_index_file() const { return freeze_index_file; }
    std::string get_out_dir() const { return out_dir; }
    std::string get_out_name() const { return out_name; }
    std::string get_model() const { return model; }
    std::string get_run_mode() const { return run_mode; }

    double get_stop_criteria_thr() const { return stop_criteria_thr; }
    double get_EM_err_thr() const { return EM_err_thr; }
    double get_rho() const { return rho; }
    double get_probit_var() const { return probit_var; }
    double get_h2() const { return h2; }
    double get_alpha_scale() const { return alpha_scale; }
    double get_gamw_init() const { return gamw_init; }
    double get_gam1_init() const { return gam1_init; }
    double get_gamma_damp() const { return gamma_damp; }

    unsigned int get_EM_max_iter() const { return EM_max_iter; }
    unsigned int get_CG_max_iter() const { return CG_max_iter; }
    unsigned int get_Mt() const { return Mt; }
    unsigned int get_Mt_test() const { return Mt_test; }
    unsigned int get_N() const { return N; }
    unsigned int get_N_test() const { return N_test; }
    unsigned int get_num_mix_comp() const { return num_mix_comp; }
    unsigned int get_use_lmmse_damp() const { return use_lmmse_damp; }
    unsigned int get_use_XXT_denoiser() const { return use_XXT_denoiser; }
    unsigned int get_store_pvals() const { return store_pvals; }
    unsigned int get_CV() const { return CV; }
    unsigned int get_C() const { return C; }
    unsigned int get_seed() const { return seed; }
    unsigned int get_redglob() const { return redglob; }
    unsigned int get_learn_vars() const { return learn_vars; }
    unsigned int get_init_est() const { return init_est; }
    unsigned int get_use_freeze() const { return use_freeze; }
    unsigned int get_iterations() const { return iterations; }

    std::vector<double> get_vars() const { return vars; }
    std::vector<double> get_probs() const { return probs; }
    std::vector<int> get_test_iter_range() const { return test_iter_range; }
    const std::vector<std::string>& get_phen_files() const { return phen_files; }
    const std::vector<std::string>& get_phen_files_test() const { return phen_files_test; }
    const std::vector<std::string>& get_true_signal_files() const { return true_signal_files; }

    void list_phen_files() const;
    int count_phen_files() const { return (int)phen_files.size(); }
    int count_phen_files_test() const { return (int)phen_files_test.size(); }
    void set_probit_var(double v) { probit_var = v; }

private:
    std::string bed_file, bed_file_test, estimate_file, freeze_index_file, cov_estimate_file, cov_file;
    std::string run_mode, bim_file, out_dir, out_name;
    std::string model = "linear";

    double stop_criteria_thr = 1e-4;
    double EM_err_thr = 1e-2;
    unsigned int EM_max_iter = 2;
    unsigned int CG_max_iter = 60;
    unsigned int Mt = 0, N = 0, N_test = 0, Mt_test = 0, num_mix_comp = 0;
    unsigned int store_pvals = 0, use_lmmse_damp = 0, use_XXT_denoiser = 0, use_freeze = 0;
    unsigned int learn_vars = 1, seed = 1;
    double alpha_scale = 1.0;
    unsigned int CV = 0, redglob = 0, C = 0, init_est = 0;
    double probit_var = 1;
    double gamw_init = 0;
    double gam1_init = -1;
    double gamma_damp = 1;
    std::vector<double> vars, probs;
    std::vector<int> test_iter_range = std::vector<int>(2, -1);
    double rho = 0.15;
    double h2 = -1;
    unsigned int iterations = 1;
    std::vector<std::string> phen_files, phen_files_test, true_signal_files;

    void fail_if_last(char** argv, const int i);
    void check_options();
};
