// options.cpp -- table-driven parser for the gVAMP command line (behavioural match of the reference's
// options.cpp:18-429: same flag names, "--flag value" form, echo of the parsed options from rank 0,
// FATAL message + exit(1) on an unknown flag, a missing value, a failed range check or a missing
// phenotype file; --out-dir is created when absent).
#include "options.hpp"

#include <sys/stat.h>

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <sstream>

#include "comm.hpp"

namespace {

enum class Kind { Str, UInt, Real, RealList, IntPair, FileList, OutDir };

struct Flag {
    const char* name;
    Kind kind;
    void* target;
    int min_value;            // UInt only: smallest accepted atoi() value
    const char* shown_name;   // name used inside the range-check message (the reference is inconsistent here)
    const char* requirement;
};

void die_range(const Flag& f, const char* got) {
    std::cout << "FATAL  : option " << f.shown_name << " has to be " << f.requirement << "! (" << got << " was passed)" << std::endl;
    exit(EXIT_FAILURE);
}

void split_reals(const std::string& s, std::vector<double>& out) {
    std::stringstream ss(s);
    std::string item;
    while (getline(ss, item, ',')) out.push_back(atof(item.c_str()));
}

}  // namespace

void Options::fail_if_last(char** argv, const int i) {
    std::cout << "FATAL  : missing argument for last option \"" << argv[i] << "\". Please check your input and relaunch." << std::endl;
    exit(EXIT_FAILURE);
}

void Options::read_command_line_options(int argc, char** argv) {
    const char* nonneg = "a non-negative integer";
    const char* pos = "a strictly positive integer";
    const char* anint = "an integer";
    const Flag table[] = {
        {"--bed-file", Kind::Str, &bed_file, 0, "", ""},
        {"--cov-file", Kind::Str, &cov_file, 0, "", ""},
        {"--bed-file-test", Kind::Str, &bed_file_test, 0, "", ""},
        {"--estimate-file", Kind::Str, &estimate_file, 0, "", ""},
        {"--freeze-index-file", Kind::Str, &freeze_index_file, 0, "", ""},
        {"--cov-estimate-file", Kind::Str, &cov_estimate_file, 0, "", ""},
        {"--run-mode", Kind::Str, &run_mode, 0, "", ""},
        {"--phen-files", Kind::FileList, &phen_files, 0, "", ""},
        {"--true-signal-files", Kind::FileList, &true_signal_files, 0, "", ""},
        {"--phen-files-test", Kind::FileList, &phen_files_test, 0, "", ""},
        {"--vars", Kind::RealList, &vars, 0, "", ""},
        {"--probs", Kind::RealList, &probs, 0, "", ""},
        {"--test-iter-range", Kind::IntPair, &test_iter_range, 0, "", ""},
        {"--use-lmmse-damp", Kind::UInt, &use_lmmse_damp, 0, "--use-lmmse-damp", nonneg},
        {"--use-freeze", Kind::UInt, &use_freeze, 0, "--use-freeze", nonneg},
        {"--seed", Kind::UInt, &seed, 0, "--seed", nonneg},
        {"--learn-vars", Kind::UInt, &learn_vars, 0, "--learn-vars", nonneg},
        {"--use-XXT-denoiser", Kind::UInt, &use_XXT_denoiser, 0, "--use-XXT-denoiser", nonneg},
        {"--iterations", Kind::UInt, &iterations, 1, "--iterations", pos},
        {"--num-mix-comp", Kind::UInt, &num_mix_comp, 1, "--num-mix-comp", pos},
        {"--store-pvals", Kind::UInt, &store_pvals, 0, "--store_pvals", anint},
        {"--red", Kind::UInt, &redglob, 0, "--red", anint},
        {"--init-est", Kind::UInt, &init_est, 0, "--init-est", anint},
        {"--out-dir", Kind::OutDir, &out_dir, 0, "", ""},
        {"--out-name", Kind::Str, &out_name, 0, "", ""},
        {"--model", Kind::Str, &model, 0, "", ""},
        {"--bim-file", Kind::Str, &bim_file, 0, "", ""},
        {"--stop-criteria-thr", Kind::Real, &stop_criteria_thr, 0, "", ""},
        {"--EM-err-thr", Kind::Real, &EM_err_thr, 0, "", ""},
        {"--alpha-scale", Kind::Real, &alpha_scale, 0, "", ""},
        {"--rho", Kind::Real, &rho, 0, "", ""},
        {"--gamma-damp", Kind::Real, &gamma_damp, 0, "", ""},
        {"--gam1-init", Kind::Real, &gam1_init, 0, "", ""},
        {"--gamw-init", Kind::Real, &gamw_init, 0, "", ""},
        {"--probit-var", Kind::Real, &probit_var, 0, "", ""},
        {"--h2", Kind::Real, &h2, 0, "", ""},
        {"--EM-max-iter", Kind::UInt, &EM_max_iter, 1, "--EM-max-iter", pos},
        {"--Mt", Kind::UInt, &Mt, 1, "--Mt", pos},
        {"--CV", Kind::UInt, &CV, 0, "--CV", "a positive integer"},
        {"--C", Kind::UInt, &C, 0, "--C", nonneg},
        {"--N", Kind::UInt, &N, 1, "--N", pos},
        {"--N-test", Kind::UInt, &N_test, 1, "--N_test", pos},
        {"--Mt-test", Kind::UInt, &Mt_test, 1, "--Mt_test", pos},
        {"--CG-max-iter", Kind::UInt, &CG_max_iter, 1, "--CG-max-iter", pos},
    };

    std::stringstream echo;
    echo << "\nardyh command line options:\n";

    for (int i = 1; i < argc; ++i) {
        const Flag* f = nullptr;
        for (const Flag& cand : table)
            if (!strcmp(argv[i], cand.name)) { f = &cand; break; }
        if (!f) {
            std::cout << "FATAL: option \"" << argv[i] << "\" unknown\n";
            exit(EXIT_FAILURE);
        }
        if (i == argc - 1) fail_if_last(argv, i);
        const char* val = argv[++i];
        switch (f->kind) {
            case Kind::Str:
                *static_cast<std::string*>(f->target) = val;
                echo << f->name << " " << val << "\n";
                break;
            case Kind::OutDir: {
                *static_cast<std::string*>(f->target) = val;
                struct stat st;
                if (stat(val, &st) != 0) mkdir(val, 0755);
                echo << f->name << " " << val << "\n";
                break;
            }
            case Kind::UInt: {
                if (atoi(val) < f->min_value) die_range(*f, val);
                unsigned int v = (unsigned int)atoi(val);
                *static_cast<unsigned int*>(f->target) = v;
                echo << f->name << " " << v << "\n";
                break;
            }
            case Kind::Real: {
                double v = atof(val);
                *static_cast<double*>(f->target) = v;
                echo << f->name << " " << v << "\n";
                break;
            }
            case Kind::RealList:
                split_reals(val, *static_cast<std::vector<double>*>(f->target));
                echo << f->name << " " << val << "\n";
                break;
            case Kind::IntPair: {
                std::vector<int>& r = *static_cast<std::vector<int>*>(f->target);
                std::stringstream ss(val);
                std::string item;
                size_t k = 0;
                while (getline(ss, item, ',') && k < r.size()) r[k++] = atoi(item.c_str());
                echo << f->name << " " << val << "\n";
                break;
            }
            case Kind::FileList: {
                std::vector<std::string>& files = *static_cast<std::vector<std::string>*>(f->target);
                echo << f->name << " " << val << "\n";
                std::stringstream ss(val);
                std::string path;
                while (getline(ss, path, ',')) {
                    std::ifstream probe(path);
                    if (!probe.is_open()) {
                        std::cout << "FATAL: file " << path << " not found\n";
                        exit(EXIT_FAILURE);
                    }
                    files.push_back(path);
                }
                break;
            }
        }
    }
    if (gvb_host::is_root()) std::cout << echo.str() << std::endl;
}

void Options::list_phen_files() const {
    for (const std::string& p : phen_files) std::cout << " phen file: " << p << std::endl;
}

// minimal setup: a bed file (training or test)
void Options::check_options() {
    if (get_bed_file() == "" && get_bed_file_test() == "") {
        std::cout << "FATAL  : no bed file provided! Please use the --bed-file option." << std::endl;
        exit(EXIT_FAILURE);
    }
}
