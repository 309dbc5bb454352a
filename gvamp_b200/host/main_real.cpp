// main_real.cpp -- driver for the linear model: the reference's main_real.exe command line
// (main_real.cpp:13-599) on the B200 hot path.  Run modes built here: infere, test, both, restart,
// predict_single, pvals-calc.  predict (Gibbs-sample prediction files) is post-processing outside the hot path
//  and reports that instead of running.
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "comm.hpp"
#include "data.hpp"
#include "options.hpp"
#include "utilities.hpp"
#include "vamp.hpp"

namespace {

struct Shard { int M, S; };
Shard shard_of(int Mt) {
    std::vector<double> ms = divide_work(Mt);
    return {(int)ms[0], (int)ms[1]};
}

std::vector<double> read_estimate(const std::string& file, int M, int S) {
    std::string ext = file.substr(file.find(".") + 1);
    return ext == "bin" ? mpi_read_vec_from_file(file, M, S) : read_vec_from_file(file, M, S);
}

// R2 of a prediction on the test set: 1 - ||y - z||^2 / (var(y) * n)
double test_r2(const std::vector<double>& y, const std::vector<double>& z, int n, double* err2_out = nullptr) {
    double err2 = 0;
    for (int i = 0; i < n; i++) err2 += (y[i] - z[i]) * (y[i] - z[i]);
    double sd = calc_stdev(y);
    if (err2_out) *err2_out = err2;
    return 1 - err2 / (sd * sd * y.size());
}

double initial_gamw(const Options& opt) { return opt.get_h2() == -1 ? 2 : 1.0 / (1.0 - opt.get_h2()); }

int run_infere(const Options& opt, int rank, bool restart) {
    Shard sh = shard_of(opt.get_Mt());
    data dataset(opt.get_phen_files()[0], opt.get_bed_file(), opt.get_N(), sh.M, opt.get_Mt(), sh.S, rank, "bed", opt.get_alpha_scale(), opt.get_bim_file());
    double gam1 = restart ? opt.get_gam1_init() : 1e-6;
    double gamw = restart ? opt.get_gamw_init() : initial_gamw(opt);
    vamp emvamp(sh.M, gam1, gamw, std::vector<double>(sh.M, 0.0), rank, opt);
    emvamp.infere(&dataset);
    return 0;
}

int run_test(const Options& opt, int rank) {
    const int N_test = opt.get_N_test();
    Shard sh = shard_of(opt.get_Mt_test());
    data dataset_test(opt.get_phen_files_test()[0], opt.get_bed_file_test(), N_test, sh.M, opt.get_Mt_test(), sh.S, rank, "bed", opt.get_alpha_scale(),
                      opt.get_bim_file());
    std::vector<double> y_test = dataset_test.get_phen();
    std::string est = opt.get_estimate_file();
    std::string ext = est.substr(est.find(".") + 1);
    std::vector<int> range = opt.get_test_iter_range();
    if (rank == 0) std::cout << "iter range = [" << range[0] << ", " << range[1] << "]" << std::endl;
    auto predict = [&](const std::string& file) {
        std::vector<double> x = read_estimate(file, sh.M, sh.S);
        x.resize(sh.M, 0.0);
        for (double& v : x) v *= sqrt((double)N_test);
        return dataset_test.Ax(x.data());
    };
    if (range[0] != -1) {
        double maxR2 = -1;
        int maxind = -1;
        size_t pos_it = est.rfind("it");
        for (int it = range[0]; it <= range[1]; it++) {
            std::string file = est.substr(0, pos_it) + "it_" + std::to_string(it) + "." + ext;
            double R2 = test_r2(y_test, predict(file), N_test);
            if (rank == 0) std::cout << R2 << ", ";
            if (R2 > maxR2) { maxR2 = R2; maxind = it; }
        }
        if (rank == 0) {
            std::cout << std::endl << "max R2 = " << maxR2 << std::endl;
            std::cout << std::endl << "max ind = " << maxind << std::endl;
        }
    } else {
        if (rank == 0) std::cout << "est_file_name = " << est << std::endl;
        double err2 = 0;
        double R2 = test_r2(y_test, predict(est), N_test, &err2);
        double sd = calc_stdev(y_test);
        if (rank == 0) {
            std::cout << "y stdev^2 = " << sd * sd << std::endl;
            std::cout << "test l2 pred err^2 = " << err2 << std::endl;
            std::cout << "test R2 = " << R2 << std::endl;
        }
    }
    return 0;
}

int run_both(const Options& opt, int rank) {
    Shard sh = shard_of(opt.get_Mt());
    data dataset(opt.get_phen_files()[0], opt.get_bed_file(), opt.get_N(), sh.M, opt.get_Mt(), sh.S, rank, "bed", opt.get_alpha_scale(), opt.get_bim_file());
    vamp emvamp(sh.M, 1e-6, initial_gamw(opt), std::vector<double>(sh.M, 0.0), rank, opt);
    std::vector<double> x_est = emvamp.infere(&dataset);
    double intercept = dataset.get_intercept(), scale = dataset.get_scale();
    if (rank == 0) {
        std::cout << "intercept = " << intercept << std::endl;
        std::cout << "scale = " << scale << std::endl;
    }
    const int N_test = opt.get_N_test();
    shard_of(opt.get_Mt());
    data dataset_test(opt.get_phen_files_test()[0], opt.get_bed_file_test(), N_test, sh.M, opt.get_Mt_test(), sh.S, rank, "bed", opt.get_alpha_scale(),
                      opt.get_bim_file());
    for (double& v : x_est) v *= sqrt((double)N_test);
    std::vector<double> z_test = dataset_test.Ax(x_est.data());
    for (double& v : z_test) v = intercept + scale * v;
    std::vector<double> y_test = dataset_test.get_phen();
    double err2 = 0;
    double R2 = test_r2(y_test, z_test, N_test, &err2);
    double sd = calc_stdev(y_test);
    if (rank == 0) {
        std::cout << std::endl;
        std::cout << "y stdev^2 = " << sd * sd << std::endl;
        std::cout << "test l2 pred err^2 = " << err2 << std::endl;
        std::cout << "test R2 = " << R2 << std::endl;
    }
    return 0;
}

int run_predict_single(const Options& opt, int rank) {
    const int N_test = opt.get_N_test();
    Shard sh = shard_of(opt.get_Mt_test());
    data dataset_test(std::vector<double>(N_test, 0.0), opt.get_bed_file_test(), N_test, sh.M, opt.get_Mt_test(), sh.S, rank, "bed", opt.get_alpha_scale(),
                      opt.get_bim_file());
    std::vector<double> x = read_estimate(opt.get_estimate_file(), sh.M, sh.S);
    x.resize(sh.M, 0.0);
    for (double& v : x) v *= sqrt((double)N_test);
    std::vector<double> z = dataset_test.Ax(x.data());
    std::string out = opt.get_out_dir() + opt.get_out_name() + "_predict.csv";
    if (rank == 0) {
        std::cout << "filepath_out = " << out << std::endl;
        store_vec_to_file(out, z);
    }
    return 0;
}

// --run-mode pvals-calc (main_real.cpp:331-452): LOO (and, with a .bim file, LOCO) association p-values of one estimate
// file or of an iteration range of estimate files; --store-pvals 0 = both, 1 = LOO only, 2 = LOCO only
int run_pvals_calc(const Options& opt, int rank) {
    const int N = opt.get_N();
    Shard sh = shard_of(opt.get_Mt());
    data dataset(opt.get_phen_files()[0], opt.get_bed_file(), N, sh.M, opt.get_Mt(), sh.S, rank, "bed", opt.get_alpha_scale(), opt.get_bim_file());
    std::string est = opt.get_estimate_file();
    std::string ext = est.substr(est.rfind(".") + 1);
    std::vector<int> range = opt.get_test_iter_range();
    if (rank == 0) std::cout << "iter range = [" << range[0] << ", " << range[1] << "]" << std::endl;
    std::vector<std::vector<double>> z1_hats, x1_hats;
    std::vector<std::string> out_loo, out_loco;
    auto add = [&](const std::string& file, const std::string& loo_name, const std::string& loco_name) {
        std::vector<double> x = ext == "bin" ? mpi_read_vec_from_file(file, sh.M, sh.S) : read_vec_from_file(file, sh.M, sh.S);
        x.resize(sh.M, 0.0);
        for (double& v : x) v *= sqrt((double)N);
        z1_hats.push_back(dataset.Ax(x.data()));
        x1_hats.push_back(x);
        if (rank == 0) std::cout << "filepath_out_pvals = " << loo_name << std::endl << "filepath_out_pvals_LOCO = " << loco_name << std::endl;
    };
    if (range[0] != -1) {
        size_t pos_it = est.rfind("it");
        for (int it = range[0]; it <= range[1]; it++) {
            std::string stem = opt.get_out_dir() + opt.get_out_name() + "_it_" + std::to_string(it);
            add(est.substr(0, pos_it) + "it_" + std::to_string(it) + "." + ext, stem, stem);   // the range form prints the stems (main_real.cpp:395-400)
            out_loo.push_back(stem + "_pvals.bin");
            out_loco.push_back(stem + "_pvals_LOCO.bin");
        }
    } else {
        if (rank == 0) std::cout << "end_est_file_name = " << ext << std::endl;
        out_loo.push_back(opt.get_out_dir() + opt.get_out_name() + "_pvals.bin");
        out_loco.push_back(opt.get_out_dir() + opt.get_out_name() + "_pvals_LOCO.bin");
        add(est, out_loo.back(), out_loco.back());   // the single-estimate form prints the full paths (main_real.cpp:429-434)
    }
    std::vector<double> y = dataset.filter_pheno();
    const unsigned store_pvals = opt.get_store_pvals();
    if (store_pvals == 0 || store_pvals == 1) dataset.pvals_calc(z1_hats, y, x1_hats, out_loo);
    if (dataset.get_bimfp() != "" && (store_pvals == 0 || store_pvals == 2)) dataset.pvals_calc_LOCO(z1_hats, y, x1_hats, out_loco);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    const gvb_host::Comm& world = gvb_host::world();
    const int rank = world.rank;
    const Options opt(argc, argv);
    const std::string mode = opt.get_run_mode();
    if (mode == "infere") return run_infere(opt, rank, false);
    if (mode == "restart") return run_infere(opt, rank, true);
    if (mode == "test") return run_test(opt, rank);
    if (mode == "both") return run_both(opt, rank);
    if (mode == "predict_single") return run_predict_single(opt, rank);
    if (mode == "pvals-calc") return run_pvals_calc(opt, rank);
    if (mode == "predict") {
        if (rank == 0) std::cout << "run mode '" << mode << "' is post-processing outside the B200 hot path and is not built" << std::endl;
        return 2;
    }
    return 0;
}
