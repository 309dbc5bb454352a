// main_real_probit.cpp -- driver for the probit (bin_class) model: the reference's
// main_real_probit.exe command line (main_real_probit.cpp:13-313), run modes infere and test.
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

int main(int argc, char** argv) {
    const int rank = gvb_host::world().rank;
    const Options opt(argc, argv);
    if (opt.get_run_mode() == "infere") {
        std::vector<double> ms = divide_work(opt.get_Mt());
        int M = (int)ms[0], S = (int)ms[1];
        data dataset(opt.get_phen_files()[0], opt.get_bed_file(), opt.get_N(), M, opt.get_Mt(), S, rank, "bed");
        dataset.read_covariates(opt.get_cov_file(), opt.get_C());
        vamp emvamp(M, 1e-8, 1, std::vector<double>(M, 0.0), rank, opt);
        emvamp.infere(&dataset);
        return 0;
    }
    if (opt.get_run_mode() == "test") {
        // liability-scale prediction z = A x (+ covariate effects) and the hit rate of sign(z) against y
        const int N_test = opt.get_N_test();
        std::vector<double> ms = divide_work(opt.get_Mt_test());
        int M = (int)ms[0], S = (int)ms[1];
        data dataset_test(opt.get_phen_files_test()[0], opt.get_bed_file_test(), N_test, M, opt.get_Mt_test(), S, rank, "bed");
        std::vector<double> y_test = dataset_test.get_phen();
        std::string est = opt.get_estimate_file();
        std::string ext = est.substr(est.find(".") + 1);
        if (rank == 0) std::cout << "est_file_name = " << est << std::endl;
        std::vector<double> x = ext == "bin" ? mpi_read_vec_from_file(est, M, S) : read_vec_from_file(est, M, S);
        x.resize(M, 0.0);
        for (double& v : x) v *= sqrt((double)N_test);
        std::vector<double> z = dataset_test.Ax(x.data());
        int tp = 0, fp = 0, tn = 0, fn = 0;
        for (int i = 0; i < N_test; i++) {
            bool pred = z[i] >= 0, truth = y_test[i] > 0;
            if (pred && truth) tp++;
            else if (pred && !truth) fp++;
            else if (!pred && truth) fn++;
            else tn++;
        }
        if (rank == 0) {
            std::cout << "TPR = " << (double)tp / std::max(1, tp + fn) << std::endl;
            std::cout << "FPR = " << (double)fp / std::max(1, fp + tn) << std::endl;
        }
        return 0;
    }
    if (rank == 0) std::cout << "run mode '" << opt.get_run_mode() << "' is not built for the probit driver" << std::endl;
    return 2;
}
