// utilities.hpp -- host-side helpers of the gVAMP drivers with the reference's names and meaning
// (utilities.hpp:11-51): work partition, vector file I/O (byte-identical formats), host reductions,
// prior initialisation, simulation helpers and special functions.
#pragma once
#include <cmath>
#include <random>
#include <string>
#include <vector>

struct gvb_ctx;

namespace gvb_host {
// context used by the host-level collectives (inner_prod(...,1), calc_stdev(...,1)); set by class data
void set_collective_ctx(gvb_ctx* ctx);
gvb_ctx* collective_ctx();
void allreduce_sum(double* buf, int n);   // MPI_Allreduce(SUM) replacement, NCCL under the hood
}  // namespace gvb_host

void initialize_prior(std::vector<double>& probs, std::vector<double>& vars, int N, int Mt, int rank);
double generate_mixture_gaussians(int K_grp, const std::vector<double>& eta, const std::vector<double>& pi, long unsigned int seed = 1);
std::vector<double> simulate(int M, std::vector<double> eta, std::vector<double> pi, long unsigned int seed = 1);
double noise_prec_calc(double SNR, std::vector<double> vars, std::vector<double> probs, int Mt, int N);

std::vector<double> read_vec_from_file(std::string filename, int M, int S);
std::vector<double> mpi_read_vec_from_file(std::string filename, int M, int S);
void store_vec_to_file(std::string filepath, std::vector<double> vec);
void mpi_store_vec_to_file(std::string filepath, std::vector<double> vec, int S, int M);
void store_doubles_at(const std::string& filepath, const double* vec, int S, int M);   // the same write without the by-value copy

double inner_prod(std::vector<double> const& u, std::vector<double> const& v, int sync);
double l2_norm2(std::vector<double> const& u, int sync);
double calc_stdev(std::vector<double> vec, int sync = 0);

std::vector<double> divide_work(int Mt);

double normal_cdf(double value);
double erfcx(double x);
int sgn(double val);
