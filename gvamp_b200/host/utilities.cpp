// utilities.cpp -- see utilities.hpp.  File formats are the reference's: text vectors with the
// default ostream precision (utilities.cpp:178-187 of the reference), binary vectors as raw doubles
// at byte offset S*8 (utilities.cpp:293-319).
#include "utilities.hpp"

#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <cassert>
#include <fstream>
#include <iostream>
#include <numeric>
#include <stdexcept>
#include <thread>

#include "comm.hpp"
#include "gvamp_b200.h"

extern "C" int gvb_allreduce_host(gvb_ctx* ctx, double* buf, int n);

namespace gvb_host {
static gvb_ctx* g_ctx = nullptr;
void set_collective_ctx(gvb_ctx* ctx) { g_ctx = ctx; }
gvb_ctx* collective_ctx() { return g_ctx; }
void allreduce_sum(double* buf, int n) {
    if (world().nranks <= 1) return;
    if (!g_ctx || gvb_allreduce_host(g_ctx, buf, n) != GVB_OK) {
        std::cout << "FATAL: host allreduce failed: " << (g_ctx ? gvb_last_error() : "no device context yet") << std::endl;
        exit(EXIT_FAILURE);
    }
}
}  // namespace gvb_host

// ---- prior -----------------------------------------------------------------------------------------
// Default prior when --probs/--vars are absent: 23 components, inclusion mass 50 000/Mt halving per
// component, variances log-spaced from 1e-5 to 1e2, divided by N.
void initialize_prior(std::vector<double>& probs, std::vector<double>& vars, int N, int Mt, int rank) {
    if (!probs.empty() || !vars.empty()) return;
    const int num_mix = 23;
    if (Mt <= 50000) throw std::invalid_argument("No probabilities or variances were specified and Mt < 50,000.");
    double p = std::min(50000.0 / Mt, 1.0) / (2 - 1.0 / pow(2, 21));
    probs.push_back(1 - 50000.0 / Mt);
    for (int k = 0; k < num_mix - 1; k++, p /= 2) probs.push_back(p);
    if (rank == 0) {
        std::cout << "probs = ";
        for (double x : probs) std::cout << x << ' ';
        std::cout << std::endl;
    }
    const double first = 1e-5, last = 1e2;
    const double step = pow(10, log10(last / first) / (num_mix - 1 - 1));
    vars.push_back(0);
    double v = first;
    for (int k = 0; k < num_mix - 1; k++, v *= step) vars.push_back(v);
    for (double& x : vars) x /= N;
    if (rank == 0) {
        std::cout << "scaled variances = ";
        for (double x : vars) std::cout << x * N << ' ';
        std::cout << std::endl;
    }
}

// ---- simulation helpers ------------------------------------------------------------------------------
// one draw from sum_j pi_j N(0, eta_j) with a generator seeded per call (the reference reseeds
// mt19937{seed} for every element, so element i of simulate() uses seed+i)
double generate_mixture_gaussians(int K_grp, const std::vector<double>& eta, const std::vector<double>& pi, long unsigned int seed) {
    std::mt19937 gen{seed};
    std::uniform_real_distribution<double> unif(0.0, 1.0);
    const double u = unif(gen);
    double cum = 0;
    for (int j = 0; j < K_grp; j++) {
        cum += pi[j];
        if (u <= cum) {
            if (eta[j] == 0) return 0;
            std::normal_distribution<double> gauss(0.0, sqrt(eta[j]));
            return gauss(gen);
        }
    }
    return 0;
}

// Element i is drawn from its own mt19937{seed + i} (utilities.cpp:77-88), so the elements are independent of each other: the loop is
// split over the host threads (N = 200k elements cost 0.5 s of generator seeding on one core, as much as a probit iteration at config 3).
std::vector<double> simulate(int M, std::vector<double> eta, std::vector<double> pi, long unsigned int seed) {
    std::vector<double> signal(M, 0.0);
    const int K = (int)eta.size();
    const int T = (int)std::max(1u, std::min(32u, std::min(std::thread::hardware_concurrency(), (unsigned)(M / 4096 + 1))));
    auto work = [&](int lo, int hi) {
        for (int i = lo; i < hi; i++) signal[i] = generate_mixture_gaussians(K, eta, pi, seed + i);
    };
    if (T <= 1) {
        work(0, M);
        return signal;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < T; t++) pool.emplace_back(work, (int)((long)M * t / T), (int)((long)M * (t + 1) / T));
    for (std::thread& th : pool) th.join();
    return signal;
}

double noise_prec_calc(double SNR, std::vector<double> vars, std::vector<double> probs, int Mt, int N) {
    (void)N;
    double expe = 0;
    for (size_t i = 0; i < vars.size(); i++) expe += vars[i] * probs[i];
    return SNR / Mt / expe;
}

// ---- vector files ------------------------------------------------------------------------------------
std::vector<double> read_vec_from_file(std::string filename, int M, int S) {
    std::vector<double> v;
    std::ifstream in(filename);
    double value;
    for (int it = 0; in >> value; it++) {
        if (it >= S + M) break;
        if (it >= S) v.push_back(value);
    }
    return v;
}

void store_vec_to_file(std::string filepath, std::vector<double> vec) {
    std::ofstream file(filepath);
    for (double x : vec) file << x << std::endl;
}

// raw doubles, this shard at byte offset S*8; the file is created but never truncated
void store_doubles_at(const std::string& filepath, const double* vec, int S, int M) {
    int fd = open(filepath.c_str(), O_CREAT | O_WRONLY, 0644);
    if (fd < 0) return;
    const char* p = reinterpret_cast<const char*>(vec);
    size_t total = (size_t)M * sizeof(double), done = 0;
    while (done < total) {
        ssize_t w = pwrite(fd, p + done, total - done, (off_t)S * (off_t)sizeof(double) + (off_t)done);
        if (w <= 0) break;
        done += (size_t)w;
    }
    close(fd);
}

void mpi_store_vec_to_file(std::string filepath, std::vector<double> vec, int S, int M) { store_doubles_at(filepath, vec.data(), S, M); }

std::vector<double> mpi_read_vec_from_file(std::string filename, int M, int S) {
    std::vector<double> vec(M, 0.0);
    int fd = open(filename.c_str(), O_RDONLY);
    if (fd < 0) return vec;
    char* p = reinterpret_cast<char*>(vec.data());
    size_t total = (size_t)M * sizeof(double), done = 0;
    while (done < total) {
        ssize_t r = pread(fd, p + done, total - done, (off_t)S * (off_t)sizeof(double) + (off_t)done);
        if (r <= 0) break;
        done += (size_t)r;
    }
    close(fd);
    return vec;
}

// ---- host reductions -----------------------------------------------------------------------------------
double inner_prod(std::vector<double> const& u, std::vector<double> const& v, int sync) {
    double acc = 0;
    for (size_t i = 0; i < u.size(); i++) acc += u[i] * v[i];
    if (sync == 1) gvb_host::allreduce_sum(&acc, 1);
    return acc;
}

double l2_norm2(std::vector<double> const& u, int sync) { return inner_prod(u, u, sync); }

double calc_stdev(std::vector<double> vec, int sync) {
    double s[3] = {std::accumulate(vec.begin(), vec.end(), 0.0), std::inner_product(vec.begin(), vec.end(), vec.begin(), 0.0),
                   (double)vec.size()};
    if (sync == 1) gvb_host::allreduce_sum(s, 3);
    double n = s[2], mean = s[0] / n;
    return std::sqrt((s[1] - n * mean * mean) / (n - 1));
}

// ---- partition -------------------------------------------------------------------------------------------
std::vector<double> divide_work(int Mt) {
    const gvb_host::Comm& w = gvb_host::world();
    long M = 0, S = 0;
    gvb_divide_work(Mt, w.nranks, w.rank, &M, &S);
    int Mm = Mt % w.nranks != 0 ? Mt / w.nranks + 1 : Mt / w.nranks;
    printf("INFO   : rank %4d has %ld markers over tot Mt = %d, max Mm = %d, starting at S = %ld\n", w.rank, M, Mt, Mm, S);
    return {(double)M, (double)S, (double)Mm};
}

// ---- special functions -----------------------------------------------------------------------------------
double normal_cdf(double value) { return 0.5 * erfc(-value * M_SQRT1_2); }
int sgn(double val) { return (0.0 < val) - (val < 0.0); }

// exp(x^2) erfc(x): minimax polynomial in q=(|x|-4)/(|x|+4) for (1+2|x|) erfcx(|x|), reflected for x<0.
// Same coefficients as the device version in csrc/vecops.cu (and the reference's utilities.cpp:345-409).
double erfcx(double x) {
    static const double c[] = {0x1.edcad78fc8044p-31, 0x1.b1548f14735d1p-30, -0x1.a1ad2e6c4a7a8p-27, -0x1.1985b48f08574p-26,
                               0x1.c6a8093ac4f83p-24, 0x1.31c2b2b44b731p-24, -0x1.b87373facb29fp-21, 0x1.3fef1358803b7p-22,
                               0x1.7eec072bb0be3p-18, -0x1.78a680a741c4ap-17, -0x1.9951f39295cf4p-16, 0x1.3be1255ce180bp-13,
                               -0x1.a1df71176b791p-13, -0x1.8d4aaa0099bc8p-11, 0x1.49c673066c831p-8, -0x1.0962386ea02b7p-6,
                               0x1.3079edf465cc3p-5, -0x1.0fb06dfedc4ccp-4, 0x1.7fee004e266dfp-4, -0x1.9ddb23c3e14d2p-4,
                               0x1.16ecefcfa4865p-4, 0x1.f7f5df66fc349p-7, -0x1.1df1ad154a27fp-3, 0x1.dd2c8b74febf6p-3};
    const double a = fmax(x, 0.0 - x);
    double r = 1.0 / (a + 4.0);
    double q = (a - 4.0) * r;
    double t = fma(q + 1.0, -4.0, a);
    q = fma(r, fma(q, -a, t), q);
    double p = c[0];
    for (size_t k = 1; k < sizeof(c) / sizeof(c[0]); k++) p = fma(p, q, c[k]);
    r = 0.5 / (a + 0.5);
    double qq = fma(p, r, r);
    double e = (p - qq) + fma(qq + qq, -a, 1.0);
    r = fma(e, r, qq);
    if (a > 0x1.fffffffffffffp1023) r = 0.0;
    if (x < 0.0) {
        double s = x * x, d = fma(x, x, -s), ex = exp(s);
        r = ex - r;
        r = fma(ex, d + d, r);
        r = r + ex;
        if (ex > 0x1.fffffffffffffp1023) r = ex;
    }
    return r;
}
