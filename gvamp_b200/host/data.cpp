// data.cpp -- see data.hpp.  Loader semantics follow the reference (data.cpp:128-331): PLINK .phen
// with "NA", phenotype scaled by 1/sd but not centred, 4-bit presence mask, covariates as a
// whitespace-separated N x C text matrix.  The .bed shard goes straight to HBM.
#include "data.hpp"

#include <immintrin.h>

#include <cassert>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <limits>
#include <sstream>

#include "comm.hpp"
#include "utilities.hpp"

namespace {
[[noreturn]] void device_fatal(const char* what) {
    std::cout << "FATAL: " << what << ": " << gvb_last_error() << std::endl;
    exit(EXIT_FAILURE);
}
gvb_ctx* g_root_ctx = nullptr;   // first context of the process owns the NCCL communicator
int g_ctx_users = 0;
}  // namespace

void data::open_device() {
    const gvb_host::Comm& w = gvb_host::world();
    if (!g_root_ctx) {
        if (gvb_ctx_create(&ctx, w.local_rank, w.rank, w.nranks, w.have_id ? w.nccl_id : nullptr) != GVB_OK) device_fatal("cannot open the CUDA device");
        gvb_host::rendezvous_done();
        g_root_ctx = ctx;
        gvb_host::set_collective_ctx(ctx);
    } else {
        if (gvb_ctx_create_shared(&ctx, g_root_ctx) != GVB_OK) device_fatal("cannot open a second device context");
    }
    owns_ctx = true;
    g_ctx_users++;
}

data::data(std::string fp, std::string genofp, const int N, const int M, const int Mt, const int S, const int rank, std::string type_data,
           double alpha_scale, std::string bimfp)
    : Mt(Mt), N(N), M(M), S(S), rank(rank), phenfp(fp), bimfp(bimfp), type_data(type_data), im4(N % 4 == 0 ? N / 4 : N / 4 + 1),
      mbytes((N % 4) ? (size_t)N / 4 + 1 : (size_t)N / 4), alpha_scale(alpha_scale) {
    if (type_data != "bed") {
        std::cout << "FATAL: only type_data == \"bed\" is supported by the B200 build" << std::endl;
        exit(EXIT_FAILURE);
    }
    mave = (double*)_mm_malloc(size_t(M) * sizeof(double), 32);
    msig = (double*)_mm_malloc(size_t(M) * sizeof(double), 32);
    bedfp = genofp;
    read_phen();
    read_genotype_data();
    compute_markers_statistics();
}

data::data(std::vector<double> y, std::string genofp, const int N, const int M, const int Mt, const int S, const int rank, std::string type_data,
           double alpha_scale, std::string bimfp)
    : Mt(Mt), N(N), M(M), S(S), rank(rank), bimfp(bimfp), type_data(type_data), phen_data(y), im4(N % 4 == 0 ? N / 4 : N / 4 + 1),
      mbytes((N % 4) ? (size_t)N / 4 + 1 : (size_t)N / 4), alpha_scale(alpha_scale) {
    if (type_data != "bed") {
        std::cout << "FATAL: only type_data == \"bed\" is supported by the B200 build" << std::endl;
        exit(EXIT_FAILURE);
    }
    mave = (double*)_mm_malloc(size_t(M) * sizeof(double), 32);
    msig = (double*)_mm_malloc(size_t(M) * sizeof(double), 32);
    // every individual carries a phenotype; only the pad bits of the last byte are cleared
    mask4.assign(mbytes, 0x0F);
    m4 = N % 4;
    if (m4 != 0) {
        for (int i = m4; i < 4; i++) mask4[N / 4] &= ~(0b1 << i);
        std::cout << "rank = " << rank << ": setting last " << 4 - m4 << " bits to NAs" << std::endl;
    }
    set_nonas(N);
    bedfp = genofp;
    read_genotype_data();
    compute_markers_statistics();
}

data::data(gvb_ctx* resident, std::vector<double> y, const int N, const int M, const int Mt, const int S, const int rank, double alpha_scale)
    : Mt(Mt), N(N), M(M), S(S), rank(rank), type_data("bed"), phen_data(y), im4(N % 4 == 0 ? N / 4 : N / 4 + 1),
      mbytes((N % 4) ? (size_t)N / 4 + 1 : (size_t)N / 4), alpha_scale(alpha_scale), ctx(resident), owns_ctx(false) {
    mave = (double*)_mm_malloc(size_t(M) * sizeof(double), 32);
    msig = (double*)_mm_malloc(size_t(M) * sizeof(double), 32);
    mask4.assign(mbytes, 0x0F);
    if (N % 4 != 0)
        for (int i = N % 4; i < 4; i++) mask4[N / 4] &= ~(0b1 << i);
    set_nonas(N);
    if (!gvb_host::collective_ctx()) gvb_host::set_collective_ctx(ctx);
    compute_markers_statistics();
}

data::~data() {
    for (gvb_vec& v : people)
        if (v && ctx) { gvb_vec_free(ctx, v); v = nullptr; }
    if (mave) _mm_free(mave);
    if (msig) _mm_free(msig);
    if (ctx && owns_ctx) {
        if (ctx == g_root_ctx) {
            // keep the root context (and its communicator) alive while other data objects use it
            if (--g_ctx_users > 0) return;
            gvb_host::set_collective_ctx(nullptr);
            g_root_ctx = nullptr;
        } else {
            --g_ctx_users;
        }
        gvb_ctx_destroy(ctx);
    }
}

// PLINK phenotype file: FID IID value, one line per individual, "NA" for missing
void data::read_phen() {
    std::ifstream infile(phenfp);
    if (!infile.is_open()) {
        std::cout << "FATAL: could not open phenotype file: " << phenfp << std::endl;
        exit(EXIT_FAILURE);
    }
    const double NA_MARK = std::numeric_limits<double>::max();
    std::string line;
    int line_n = 0;
    double sum = 0.0;
    nonas = 0;
    nas = 0;
    while (getline(infile, line)) {
        if (line_n % 4 == 0) mask4.push_back(0x0F);
        std::istringstream fields(line);
        std::string fid, iid, value;
        fields >> fid >> iid >> value;
        if (value == "NA") {
            nas++;
            phen_data.push_back(NA_MARK);
            mask4[line_n / 4] &= ~(0b1 << (line_n % 4));
        } else {
            nonas++;
            double v = atof(value.c_str());
            phen_data.push_back(v);
            sum += v;
        }
        line_n++;
    }
    assert(nas + nonas == N);
    if (line_n % 4 != 0) {
        for (int i = line_n % 4; i < 4; i++) mask4[line_n / 4] &= ~(0b1 << i);
        std::cout << "rank = " << rank << ": setting last " << 4 - line_n % 4 << " bits to NAs" << std::endl;
    }
    // scale by the inverse standard deviation of the observed values (no centring)
    const double avg = sum / double(nonas);
    double sq = 0.0;
    for (double v : phen_data)
        if (v != NA_MARK) sq += (v - avg) * (v - avg);
    const double inv_sd = sqrt(double(nonas - 1) / sq);
    for (double& v : phen_data) v *= inv_sd;
    intercept = avg;
    scale = inv_sd;
}

void data::upload_mask() {
    if (gvb_set_mask(ctx, mask4.data(), nonas) != GVB_OK) device_fatal("cannot upload the phenotype mask");
}

// the shard [S, S+M) of the .bed file goes to HBM (file offset 3 + S*mbytes)
void data::read_genotype_data() {
    double ts = gvb_host::wtime();
    open_device();
    size_t size_bytes = size_t(M) * mbytes;
    printf("INFO   : rank %d has allocated %zu bytes (%.3f GB) of HBM for raw data.\n", rank, size_bytes, double(size_bytes) / 1.0E9);
    int rc = gvb_bed_load_file(ctx, bedfp.c_str(), N, Mt, S, M);
    if (rc != GVB_OK) device_fatal("reading genotype data failed");
    upload_mask();
    double te = gvb_host::wtime();
    if (rank == 0) std::cout << "reading genotype data took " << te - ts << " seconds." << std::endl;
}

void data::read_covariates(std::string covfp, int C) {
    if (C == 0) return;
    double start = gvb_host::wtime();
    std::ifstream covf(covfp);
    std::string line;
    while (std::getline(covf, line)) {
        std::vector<double> entries;
        std::istringstream fields(line);
        std::string tok;
        while (fields >> tok) entries.push_back(std::stod(tok));
        if ((int)entries.size() != C) {
            std::cout << "FATAL: number of covariates = " << entries.size() << " does not match to the specified number of covariates = " << C
                      << std::endl;
            exit(EXIT_FAILURE);
        }
        covs.push_back(entries);
    }
    if (rank == 0) std::cout << "rank = " << rank << ": reading covariates took " << gvb_host::wtime() - start << " seconds to run." << std::endl;
}

void data::compute_markers_statistics() {
    double start = gvb_host::wtime();
    upload_mask();   // the mask (and nonas) may have been changed since the matrix was loaded
    if (gvb_compute_stats(ctx, alpha_scale) != GVB_OK) device_fatal("marker statistics failed");
    if (gvb_get_stats(ctx, mave, msig) != GVB_OK) device_fatal("marker statistics download failed");
    if (rank == 0) std::cout << "rank = " << rank << ": statistics took " << gvb_host::wtime() - start << " seconds to run." << std::endl;
}

// per-individual statistics over all markers of all shards (data.cpp:548-640): three X.v-type sweeps on the device
void data::compute_people_statistics() {
    double start = gvb_host::wtime();
    for (int k = 0; k < 3; k++)
        if (!people[k] && gvb_vec_alloc_N(ctx, &people[k]) != GVB_OK) device_fatal("device vector allocation failed");
    if (gvb_people_stats(ctx, people[0], people[1], people[2]) != GVB_OK) device_fatal("people statistics failed");
    mave_people.assign(4 * mbytes, 0.0);
    msig_people.assign(4 * mbytes, 0.0);
    numb_people.assign(4 * mbytes, 0.0);
    if (gvb_vec_download(ctx, people[0], mave_people.data(), 4 * (long)mbytes) != GVB_OK ||
        gvb_vec_download(ctx, people[1], msig_people.data(), 4 * (long)mbytes) != GVB_OK ||
        gvb_vec_download(ctx, people[2], numb_people.data(), 4 * (long)mbytes) != GVB_OK)
        device_fatal("people statistics download failed");
    if (rank == 0) std::cout << "rank = " << rank << ": people statistics took " << gvb_host::wtime() - start << " seconds to run." << std::endl;
}

unsigned char* data::get_bed_data() {
    if (bed_host.empty()) {
        bed_host.resize(size_t(M) * mbytes);
        if (gvb_bed_decode(ctx, 0, M, bed_host.data()) != GVB_OK) device_fatal("decoding the genotype matrix failed");
    }
    return bed_host.data();
}

// ---- design-matrix products (host pointers in, std::vector out, like the reference) -----------------
std::vector<double> data::Ax(double* __restrict__ phen) { return Ax(phen, 0, (int)mbytes); }

std::vector<double> data::Ax(double* __restrict__ phen, int SB, int LB) {
    std::vector<double> out(4 * (size_t)LB, 0.0);
    if (gvb_Ax(ctx, phen, out.data(), SB, LB) != GVB_OK) device_fatal("Ax failed");
    return out;
}

std::vector<double> data::ATx(double* __restrict__ phen) { return ATx(phen, 0, (int)mbytes); }

std::vector<double> data::ATx(double* __restrict__ phen, int SB, int LB) {
    std::vector<double> out(M, 0.0);
    // the reference reads 4*LB entries; callers that pass an N-long vector (vamp passes y) rely on the pad entries
    // not mattering, so only the entries that address real individuals are taken
    if (gvb_ATx(ctx, phen, out.data(), SB, LB) != GVB_OK) device_fatal("ATx failed");
    return out;
}

double data::dot_product(const int mloc, double* __restrict__ phen, const double mu, const double sigma_inv, const int SB, const int LB) {
    double dpa = 0, dpb = 0;
    if (gvb_marker_dot(ctx, mloc, phen, SB, LB, &dpa, &dpb) != GVB_OK) device_fatal("dot_product failed");
    return sigma_inv * (dpa - mu * dpb);
}

std::vector<double> data::Zx(std::vector<double> phen) {
    std::vector<double> out(4 * mbytes, 0.0);
    for (int i = 0; i < N; i++) out[i] = inner_prod(covs[i], phen, 0);
    return out;
}

std::vector<double> data::filter_pheno() {
    std::vector<double> y = get_phen();
    for (int i = 0; i < N && i < (int)y.size(); i++)
        if (((mask4[i / 4] >> (i % 4)) & 1) == 0) y[i] = 0;
    return y;
}

std::vector<double> data::filter_pheno(int* nonnan) {
    *nonnan = 0;
    std::vector<double> y = get_phen();
    for (int i = 0; i < N && i < (int)y.size(); i++) {
        if (((mask4[i / 4] >> (i % 4)) & 1) == 0)
            y[i] = 0;
        else
            (*nonnan)++;
    }
    return y;
}

// ---- association tests ------------------------------------------------------------------------------
// chromosome index of every local marker from column 1 of the .bim file, "X" -> 23 (data.cpp:346-380)
std::vector<int> data::read_chromosome_info(std::string bim_file) {
    std::vector<int> chroms;
    std::ifstream infile(bim_file);
    if (!infile.is_open()) {
        std::cout << "FATAL: could not open bim file: " << bim_file << std::endl;
        exit(EXIT_FAILURE);
    }
    std::string line;
    for (int line_n = 0; std::getline(infile, line); line_n++) {
        if (line_n < S || line_n >= S + M) continue;
        std::istringstream fields(line);
        std::string tok;
        fields >> tok;
        chroms.push_back(tok == "X" ? 23 : (int)atof(tok.c_str()));
    }
    return chroms;
}

namespace {
struct DevVec {   // scoped device vector
    gvb_ctx* ctx;
    gvb_vec v = nullptr;
    DevVec(gvb_ctx* c, bool n_vector) : ctx(c) {
        if ((n_vector ? gvb_vec_alloc_N(c, &v) : gvb_vec_alloc_M(c, &v)) != GVB_OK) device_fatal("device vector allocation failed");
    }
    ~DevVec() { gvb_vec_free(ctx, v); }
};
}  // namespace

// data::pvals_calc, data.cpp:1108-1180: leave-one-out t-test of every marker, y - z1 + x_k x1_hat_k regressed on x_k.
// One call of the device association pass (two bed sweeps) per estimator.
std::vector<std::vector<double>> data::pvals_calc(std::vector<std::vector<double>> z1, std::vector<double> y, std::vector<std::vector<double>> x1_hat,
                                                  std::vector<std::string> filepath) {
    const int nE = (int)z1.size();
    std::vector<std::vector<double>> pvals(nE, std::vector<double>(M, 0.0));
    DevVec yres(ctx, true), coef(ctx, false), pv(ctx, false);
    for (int ie = 0; ie < nE; ie++) {
        std::vector<double> ymod(4 * mbytes, 0.0);
        for (int i = 0; i < N; i++) ymod[i] = y[i] - z1[ie][i];
        if (gvb_vec_upload(ctx, yres.v, ymod.data(), (long)ymod.size()) != GVB_OK || gvb_vec_upload(ctx, coef.v, x1_hat[ie].data(), M) != GVB_OK ||
            gvb_assoc_pvals(ctx, yres.v, coef.v, nullptr, pv.v) != GVB_OK || gvb_vec_download(ctx, pv.v, pvals[ie].data(), M) != GVB_OK)
            device_fatal("LOO p-values failed");
        if (rank == 0) std::cout << "so far worker 0 has calculated " << M << " pvals." << std::endl;
        if (ie < (int)filepath.size()) mpi_store_vec_to_file(filepath[ie], pvals[ie], S, M);
    }
    return pvals;
}

// data::pvals_calc_LOCO, data.cpp:1220-1353: per chromosome, the predictor of the chromosome's own markers (all shards,
// all-reduced inside Ax) is added back to y - z1 and the chromosome's markers are tested against it.
std::vector<std::vector<double>> data::pvals_calc_LOCO(std::vector<std::vector<double>> z1, std::vector<double> y,
                                                       std::vector<std::vector<double>> x1_hat, std::vector<std::string> filepath) {
    const int nE = (int)z1.size();
    std::vector<std::vector<double>> pvals(nE, std::vector<double>(M, 0.0));
    std::vector<int> ch_info = read_chromosome_info(bimfp);
    ch_info.resize(M, 0);
    DevVec yres(ctx, true), sel(ctx, false), pv(ctx, false);
    for (int ie = 0; ie < nE; ie++) {
        if (gvb_vec_fill(ctx, pv.v, 0.0) != GVB_OK) device_fatal("LOCO p-values failed");
        for (int ch = 1; ch <= 23; ch++) {
            std::vector<double> xch(M, 0.0), select(M, 0.0);
            for (int m = 0; m < M; m++)
                if (ch_info[m] == ch) { xch[m] = x1_hat[ie][m]; select[m] = 1.0; }
            std::vector<double> y_chrom = Ax(xch.data());   // collective: every rank walks all 23 chromosomes
            for (size_t i = N; i < y_chrom.size(); i++) y_chrom[i] = 0.0;
            std::string filepath_predictors = filepath[ie] + "_LOCO_chr_" + std::to_string(ch) + ".csv";
            if (rank == 0) {
                store_vec_to_file(filepath_predictors, y_chrom);
                std::cout << "filepath predictors = " << filepath_predictors << std::endl;
            }
            for (int i = 0; i < N; i++) y_chrom[i] += y[i] - z1[ie][i];
            if (gvb_vec_upload(ctx, yres.v, y_chrom.data(), (long)y_chrom.size()) != GVB_OK || gvb_vec_upload(ctx, sel.v, select.data(), M) != GVB_OK ||
                gvb_assoc_pvals(ctx, yres.v, nullptr, sel.v, pv.v) != GVB_OK)
                device_fatal("LOCO p-values failed");
        }
        if (gvb_vec_download(ctx, pv.v, pvals[ie].data(), M) != GVB_OK) device_fatal("LOCO p-values failed");
        mpi_store_vec_to_file(filepath[ie] + "_pvals_LOCO.bin", pvals[ie], S, M);
    }
    return pvals;
}
