// gvb_internal.cuh -- shared declarations of libgvamp_b200 (not part of the public C ABI).
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/gvamp_b200.h"

// ------------------------------------------------------------------------------------------------
// HBM layout of the packed genotype matrix ("striped, 4-marker interleaved SNP-major")
//
//   position p   = byte index inside a marker column (4 individuals), p in [0, mbytes)
//   stripe  t    = p / 32            (128 individuals, one 128-byte line per marker group)
//   group   g    = local marker / 4  (4 consecutive markers)
//   word(t,g,l)  = bed32[(t * Mg_pad + g) * 32 + l],  byte q of the word = PLINK byte of marker
//                  4g+q at position 32t+l.
//
// A stripe is one contiguous run of Mg_pad*128 bytes, so X^T.u (marker-stationary) and X.v
// (individual-stationary) both stream long contiguous runs, every PLINK byte is stored exactly once
// and unchanged, and a 32-bit word holds a 4x4 (marker x individual) block: byte q indexes a
// per-position table for X^T.u, and the four codes of one individual are gathered with one
// multiply for X.v (see matvec_lut.cu).  Padding bytes are 0x55 (four "missing" codes): they
// contribute nothing to any product or count.
// ------------------------------------------------------------------------------------------------
#define GVB_STRIPE_POS 32         // byte positions per stripe
#define GVB_STRIPE_IND 128        // individuals per stripe
#define GVB_GROUP 4               // markers per interleaved word
#define GVB_GROUP_TILE 32         // marker groups per table tile (Mg_pad is a multiple of this)
#define GVB_PAD_BYTE 0x55u
#define GVB_SNAP_SLOTS 16

struct gvb_vec_s {
    double* d;
    long n;      // logical length
    long cap;    // allocated doubles
};

struct gvb_ctx {
    int device = 0, rank = 0, nranks = 1;
    cudaStream_t stream = nullptr;
    ncclComm_t comm = nullptr;
    bool owns_comm = true;
    int sm_count = 148;

    // problem dimensions
    long N = 0, Mt = 0, S = 0, M = 0, mbytes = 0;
    long n_stripes = 0, Npad = 0;   // Npad = n_stripes * 128
    long Mg = 0, Mg_pad = 0;        // marker groups, padded to GVB_GROUP_TILE
    size_t bed_words = 0;
    uint32_t* bed = nullptr;        // the HBM-resident matrix
    // individual-major twin of the matrix (twin.cu): same indexing, every word 4x4-transposed, so that byte k of a word is
    // the ready-made table index of individual k for X.v; held only while spare HBM allows it
    uint32_t* bed_twin = nullptr;
    int twin_state = 0;             // 0: not built yet, 1: built, 2: built for the first twin_stripes stripes only, -1: not available (X.v gathers from `bed`)
    long twin_stripes = 0;          // stripes [0, twin_stripes) have a twin (== n_stripes when twin_state == 1)

    // phenotype mask / valid-individual mask, one 32-bit word per position: bits 0,2,4,6 of every
    // byte are set when individual k of that position is present (resp. < N)
    uint32_t* maskw = nullptr;
    uint32_t* validw = nullptr;
    int nonas = 0;
    long mask_present = 0;          // individuals the mask selects (== nonas when the caller is consistent)
    bool have_mask = false, have_stats = false;

    // marker statistics
    double* mave = nullptr;   // Mg_pad*4 (padded markers: 0)
    double* msig = nullptr;   // padded markers: 0 (so they contribute nothing)
    int64_t* counts = nullptr;
    long total_missing = 0;   // sum over local markers of missing genotypes (i < N)
    // sparse list of the missing genotypes (misslist.cu): the X^T.u correction sum_{i missing in j} U_i as a gather
    uint16_t* miss_idx = nullptr;             // 16-bit indices into a 32768-individual block, segments padded to groups of 4
    unsigned long long* miss_off = nullptr;   // [miss_nblk * Mg_pad*4 + 1] segment offsets in groups, block-major / marker-minor
    int* uq = nullptr;                        // quantised u of the running X^T.u sweep (Npad)
    long miss_nblk = 0, miss_entries = 0;
    int miss_state = 0;                       // 0: not built yet, 1: built, -1: not available (X^T.u walks the bed a second time)
    double alpha_scale = 1.0;

    // scratch
    double* red_partial = nullptr;  // [RED_BLOCKS][RED_MAXK]
    double* red_result = nullptr;   // [RED_MAXK] device
    double* h_red = nullptr;        // pinned host mirror
    double* ax_partial = nullptr;   // Ax v0 partial sums
    size_t ax_partial_cap = 0;
    double* tmpN = nullptr;         // Npad
    double* tmpN2 = nullptr;
    double* tmpM = nullptr;         // Mg_pad*4
    double* tmpM2 = nullptr;
    double* wv = nullptr;           // sigma_j * v_j            (Mg_pad*4)
    double* cv = nullptr;           // mu_j * sigma_j * v_j
    // LUT-kernel scratch (matvec_lut.cu)
    int32_t* tab_u = nullptr;       // per-position tables for X^T.u
    size_t tab_u_cap = 0;
    int32_t* tab_v = nullptr;       // per-marker-group tables for X.v
    size_t tab_v_cap = 0;
    int* shift_u = nullptr;         // per stripe: left shift of its table's int32 windows (scale classes, matvec_tile.cu)
    int* shift_v = nullptr;         // per marker tile, likewise
    int32_t* tab_v2 = nullptr;      // dual X.v (two right-hand sides per bed read): interleaved tables, the second product's shifts / accumulators
    size_t tab_v2_cap = 0;
    int* shift_v2 = nullptr;
    unsigned long long* acc_dual = nullptr;
    int32_t* tab_u2 = nullptr;      // dual X^T.u: interleaved per-stripe tables and the second product's shifts
    size_t tab_u2_cap = 0;
    int* shift_u2 = nullptr;
    size_t acc_dual_cap = 0;
    size_t shift_cap = 0;
    unsigned long long* acc_i64 = nullptr;  // fixed-point accumulators
    size_t acc_i64_cap = 0;
    int* work_counter = nullptr;
    // device-side predicate of the sweep kernels (matvec_tile.cu, misslist.cu): when non-null and *skip != 0 every kernel of a
    // sweep returns at once.  Set by the CG driver (cg.cu) around iterations it enqueues before it knows the solver has stopped.
    const int* skip = nullptr;
    double* scal = nullptr;         // small device scalars
    // layout of the lookup tables in global memory (matvec_tile.cu): 0 = one 32 KB tile per step, [step][256 entries][32 slots]; 1 = the
    // tiles of two consecutive steps interleaved, [pair][256][2][32], so that one TMA bulk copy fills a 64 KB table region ("pair" mode)
    int tab_pairs = 0;
    int kernel_gen = 2;             // 0: simple FP64 kernels, 1: gen-1 table kernels, 2: gen-2 tile kernels (env GVB_KERNELS)

    // asynchronous device -> host snapshots of vectors (capi.cu): the iteration outputs leave over a copy stream while
    // the sweeps go on; staging in HBM decouples them from later updates of the vector
    struct Snap {
        double* dev = nullptr;    // staging copy in HBM
        double* host = nullptr;   // pinned
        long cap = 0, n = 0;
        cudaEvent_t ready = nullptr, done = nullptr;
        bool pending = false;
    } snap[GVB_SNAP_SLOTS];
    cudaStream_t copy_stream = nullptr;

    // timers and counters
    cudaEvent_t ev_start[8], ev_stop[8];
    long launches = 0, sweeps = 0, dual_sweeps = 0;   // a dual sweep (two products, one bed read) counts once in `sweeps`
    long host_syncs = 0;            // stream synchronisations that return a reduction to the host (bench.py reports them per iteration)
    bool profile = false;
    std::vector<cudaEvent_t> prof_ev[3];   // [0] X.v, [1] X^T.u, [2] dual X.v : start/stop pairs
    size_t prof_used[3] = {0, 0, 0};
    std::vector<gvb_vec_s*> vecs;
    size_t cg_ap_cap = 0;
    double* cg_ap = nullptr;        // A p0 of a prepared solve (gvb_cg_prepare), consumed by iteration 0 of gvb_cg_solve_prepared
    bool cg_prepared = false;
    gvb_cg_companion_fn cg_companion = nullptr;   // gvb_cg_set_companion: consumed by the next solve
    void* cg_companion_user = nullptr;
    gvb_vec_s* cg_ws[3] = {nullptr, nullptr, nullptr};   // r, p, d of the CG solver (allocated once: no cudaMalloc in the loop)
    // device-resident scalars of the CG solver (cg.cu): the iteration never returns to the host for alpha / beta / the exit tests
    double* cg_dev = nullptr;       // [GVB_CG_NSCAL scalars][4 log doubles per iteration]
    double* cg_host = nullptr;      // pinned: 4 ring slots of GVB_CG_NSCAL scalars, then the final copy of cg_dev
    int cg_log_cap = 0;             // iterations the log has room for
    int* cg_flags = nullptr;        // device: [0] != 0 once the solver has stopped (the predicate of every later kernel)
    cudaEvent_t cg_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    long layout_gen = 0;            // bumped whenever the matrix or the mask is (re)loaded: callers drop cached by-products
};

#define GVB_CG_NSCAL 16
#define GVB_RED_BLOCKS 1184   // 8 blocks per SM: the FP64 exp-heavy denoiser / EM kernels are latency-bound and want the warps (0.6 -> 0.2 ms at 1.6M markers)
#define GVB_RED_MAXK 80

void gvb_set_error(const char* fmt, ...);

#define GVB_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            gvb_set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #call); \
            return GVB_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define GVB_NCCL(call)                                                                         \
    do {                                                                                       \
        ncclResult_t r_ = (call);                                                              \
        if (r_ != ncclSuccess) {                                                               \
            gvb_set_error("NCCL error %s at %s:%d (%s)", ncclGetErrorString(r_), __FILE__, __LINE__, #call); \
            return GVB_ERR_NCCL;                                                               \
        }                                                                                      \
    } while (0)

#define GVB_CHECK(call)                 \
    do {                                \
        int rc_ = (call);               \
        if (rc_ != GVB_OK) return rc_;  \
    } while (0)

#define GVB_ARG(cond, msg)                                     \
    do {                                                       \
        if (!(cond)) {                                         \
            gvb_set_error("invalid argument: %s (%s:%d)", msg, __FILE__, __LINE__); \
            return GVB_ERR_ARG;                                \
        }                                                      \
    } while (0)

// count + launch-error check after every kernel launch
#define GVB_LAUNCHED(ctx)                     \
    do {                                      \
        (ctx)->launches++;                    \
        GVB_CUDA(cudaGetLastError());         \
    } while (0)

static inline long gvb_roundup(long x, long m) { return (x + m - 1) / m * m; }

// index of table entry (step, entry e, slot s) in the global table array, see gvb_ctx::tab_pairs
__host__ __device__ static inline size_t gvb_tab_index(long step, int e, int s, int pairs) {
    return pairs ? (((size_t)(step >> 1) * 256 + e) * 2 + (step & 1)) * 32 + s : ((size_t)step * 256 + e) * 32 + s;
}

// Frees every individual-major twin held on `device` (the one optional, re-creatable consumer of HBM: X.v then gathers from the one
// matrix, bit-identical results); returns the number of twins released.  capi.cu keeps the registry of live contexts.
int gvb_release_twins(int device);

// cudaMalloc that gives the twins back before it gives up: a later allocation (a second matrix in run mode `both`, the scratch of the
// association tests, solver vectors) must not fail because spare HBM was spent on the optional second orientation.
template <typename T>
static inline cudaError_t gvb_malloc(gvb_ctx* c, T** p, size_t bytes) {
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation && gvb_release_twins(c->device) > 0) {
        cudaGetLastError();
        e = cudaMalloc(p, bytes);
    }
    return e;
}

// ---- internal entry points (implemented across the .cu files) ----
int gvb_layout_alloc(gvb_ctx* c, long N, long Mt, long S, long M);
int gvb_layout_retile(gvb_ctx* c, const uint8_t* d_snp_major, long j0, long nmark);   // device SNP-major chunk -> layout
int gvb_stats_run(gvb_ctx* c);

// full-range device matvecs; u/out N-vectors have Npad entries, v/out M-vectors have Mg_pad*4
int gvb_ax_simple(gvb_ctx* c, const double* v, double* out);
int gvb_atx_simple(gvb_ctx* c, const double* u, double* out);
int gvb_ax_lut(gvb_ctx* c, const double* v, double* out);
int gvb_atx_lut(gvb_ctx* c, const double* u, double* out);
int gvb_ax_tile(gvb_ctx* c, const double* v, double* out, int mode = 0);     // gen-2 sweeps (matvec_tile.cu); mode: ax_code_values
int gvb_ax2_dev(gvb_ctx* c, const double* v0, const double* v1, double* out0, double* out1);   // dual sweep + all-reduce (capi.cu)
int gvb_atx2_dev(gvb_ctx* c, const double* u0, const double* u1, double* out0, double* out1);   // dual X^T.u (capi.cu)
int gvb_atx_tile_dual(gvb_ctx* c, const double* u0, const double* u1, double* out0, double* out1);
int gvb_ax_tile_dual(gvb_ctx* c, const double* v0, const double* v1, double* out0, double* out1);   // two products, one bed read
int gvb_atx_tile(gvb_ctx* c, const double* u, double* out, double* outB = nullptr);   // outB[j] = sum_i b_ij u_i (optional)
int gvb_count_tile_main(gvb_ctx* c, const int* tab, unsigned long long* acc);
void gvb_twin_reset(gvb_ctx* c);                                       // twin.cu
int gvb_twin_build(gvb_ctx* c);
void gvb_misslist_reset(gvb_ctx* c);                                   // misslist.cu
int gvb_misslist_build(gvb_ctx* c);
int gvb_misslist_sum(gvb_ctx* c, unsigned long long* accm);
int gvb_ax_dev(gvb_ctx* c, const double* v, double* out, bool allreduce);
int gvb_atx_dev(gvb_ctx* c, const double* u, double* out);

// reductions: res (host) <- sum over blocks (and ranks if sync) of the K per-block partials written
// by the caller's kernel into c->red_partial[b*K + k], b < nblocks
int gvb_reduce_finish(gvb_ctx* c, int nblocks, int K, bool sync, double* res_host);
int gvb_reduce_device(gvb_ctx* c, int nblocks, int K, bool sync);                          // the same, result left in c->red_result
int gvb_vec_dots_device(gvb_ctx* c, int n, const gvb_vec* x, const gvb_vec* y, bool sync); // gvb_vec_dots without the host round trip
