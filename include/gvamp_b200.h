/* gvamp_b200.h -- C ABI of the B200-native gVAMP hot path (libgvamp_b200.so).
 *
 * One context (gvb_ctx) = one B200 = one contiguous marker shard [S, S+M) of an N x Mt genotype
 * matrix, i.e. exactly what one MPI rank owns in the reference (utilities.cpp:259-291).  All
 * entry points take plain pointers and sizes, return 0 on success and a non-zero code on error
 * (text via gvb_last_error()); no C++ exception crosses this boundary.  There is no CPU fallback:
 * every call fails with GVB_ERR_CUDA when no sm_100 device is usable.
 *
 * The reference has no FFI layer; the functions below replace the C++ member functions of
 * `class data` / `class vamp` that make up the hot path.  Each declaration cites the reference
 * interface it stands in for (file:line relative to the gVAMP source tree).
 *
 * Threading and devices: like a rank of the reference, a context is driven by ONE host thread, and a process normally holds
 * the context(s) of one GPU (gvb_ctx_create makes its device current and the later calls expect it to stay current).
 *
 * Vector conventions: "host" pointers are ordinary host memory (pinned or not); `gvb_vec` handles
 * are device-resident FP64 vectors owned by the context.  M-vectors are sharded like the markers,
 * N-vectors (length 4*ceil(N/4), padded entries are zero) are replicated on every rank.
 */
#ifndef GVAMP_B200_H
#define GVAMP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gvb_ctx gvb_ctx;
typedef struct gvb_vec_s* gvb_vec; /* device vector handle */

enum {
    GVB_OK = 0,
    GVB_ERR_CUDA = 1,     /* CUDA runtime / driver / no device */
    GVB_ERR_NCCL = 2,
    GVB_ERR_ARG = 3,      /* bad argument or call order */
    GVB_ERR_IO = 4,       /* file could not be read */
    GVB_ERR_NOMEM = 5
};

#define GVB_NCCL_ID_BYTES 128
#define GVB_MAX_MIX 32 /* max mixture components L (the reference's default prior has 23) */

/* ---- library / context -------------------------------------------------------------------- */
const char* gvb_last_error(void);
const char* gvb_version(void);
int gvb_device_count(int* n);

/* rank 0 calls this and ships the 128 bytes to the other ranks (torch.distributed broadcast, a
 * file, ...); replaces MPI_Init's implicit communicator (main_real.cpp:17-27). */
int gvb_nccl_unique_id(void* id128);

/* device: CUDA ordinal; rank/nranks: position of this shard; id128 may be NULL when nranks==1. */
int gvb_ctx_create(gvb_ctx** out, int device, int rank, int nranks, const void* id128);
/* a second context (e.g. the test-set matrix of --run-mode both) on the same device that shares the
 * parent's NCCL communicator; destroy it before the parent. */
int gvb_ctx_create_shared(gvb_ctx** out, gvb_ctx* parent);
void gvb_ctx_destroy(gvb_ctx* ctx);
int gvb_ctx_sync(gvb_ctx* ctx);                    /* cudaStreamSynchronize on the context's stream */
void* gvb_ctx_stream(gvb_ctx* ctx);                /* the cudaStream_t every kernel is launched on */
int gvb_ctx_info(gvb_ctx* ctx, long* N, long* Mt, long* S, long* M, long* mbytes);

/* CUDA-event timers on the context's stream (bench.py times with these; torch.cuda.Event only
 * sees torch's own stream). */
int gvb_timer_start(gvb_ctx* ctx, int slot);
int gvb_timer_stop(gvb_ctx* ctx, int slot);
int gvb_timer_elapsed_ms(gvb_ctx* ctx, int slot, float* ms); /* synchronises on the stop event */
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
long gvb_launch_count(gvb_ctx* ctx);
/* bed sweeps (Ax + ATx passes over the packed matrix) since creation */
long gvb_sweep_count(gvb_ctx* ctx);
/* stream synchronisations that returned a reduction (or a solver's exit flag) to the host since the context was created */
long gvb_host_sync_count(gvb_ctx* ctx);
/* bumped by every (re)load of the matrix and every gvb_set_mask: cached by-products of earlier solves are stale after a change */
long gvb_layout_generation(gvb_ctx* ctx);
/* per-sweep device timing: when enabled every X.v / X^T.u sweep is bracketed by CUDA events on the
 * context's stream.  gvb_profile_read synchronises, returns out[0..3] = {X.v total ms, X.v sweeps,
 * X^T.u total ms, X^T.u sweeps} since the last read and resets the counters. */
int gvb_profile_enable(gvb_ctx* ctx, int on);
int gvb_profile_read(gvb_ctx* ctx, double* out4);
/* the same for the dual sweeps (gvb_dAx2): out[0..1] = {total ms, sweeps} since the last read */
int gvb_profile_read_dual(gvb_ctx* ctx, double* out2);
long gvb_dual_sweep_count(gvb_ctx* ctx); /* dual sweeps so far; each also counts once in gvb_sweep_count */

/* ---- partition ------------------------------------------------------------------------------ */
/* divide_work, utilities.cpp:259-291: first Mt%nranks ranks own Mt/nranks+1 contiguous markers */
void gvb_divide_work(long Mt, int nranks, int rank, long* M, long* S);

/* ---- genotype matrix ------------------------------------------------------------------------- */
/* data::read_genotype_data, data.cpp:201-234: reads the shard's bytes at file offset 3+S*mbytes
 * (magic bytes skipped, never validated -- same as the reference) and re-tiles them into the
 * HBM-resident layout. */
int gvb_bed_load_file(gvb_ctx* ctx, const char* bed_path, long N, long Mt, long S, long M);
/* same, from a host buffer of M*ceil(N/4) SNP-major PLINK bytes (data::get_bed_data, data.hpp:66) */
int gvb_bed_load_host(gvb_ctx* ctx, const uint8_t* bed, long N, long Mt, long S, long M);
/* synthetic Binomial(2,p_j) genotypes generated directly in HBM (bench workloads up to 840 GB that
 * never touch disk); byte-identical to oracle/gvamp_oracle.c:orc_synth_bed. */
int gvb_bed_synth(gvb_ctx* ctx, uint64_t seed, long N, long Mt, long S, long M, double miss_rate);
/* PLINK bytes of local markers [j0, j0+n) back to the host: the bit-exactness check of the layout */
int gvb_bed_decode(gvb_ctx* ctx, long j0, long n, uint8_t* out);

/* phenotype-present mask: data::mask4 / nonas, data.cpp:136-169 and :86-100.  mask4==NULL means
 * "all N individuals present" (the y-vector constructor). */
int gvb_set_mask(gvb_ctx* ctx, const uint8_t* mask4, int nonas);

/* data::compute_markers_statistics, data.cpp:392-485 (scalar-branch semantics: mean 0 for an empty
 * column, inverse sd 1 for a constant one, alpha_scale exponent). */
int gvb_compute_stats(gvb_ctx* ctx, double alpha_scale);
int gvb_get_stats(gvb_ctx* ctx, double* mave, double* msig);   /* data::get_mave / get_msig */
/* counts[j*8+c]: individuals with code c and phenotype present; counts[j*8+4+c]: all i<N */
int gvb_get_counts(gvb_ctx* ctx, int64_t* counts);

/* entries (16-bit indices, padding included) of the sparse list of missing genotypes that X^T.u uses for the na_lut
 * term of data::dot_product (data.cpp:766) on shards with missing genotypes; 0 when the shard has none, the list is
 * not built yet (it is built by the first X^T.u after gvb_compute_stats) or X^T.u walks the bed a second time instead
 * (list larger than half the bed / out of memory / env GVB_MISS=twopass). */
long gvb_missing_list_entries(gvb_ctx* ctx);

/* state of the individual-major twin of the matrix that X.v (data::Ax, data.cpp:848-1011) walks when spare HBM holds one:
 * 1 built, 2 built for a prefix of the stripes only (as many as spare HBM holds; the other stripes gather from the one matrix),
 * 0 not decided yet (the first X.v after gvb_compute_stats decides), -1 not held (no spare HBM, or env GVB_TWIN=0): X.v then
 * gathers its table indices from the one matrix.  Results are bit-identical in every state. */
int gvb_twin_state(gvb_ctx* ctx);
long gvb_twin_stripes(gvb_ctx* ctx); /* stripes (of 128 individuals) that have a twin */
/* gives the twin's HBM back (state becomes -1: it is not rebuilt).  The library does this by itself for every context of the device
 * when one of its own allocations fails (a second matrix in run mode `both`, main_real.cpp:255-330, scratch of later stages). */
int gvb_twin_release(gvb_ctx* ctx);

/* ---- X.v and X^T.u, host-pointer drop-in --------------------------------------------------------- */
/* Precision contract of both products (and of everything built on them).  The sweeps run in exact fixed point after ONE
 * quantisation of the input vector; every stripe of 128 individuals (X^T.u) / tile of 128 markers (X.v) is quantised with its own
 * power-of-two scale class, so an outlier costs only its own stripe / tile resolution.  Against the reference's FP64 products:
 *   ||out - ref||_2 / ||ref||_2  < 1e-6   and   ||out - ref||_inf / ||ref||_inf < 1e-6
 * also under adversarial dynamic range (one entry 10^6 x the rest in u or v, stripes 10^4 apart, 128 clustered large effects in one
 * tile over a tiny background, 99.9 % sparse v, |u| ~ 1e-150 with |v| ~ 1e150; tests/test_gpu_kernels.py::
 * test_fixed_point_dynamic_range); typical inputs give 5e-8 (X.v) / 7e-8 (X^T.u).  On shards with missing genotypes the
 * missing-genotype term of X^T.u keeps the common scale of the whole vector: with a fraction p of missing genotypes the bound
 * holds while max|u| / rms(u) stays below ~30 / sqrt(p) (300 at 1 % missing; a Gaussian u of 400k entries has 4.8).  Outputs that
 * nearly cancel are accurate to that absolute error, not to 1e-6 of their own value.  Results do not depend on the tiling, the
 * table staging mode or the amount of twin (bit-identical), and are the same on every run.  A NaN / infinity in the input turns every
 * output into NaN, like the reference's sums. */
/* data::Ax(double* v, SB, LB), data.cpp:848-1011 incl. the MPI_Allreduce at :995.
 * v: M local entries; out: 4*LB entries.  COLLECTIVE when nranks>1. */
int gvb_Ax(gvb_ctx* ctx, const double* v, double* out, long SB, long LB);
/* data::ATx(double* u, SB, LB) + dot_product, data.cpp:728-835.  u: 4*LB entries (entries that
 * would index individuals >= N are ignored); out: M entries.  Local, no communication. */
int gvb_ATx(gvb_ctx* ctx, const double* u, double* out, long SB, long LB);

/* sum_i a_i u_i and sum_i b_i u_i of ONE local marker over the byte range [SB, SB+LB): the two
 * accumulators of data::dot_product (data.cpp:728-801) before the caller's mu / sigma are applied. */
int gvb_marker_dot(gvb_ctx* ctx, long mloc, const double* u, long SB, long LB, double* dpa, double* dpb);
/* in-place sum over ranks of n host doubles: MPI_Allreduce(MPI_SUM) for the driver's host scalars */
int gvb_allreduce_host(gvb_ctx* ctx, double* buf, int n);

/* ---- device vectors ------------------------------------------------------------------------------ */
int gvb_vec_alloc(gvb_ctx* ctx, long n, gvb_vec* out);          /* zero-initialised */
int gvb_vec_alloc_M(gvb_ctx* ctx, gvb_vec* out);                /* length M (local markers) */
int gvb_vec_alloc_N(gvb_ctx* ctx, gvb_vec* out);                /* length 4*ceil(N/4) */
void gvb_vec_free(gvb_ctx* ctx, gvb_vec v);
long gvb_vec_len(gvb_vec v);
void* gvb_vec_ptr(gvb_vec v);                                   /* raw device pointer */
int gvb_vec_upload(gvb_ctx* ctx, gvb_vec dst, const double* src, long n);
int gvb_vec_download(gvb_ctx* ctx, gvb_vec src, double* dst, long n);
/* gvb_vec_upload that also reports whether any bit of dst[0..n) changed (*changed = 0 / 1).  `stage` is a scratch vector of at
 * least n entries (overwritten).  For host code that keeps products of an input vector: vamp::infere_linear computes A^T y in every
 * iteration from the y it holds in host memory (vamp.cpp:588); here A^T y is kept while the y that arrives is the one already
 * resident. */
int gvb_vec_upload_changed(gvb_ctx* ctx, gvb_vec dst, gvb_vec stage, const double* src, long n, int* changed);
/* Asynchronous snapshot of the first n entries of a vector into pinned host memory: the per-iteration outputs of
 * vamp::infere_linear (z1 / x1_hat / r1 / r2 / x2_hat files, vamp.cpp:435-462,542,612) leave the device while the LMMSE
 * sweeps run.  begin: ordered on the library stream like a copy kernel (the vector may be overwritten right after the
 * call); wait: blocks until that snapshot is in host memory and returns the library-owned pinned buffer, valid until
 * the next begin on the same slot (0..15) or gvb_ctx_destroy. */
int gvb_snapshot_begin(gvb_ctx* ctx, gvb_vec src, long n, int slot);
int gvb_snapshot_wait(gvb_ctx* ctx, int slot, const double** host, long* n);
int gvb_vec_copy(gvb_ctx* ctx, gvb_vec dst, gvb_vec src);
int gvb_vec_fill(gvb_ctx* ctx, gvb_vec dst, double value);
/* out = a*x + b*y (y may be NULL); the M-vector algebra of infere_linear, vamp.cpp:348-354,485-486,
 * 590-591,706-707 */
int gvb_vec_axpby(gvb_ctx* ctx, gvb_vec out, double a, gvb_vec x, double b, gvb_vec y);
/* out = (a*x + b*y) / div, evaluated in that order like r2 = (eta1*x1 - gam1*r1)/gam2, vamp.cpp:486,707 */
int gvb_vec_axpby_div(gvb_ctx* ctx, gvb_vec out, double a, gvb_vec x, double b, gvb_vec y, double div);
/* res[k] = <x[k], y[k]> (y[k]==NULL: squared norm); inner_prod / l2_norm2, utilities.cpp:190-214.
 * sync!=0 sums over ranks (one fused NCCL allreduce of n scalars instead of n MPI_Allreduce). */
int gvb_vec_dots(gvb_ctx* ctx, int n, const gvb_vec* x, const gvb_vec* y, int sync, double* res);
/* A batch of reductions over vectors of any lengths in ONE host synchronisation (two launches, one fused allreduce):
 *   GVB_RED_DOT: res[k] = <x, y>  (y == NULL: ||x||^2)      GVB_RED_SQ: res[k] = ||a*x + b*y||^2  (y == NULL: ||a*x||^2)
 * sync != 0 sums over the ranks (M-vectors); such operations must come first in the batch.  Replaces the separate
 * inner_prod / l2_norm2 allreduces of the iteration's diagnostics and exit tests (vamp.cpp:741-749, 892-927, 1232-1317). */
#define GVB_RED_BATCH 16
enum { GVB_RED_DOT = 0, GVB_RED_SQ = 1 };
typedef struct {
    gvb_vec x, y;
    double a, b;
    int kind, sync;
} gvb_red_op;
int gvb_vec_reduce_batch(gvb_ctx* ctx, int nops, const gvb_red_op* ops, double* res);
/* res[0] = ||x - y||^2 (local or rank-summed) -- the gamma re-estimation residuals, vamp.cpp:326,692 */
int gvb_vec_dist2(gvb_ctx* ctx, gvb_vec x, gvb_vec y, int sync, double* res);

/* device-resident matvecs used by the fused loop (same math as gvb_Ax / gvb_ATx, full range) */
int gvb_dAx(gvb_ctx* ctx, gvb_vec v, gvb_vec out);
int gvb_dATx(gvb_ctx* ctx, gvb_vec u, gvb_vec out);
/* Two products from ONE pass over the bed: out0 = X.v0, out1 = X.v1 (both all-reduced like gvb_dAx).  Where the reference calls
 * data::Ax twice on independent vectors (z1 = A x1_hat, vamp.cpp:430, and the first operator application of the LMMSE solve,
 * vamp.cpp:1146) the bed is read once.  Bit-identical to two gvb_dAx calls. */
int gvb_dAx2(gvb_ctx* ctx, gvb_vec v0, gvb_vec v1, gvb_vec out0, gvb_vec out1);
/* ... and the same for X^T.u: out0 = X^T.u0, out1 = X^T.u1 from one pass (shards with missing genotypes: two single sweeps). */
int gvb_dATx2(gvb_ctx* ctx, gvb_vec u0, gvb_vec u1, gvb_vec out0, gvb_vec out1);

/* ---- denoiser and EM prior update ---------------------------------------------------------------- */
/* x1_hat = g1(r1), sums[0] = sum_i g1d(r1_i) over ALL ranks, sums[1] = ||x1_hat - r1||^2 over all
 * ranks: vamp::g1 / g1d, vamp.cpp:805-869 and the allreduce at :313. */
int gvb_denoise(gvb_ctx* ctx, gvb_vec r1, double gam1, const double* probs, const double* vars, int L, gvb_vec x1_hat,
                double* sums);
/* One E-step of vamp::updatePrior, vamp.cpp:944-1013: sums[0] = sum pin; sums[1..L-1] = sum
 * beta_j*pin; sums[L..2L-2] = sum beta_j*(m_j^2+v_j)*pin, all summed over ranks in ONE allreduce. */
int gvb_em_stats(gvb_ctx* ctx, gvb_vec r1, double gam1, double lambda, const double* omegas, const double* vars, int L,
                 double* sums);

/* ---- LMMSE: operator, preconditioned CG, Onsager, noise precision ---------------------------------- */
/* vamp::lmmse_mult, vamp.cpp:1074-1118: out = tau * ATx(Ax(v)) + gam2 * v */
int gvb_lmmse_mult(gvb_ctx* ctx, gvb_vec v, double tau, double gam2, gvb_vec out);
/* vamp::precondCG_solver, vamp.cpp:1130-1229.  mu holds the start vector on entry and the solution
 * on exit.  denoiser==0 is the Onsager mode (exit on the trace estimate, :1174-1193).
 * iters: iterations run; rel_res: ||r||/||rhs|| per iteration (room for max_iter doubles, may be NULL). */
int gvb_cg_solve(gvb_ctx* ctx, gvb_vec rhs, gvb_vec mu, double tau, double gam2, int max_iter, int denoiser, int* iters,
                 double* rel_res);

/* ---- association tests (LOO / LOCO p-values) --------------------------------------------------------- */
/* The per-marker test of data::pvals_calc (data.cpp:1108-1180) and of the chromosome loop of data::pvals_calc_LOCO
 * (data.cpp:1311-1343) with linear_reg1d_pvals (utilities.cpp:321-334): for every local marker j, regress
 * yres + x_j * coef[j]/sqrt(N) on the standardised genotype column x_j over the individuals whose phenotype AND genotype
 * are present, pvals[j] = two-sided Student-t p-value of the slope.  yres: N-vector (y - z1, resp. y - z1 + the
 * chromosome's predictor); coef: the estimate whose own contribution is added back (x1_hat, LOO) or NULL (LOCO);
 * select: optional M-vector, markers with select[j]==0 keep their old pvals[j] (LOCO: the markers of one chromosome).
 * Two bed sweeps, local to the shard (no communication). */
int gvb_assoc_pvals(gvb_ctx* ctx, gvb_vec yres, gvb_vec coef, gvb_vec select, gvb_vec pvals);

/* The same solver with its by-products (no extra bed sweep; either may be NULL): ax_mu <- A * (returned mu), accumulated from
 * the A p_k the iterations form anyway (replaces the separate data::Ax(x2_hat) of vamp.cpp:897,1301 and of vamp_probit.cpp:567);
 * dots3 <- {<rhs,rhs>, <rhs,mu>, <rhs,r>} over all ranks, r = rhs - (tau A^T A + gam2) mu, from which
 * <rhs, A^T A mu> = (dots3[0] - gam2*dots3[1] - dots3[2]) / tau (replaces the Ax + ATx of vamp.cpp:908-915). */
int gvb_cg_solve_ex(gvb_ctx* ctx, gvb_vec rhs, gvb_vec mu, double tau, double gam2, int max_iter, int denoiser, int* iters,
                    double* rel_res, gvb_vec ax_mu, double* dots3);

/* Warm-started variant: ata_mu (M-vector) additionally carries A^T A * mu.  have_start == 1: on entry ax_mu / ata_mu hold A mu and
 * A^T A mu of the START vector in mu (the outputs of the previous call that produced it; tau / gam2 may have changed since), and
 * the initial residual of precondCG_solver (vamp.cpp:1141-1146, two sweeps in the reference) is formed from them without a sweep.
 * have_start == 2: the caller states that the start vector is zero (mu is cleared): no sweep and no host-visible norm -- the
 * zero-vector shortcut of lmmse_mult, vamp.cpp:1079-1080.  have_start == 0: any start vector (one host-visible norm decides
 * whether the shortcut applies).  An Onsager-mode solve (denoiser == 0) with have_start == 1 stops on ||r||/||rhs|| < 1e-5 only:
 * the reference's successive-difference exit (vamp.cpp:1174-1193) presumes the monotone <u,mu_k> of a zero start.
 * On exit both by-products describe the returned mu.
 * The solver's scalars (alpha, beta, the exit tests) live on the device: the stream does not drain inside a solve. */
int gvb_cg_solve_warm(gvb_ctx* ctx, gvb_vec rhs, gvb_vec mu, double tau, double gam2, int max_iter, int denoiser, int* iters,
                      double* rel_res, gvb_vec ax_mu, gvb_vec ata_mu, int have_start, double* dots3);
/* The same solve in two calls, so that its first product A p0 shares ONE bed read with a product the caller needs anyway (z1 = A x1_hat
 * of the VAMP iteration, vamp.cpp:430): gvb_cg_prepare forms the initial residual and the first search direction (have_start 1 or 2 as
 * above) and runs one dual sweep {A p0, extra_out = A extra_v}; gvb_cg_solve_prepared, called with the same vectors and scalars, runs
 * the iterations.  Anything may be enqueued between the two calls except another solve.  Bit-identical to gvb_dAx(extra_v, extra_out)
 * followed by gvb_cg_solve_warm. */
int gvb_cg_prepare(gvb_ctx* ctx, gvb_vec rhs, gvb_vec mu, double tau, double gam2, int max_iter, gvb_vec ax_mu, gvb_vec ata_mu, int have_start,
                   gvb_vec extra_v, gvb_vec extra_out);
int gvb_cg_solve_prepared(gvb_ctx* ctx, gvb_vec rhs, gvb_vec mu, double tau, double gam2, int max_iter, int denoiser, int* iters, double* log4,
                          gvb_vec ax_mu, gvb_vec ata_mu, int have_start, double* dots3);
/* A companion for the NEXT solve of this context (consumed by it; fn = NULL clears).  Before the product A p of iteration i the solver
 * calls fn(user, 0, i, &v, &av, &w); when that returns 1 and sets an M-vector v and an N-vector av, the iteration's X.v is ONE dual
 * sweep {A p, av = A v}; when it also sets an M-vector w, the iteration's X^T.u is a dual sweep too, {A^T (A p), w = A^T av} (w = NULL:
 * the companion sweeps for it itself).  Then the solver calls fn(user, 1, i, &v, &av, &w), where the companion finishes its own step (it
 * may enqueue sweeps and synchronise; a negative return aborts the solve).  The solver's own results do not change by a bit.  Used by
 * vamp::infere_linear to let the Lanczos steps of the Onsager projection (the reference's second solve with the same operator,
 * vamp.cpp:884) ride on the LMMSE solve of the first iteration (vamp.cpp:594-599). */
typedef int (*gvb_cg_companion_fn)(void* user, int stage, int iteration, gvb_vec* v, gvb_vec* av, gvb_vec* w);
int gvb_cg_set_companion(gvb_ctx* ctx, gvb_cg_companion_fn fn, void* user);

/* Zero-start solve against a right-hand side that recurs (the Onsager probe: the same Rademacher vector in every VAMP iteration,
 * vamp.cpp:875-882).  mu is cleared.  ata_rhs (M-vector) caches A^T A rhs: with a zero start the first search direction is rhs / diag,
 * so the operator product of CG iteration 0 is (tau * ata_rhs + gam2 * rhs) / diag and needs no bed sweep.  *ata_rhs_state: 0 = the
 * cache is empty (iteration 0 sweeps and fills it, the state becomes 1), 1 = filled for THIS rhs and THIS matrix / mask (the caller
 * resets it to 0 when either changes, see gvb_layout_generation).  Same iterates as gvb_cg_solve up to the fixed-point error of the
 * sweeps; two sweeps less per solve. */
int gvb_cg_solve_cached(gvb_ctx* ctx, gvb_vec rhs, gvb_vec mu, double tau, double gam2, int max_iter, int denoiser, int* iters,
                        double* rel_res, gvb_vec ata_rhs, int* ata_rhs_state, double* dots3);

/* ---- XXT form of the LMMSE step (--use-XXT-denoiser 1) ---------------------------------------------- */
/* data::compute_people_statistics, data.cpp:548-640: per individual, over ALL markers (summed over the ranks): number of
 * non-missing genotypes, mean and inverse-variance-like scale sqrt((n-1)/(S2 - n mean^2)) of the standardised genotypes;
 * 0 for individuals without phenotype.  Three X.v-type sweeps.  COLLECTIVE when nranks>1. */
int gvb_people_stats(gvb_ctx* ctx, gvb_vec mave_people, gvb_vec msig_people, gvb_vec numb_people);
/* vamp::CG_solverAAT + lmmse_multAAT, denoiserXXT.cpp:15-135: (tau A A^T + gam2 I) mu = rhs on N-vectors with the
 * per-individual diagonal preconditioner built from the people statistics; mu holds the start vector on entry; exit at
 * ||r||/||rhs|| < 1e-4.  log3: 3 doubles per iteration (||r||/||rhs||, ||mu||, ||z||/||rhs||), may be NULL.  COLLECTIVE. */
int gvb_cg_solve_aat(gvb_ctx* ctx, gvb_vec rhs, gvb_vec mu, double tau, double gam2, gvb_vec mave_people, gvb_vec msig_people,
                     gvb_vec numb_people, int max_iter, int* iters, double* log3);

/* ---- probit z-denoiser ---------------------------------------------------------------------------- */
/* vamp::g1_bin_class / g1d_bin_class, vamp_probit.cpp:661-726 with erfcx (utilities.cpp:345-409):
 * z1_hat[i] = g(p1[i]); sums[0] = sum_i g'(p1[i]) (i<N); sums[1] = ||z1_hat - p1||^2.  mcov may be NULL. */
int gvb_probit_denoise(gvb_ctx* ctx, gvb_vec p1, gvb_vec y, gvb_vec mcov, double tau1, double probit_var, gvb_vec z1_hat,
                       double* sums);

/* ---- probit covariate effects --------------------------------------------------------------------- */
/* One pass over the N x C covariate matrix Z (row-major device vector of N*C doubles, 1 <= C <= 32) at the covariate
 * effects eta (C host doubles): the inner loops of vamp::Newton_method_cov (vamp_probit.cpp:936-1067), grad_cov
 * (:814-839) and mlogL_probit (:841-858).  gg: genetic part of the liability (N-vector) or NULL for zeros.
 * out (host, 1 + 2C + C*C doubles): [0] mlogL_probit, [1..C] grad_cov, [1+C..2C] Newton numerator sum_i Z_ij lambda_i,
 * [1+2C..] Hessian, row-major.  `what` selects the Hessian (bit 2); the cheap parts are always returned. */
int gvb_probit_cov_pass(gvb_ctx* ctx, gvb_vec y, gvb_vec gg, gvb_vec Z, int C, const double* eta, double probit_var, int what, double* out);
/* mcov[i] = <Z_i, eta>: the covariate part of the liability (vamp_probit.cpp:132-137) */
int gvb_probit_cov_apply(gvb_ctx* ctx, gvb_vec Z, int C, const double* eta, gvb_vec mcov);

#ifdef __cplusplus
}
#endif
#endif /* GVAMP_B200_H */
