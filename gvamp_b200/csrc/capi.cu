// capi.cu -- context management, device vectors, the host-pointer Ax/ATx drop-ins, the LMMSE
// operator and the preconditioned CG driver of the C ABI declared in include/gvamp_b200.h.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "gvb_internal.cuh"

static thread_local char g_err[512] = "";

void gvb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* gvb_last_error(void) { return g_err; }
extern "C" const char* gvb_version(void) { return "gvamp_b200 0.1 (sm_100a)"; }

extern "C" int gvb_device_count(int* n) {
    GVB_ARG(n, "n");
    *n = 0;
    GVB_CUDA(cudaGetDeviceCount(n));
    return GVB_OK;
}

extern "C" int gvb_nccl_unique_id(void* id128) {
    GVB_ARG(id128, "id128");
    static_assert(sizeof(ncclUniqueId) == GVB_NCCL_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    GVB_NCCL(ncclGetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return GVB_OK;
}

extern "C" void gvb_divide_work(long Mt, int nranks, int rank, long* M, long* S) {
    // utilities.cpp:266-281
    long modu = Mt % nranks, size = Mt / nranks;
    long start = 0;
    for (int i = 0; i < rank; i++) start += (i < modu) ? size + 1 : size;
    if (M) *M = (rank < modu) ? size + 1 : size;
    if (S) *S = start;
}

// The product library ships ONE implementation of the sweeps (the tile kernels of matvec_tile.cu).  The earlier generations -- FP64
// decode-multiply kernels ("simple") and the first table kernels ("lut1") -- are cross-checks for the tests and exist only in the
// test build of the library (gvamp_b200/lib/xcheck/, compiled with -DGVB_LEGACY_KERNELS), selected there by env GVB_KERNELS.
static int pick_kernel_gen(int* out) {
    const char* gen = getenv("GVB_KERNELS");
    *out = 2;
    if (gen && (!strcmp(gen, "simple") || !strcmp(gen, "lut1"))) {
#ifdef GVB_LEGACY_KERNELS
        *out = !strcmp(gen, "simple") ? 0 : 1;
#else
        gvb_set_error("GVB_KERNELS=%s: the cross-check kernels are not part of the product library (the tests load gvamp_b200/lib/xcheck/libgvamp_b200.so)", gen);
        return GVB_ERR_ARG;
#endif
    }
    return GVB_OK;
}

static std::vector<gvb_ctx*> g_live_ctx;   // every context of this process (one host thread drives the library, see the header)

int gvb_release_twins(int device) {
    int n = 0;
    for (gvb_ctx* x : g_live_ctx) {
        if (x->device != device || !x->bed_twin) continue;
        cudaStreamSynchronize(x->stream);   // a sweep on the twin may still be running
        gvb_twin_reset(x);
        x->twin_state = -1;                 // not rebuilt: the memory is needed elsewhere
        n++;
    }
    return n;
}

static int ctx_common_init(gvb_ctx* c) {
    g_live_ctx.push_back(c);
    GVB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 8; i++) {
        GVB_CUDA(cudaEventCreate(&c->ev_start[i]));
        GVB_CUDA(cudaEventCreate(&c->ev_stop[i]));
    }
    GVB_CUDA(gvb_malloc(c, &c->red_partial, sizeof(double) * GVB_RED_BLOCKS * GVB_RED_MAXK));
    GVB_CUDA(gvb_malloc(c, &c->red_result, sizeof(double) * GVB_RED_MAXK));
    GVB_CUDA(cudaMallocHost(&c->h_red, sizeof(double) * GVB_RED_MAXK));
    GVB_CUDA(gvb_malloc(c, &c->scal, sizeof(double) * 64));
    GVB_CUDA(cudaMemset(c->scal, 0, sizeof(double) * 64));
    GVB_CUDA(gvb_malloc(c, &c->work_counter, sizeof(int) * 16));
    GVB_CUDA(cudaMemset(c->work_counter, 0, sizeof(int) * 16));
    GVB_CHECK(pick_kernel_gen(&c->kernel_gen));
    c->tab_pairs = c->kernel_gen == 2 ? 1 : 0;   // the tile kernels' tables are pair-interleaved in global memory (matvec_tile.cu)
    return GVB_OK;
}

extern "C" int gvb_ctx_create_shared(gvb_ctx** out, gvb_ctx* parent) {
    GVB_ARG(out && parent, "out / parent");
    GVB_CUDA(cudaSetDevice(parent->device));
    int gen_probe = 2;
    GVB_CHECK(pick_kernel_gen(&gen_probe));
    gvb_ctx* c = new gvb_ctx();
    c->device = parent->device;
    c->rank = parent->rank;
    c->nranks = parent->nranks;
    c->sm_count = parent->sm_count;
    c->comm = parent->comm;
    c->owns_comm = false;
    int rc = ctx_common_init(c);
    if (rc != GVB_OK) return rc;
    *out = c;
    return GVB_OK;
}

extern "C" int gvb_ctx_create(gvb_ctx** out, int device, int rank, int nranks, const void* id128) {
    GVB_ARG(out, "out");
    GVB_ARG(nranks >= 1 && rank >= 0 && rank < nranks, "rank / nranks");
    GVB_ARG(nranks == 1 || id128, "a NCCL unique id is required when nranks > 1");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        gvb_set_error("no CUDA device available (%s): libgvamp_b200 has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return GVB_ERR_CUDA;
    }
    GVB_ARG(device >= 0 && device < ndev, "device ordinal");
    GVB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    GVB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        gvb_set_error("device %d (%s, sm_%d%d) is not a Blackwell part: this library ships sm_100a code only", device, prop.name, prop.major, prop.minor);
        return GVB_ERR_CUDA;
    }
    int gen_probe = 2;
    GVB_CHECK(pick_kernel_gen(&gen_probe));
    gvb_ctx* c = new gvb_ctx();
    c->device = device;
    c->rank = rank;
    c->nranks = nranks;
    c->sm_count = prop.multiProcessorCount;
    GVB_CHECK(ctx_common_init(c));
    if (nranks > 1) {
        ncclUniqueId id;
        memcpy(&id, id128, sizeof(id));
        GVB_NCCL(ncclCommInitRank(&c->comm, nranks, id, rank));
    }
    *out = c;
    return GVB_OK;
}

extern "C" void gvb_ctx_destroy(gvb_ctx* c) {
    if (!c) return;
    g_live_ctx.erase(std::remove(g_live_ctx.begin(), g_live_ctx.end(), c), g_live_ctx.end());
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (gvb_vec_s* v : c->vecs) {
        if (v) { cudaFree(v->d); delete v; }
    }
    auto fr = [](auto*& p) { if (p) { cudaFree(p); p = nullptr; } };
    fr(c->bed); fr(c->maskw); fr(c->validw); fr(c->mave); fr(c->msig); fr(c->counts);
    fr(c->tmpN); fr(c->tmpN2); fr(c->tmpM); fr(c->tmpM2); fr(c->wv); fr(c->cv); fr(c->ax_partial);
    gvb_misslist_reset(c);
    gvb_twin_reset(c);
    fr(c->tab_u); fr(c->tab_v); fr(c->tab_v2); fr(c->shift_v2); fr(c->acc_dual); fr(c->tab_u2); fr(c->shift_u2); fr(c->cg_ap); fr(c->shift_u); fr(c->shift_v); fr(c->acc_i64); fr(c->red_partial); fr(c->red_result); fr(c->scal); fr(c->work_counter);
    if (c->h_red) cudaFreeHost(c->h_red);
    fr(c->cg_dev); fr(c->cg_flags);
    if (c->cg_host) cudaFreeHost(c->cg_host);
    for (auto& e : c->cg_ev) if (e) cudaEventDestroy(e);
    for (auto& sn : c->snap) {
        if (sn.dev) cudaFree(sn.dev);
        if (sn.host) cudaFreeHost(sn.host);
        if (sn.ready) cudaEventDestroy(sn.ready);
        if (sn.done) cudaEventDestroy(sn.done);
    }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (int i = 0; i < 8; i++) { cudaEventDestroy(c->ev_start[i]); cudaEventDestroy(c->ev_stop[i]); }
    if (c->comm && c->owns_comm) ncclCommDestroy(c->comm);
    cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int gvb_ctx_sync(gvb_ctx* c) {
    GVB_ARG(c, "ctx");
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    return GVB_OK;
}
extern "C" void* gvb_ctx_stream(gvb_ctx* c) { return c ? (void*)c->stream : nullptr; }
extern "C" int gvb_ctx_info(gvb_ctx* c, long* N, long* Mt, long* S, long* M, long* mbytes) {
    GVB_ARG(c, "ctx");
    if (N) *N = c->N;
    if (Mt) *Mt = c->Mt;
    if (S) *S = c->S;
    if (M) *M = c->M;
    if (mbytes) *mbytes = c->mbytes;
    return GVB_OK;
}
extern "C" int gvb_timer_start(gvb_ctx* c, int slot) {
    GVB_ARG(c && slot >= 0 && slot < 8, "slot");
    GVB_CUDA(cudaEventRecord(c->ev_start[slot], c->stream));
    return GVB_OK;
}
extern "C" int gvb_timer_stop(gvb_ctx* c, int slot) {
    GVB_ARG(c && slot >= 0 && slot < 8, "slot");
    GVB_CUDA(cudaEventRecord(c->ev_stop[slot], c->stream));
    return GVB_OK;
}
extern "C" int gvb_timer_elapsed_ms(gvb_ctx* c, int slot, float* ms) {
    GVB_ARG(c && slot >= 0 && slot < 8 && ms, "slot");
    GVB_CUDA(cudaEventSynchronize(c->ev_stop[slot]));
    GVB_CUDA(cudaEventElapsedTime(ms, c->ev_start[slot], c->ev_stop[slot]));
    return GVB_OK;
}
extern "C" long gvb_launch_count(gvb_ctx* c) { return c ? c->launches : 0; }
extern "C" long gvb_sweep_count(gvb_ctx* c) { return c ? c->sweeps : 0; }
extern "C" long gvb_host_sync_count(gvb_ctx* c) { return c ? c->host_syncs : 0; }
extern "C" long gvb_layout_generation(gvb_ctx* c) { return c ? c->layout_gen : 0; }

// ------------------------------------------------------------------------------------------------
// device vectors
// ------------------------------------------------------------------------------------------------
static int vec_alloc_cap(gvb_ctx* c, long n, long cap, gvb_vec* out) {
    GVB_ARG(c && out && n > 0 && cap >= n, "vector length");
    GVB_CUDA(cudaSetDevice(c->device));
    gvb_vec_s* v = new gvb_vec_s();
    v->n = n;
    v->cap = cap;
    cudaError_t e = gvb_malloc(c, &v->d, cap * sizeof(double));
    if (e != cudaSuccess) {
        delete v;
        gvb_set_error("cudaMalloc of a %ld-double vector failed: %s", cap, cudaGetErrorString(e));
        return GVB_ERR_NOMEM;
    }
    GVB_CUDA(cudaMemsetAsync(v->d, 0, cap * sizeof(double), c->stream));
    c->vecs.push_back(v);
    *out = v;
    return GVB_OK;
}
extern "C" int gvb_vec_alloc(gvb_ctx* c, long n, gvb_vec* out) { return vec_alloc_cap(c, n, n, out); }
// M- and N-vectors carry zero padding up to the layout's tile sizes so the sweep kernels never branch on the tail
extern "C" int gvb_vec_alloc_M(gvb_ctx* c, gvb_vec* out) {
    GVB_ARG(c && c->bed, "matrix not loaded");
    return vec_alloc_cap(c, c->M, c->Mg_pad * 4, out);
}
extern "C" int gvb_vec_alloc_N(gvb_ctx* c, gvb_vec* out) {
    GVB_ARG(c && c->bed, "matrix not loaded");
    return vec_alloc_cap(c, 4 * c->mbytes, c->Npad, out);
}
extern "C" void gvb_vec_free(gvb_ctx* c, gvb_vec v) {
    if (!c || !v) return;
    auto it = std::find(c->vecs.begin(), c->vecs.end(), v);
    if (it != c->vecs.end()) c->vecs.erase(it);
    cudaStreamSynchronize(c->stream);
    cudaFree(v->d);
    delete v;
}
extern "C" long gvb_vec_len(gvb_vec v) { return v ? v->n : 0; }
extern "C" void* gvb_vec_ptr(gvb_vec v) { return v ? (void*)v->d : nullptr; }
extern "C" int gvb_vec_upload(gvb_ctx* c, gvb_vec dst, const double* src, long n) {
    GVB_ARG(c && dst && src && n >= 0 && n <= dst->n, "upload length");
    GVB_CUDA(cudaMemcpyAsync(dst->d, src, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    GVB_CUDA(cudaStreamSynchronize(c->stream));   // src may be pageable / reused by the caller
    return GVB_OK;
}
extern "C" int gvb_vec_download(gvb_ctx* c, gvb_vec src, double* dst, long n) {
    GVB_ARG(c && dst && src && n >= 0 && n <= src->n, "download length");
    GVB_CUDA(cudaMemcpyAsync(dst, src->d, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    return GVB_OK;
}
// Upload that reports whether the vector's content changed (bitwise): host code that caches a product of the vector (A^T y of the
// linear model) keeps it when the caller re-sends identical data.
__global__ void upload_compare_kernel(unsigned long long* __restrict__ dst, const unsigned long long* __restrict__ stage, long n, int* __restrict__ changed) {
    int diff = 0;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        unsigned long long a = stage[i];
        if (a != dst[i]) {
            dst[i] = a;
            diff = 1;
        }
    }
    if (__syncthreads_or(diff) && threadIdx.x == 0) atomicOr(changed, 1);
}
extern "C" int gvb_vec_upload_changed(gvb_ctx* c, gvb_vec dst, gvb_vec stage, const double* src, long n, int* changed) {
    GVB_ARG(c && dst && stage && src && changed && dst != stage && n >= 0 && n <= dst->n && n <= stage->n, "upload length / staging vector");
    int* flag = reinterpret_cast<int*>(c->scal + 60);
    GVB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), c->stream));
    GVB_CUDA(cudaMemcpyAsync(stage->d, src, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (n > 0) {
        upload_compare_kernel<<<(unsigned)std::min<long>((n + 255) / 256, 1184), 256, 0, c->stream>>>(
            reinterpret_cast<unsigned long long*>(dst->d), reinterpret_cast<const unsigned long long*>(stage->d), n, flag);
        GVB_LAUNCHED(c);
    }
    GVB_CUDA(cudaMemcpyAsync(changed, flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    c->host_syncs++;
    return GVB_OK;
}
// Snapshots: the vector is copied to a staging buffer on the library stream (ordered with the kernels that produce and later
// overwrite it), then leaves for pinned host memory on a second stream, overlapping whatever the library stream does next.
extern "C" int gvb_snapshot_begin(gvb_ctx* c, gvb_vec src, long n, int slot) {
    GVB_ARG(c && src && n >= 0 && n <= src->n, "snapshot length");
    GVB_ARG(slot >= 0 && slot < GVB_SNAP_SLOTS, "snapshot slot");
    gvb_ctx::Snap& sn = c->snap[slot];
    if (!c->copy_stream) GVB_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    if (!sn.ready) {
        GVB_CUDA(cudaEventCreateWithFlags(&sn.ready, cudaEventDisableTiming));
        GVB_CUDA(cudaEventCreateWithFlags(&sn.done, cudaEventDisableTiming));
    }
    if (sn.pending) GVB_CUDA(cudaEventSynchronize(sn.done));   // an unread snapshot in this slot is simply replaced
    if (sn.cap < n) {
        if (sn.dev) cudaFree(sn.dev);
        if (sn.host) cudaFreeHost(sn.host);
        sn.dev = sn.host = nullptr;
        sn.cap = 0;
        GVB_CUDA(gvb_malloc(c, &sn.dev, std::max(n, 1l) * sizeof(double)));
        GVB_CUDA(cudaMallocHost(&sn.host, std::max(n, 1l) * sizeof(double)));
        sn.cap = n;
    }
    sn.n = n;
    GVB_CUDA(cudaMemcpyAsync(sn.dev, src->d, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    GVB_CUDA(cudaEventRecord(sn.ready, c->stream));
    GVB_CUDA(cudaStreamWaitEvent(c->copy_stream, sn.ready, 0));
    GVB_CUDA(cudaMemcpyAsync(sn.host, sn.dev, n * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
    GVB_CUDA(cudaEventRecord(sn.done, c->copy_stream));
    sn.pending = true;
    return GVB_OK;
}
extern "C" int gvb_snapshot_wait(gvb_ctx* c, int slot, const double** host, long* n) {
    GVB_ARG(c && slot >= 0 && slot < GVB_SNAP_SLOTS && host, "snapshot slot / pointer");
    gvb_ctx::Snap& sn = c->snap[slot];
    GVB_ARG(sn.done && sn.host, "no snapshot was begun in this slot");
    if (sn.pending) GVB_CUDA(cudaEventSynchronize(sn.done));
    sn.pending = false;
    *host = sn.host;
    if (n) *n = sn.n;
    return GVB_OK;
}
extern "C" int gvb_vec_copy(gvb_ctx* c, gvb_vec dst, gvb_vec src) {
    GVB_ARG(c && dst && src && dst->n == src->n, "vector lengths");
    GVB_CUDA(cudaMemcpyAsync(dst->d, src->d, src->n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    return GVB_OK;
}
__global__ void fill_kernel(double* d, double v, long n) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) d[i] = v;
}
extern "C" int gvb_vec_fill(gvb_ctx* c, gvb_vec dst, double value) {
    GVB_ARG(c && dst, "vector");
    if (value == 0.0) {
        GVB_CUDA(cudaMemsetAsync(dst->d, 0, dst->n * sizeof(double), c->stream));
        return GVB_OK;
    }
    fill_kernel<<<(unsigned)std::min((dst->n + 255) / 256, 1184l), 256, 0, c->stream>>>(dst->d, value, dst->n);
    GVB_LAUNCHED(c);
    return GVB_OK;
}

// ------------------------------------------------------------------------------------------------
// mat-vec dispatch (kernel generation) + NCCL allreduce of the partial N-vector
// ------------------------------------------------------------------------------------------------
static void prof_mark(gvb_ctx* c, int which) {
    if (!c->profile) return;
    if (c->prof_used[which] >= c->prof_ev[which].size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        c->prof_ev[which].push_back(e);
    }
    cudaEventRecord(c->prof_ev[which][c->prof_used[which]++], c->stream);
}

extern "C" int gvb_profile_enable(gvb_ctx* c, int on) {
    GVB_ARG(c, "ctx");
    c->profile = on != 0;
    c->prof_used[0] = c->prof_used[1] = c->prof_used[2] = 0;
    return GVB_OK;
}

static int profile_sum(gvb_ctx* c, int w, double* tot, double* cnt) {
    *tot = 0;
    size_t n = c->prof_used[w] / 2;
    for (size_t i = 0; i < n; i++) {
        float ms = 0;
        GVB_CUDA(cudaEventElapsedTime(&ms, c->prof_ev[w][2 * i], c->prof_ev[w][2 * i + 1]));
        *tot += ms;
    }
    *cnt = (double)n;
    c->prof_used[w] = 0;
    return GVB_OK;
}
extern "C" int gvb_profile_read(gvb_ctx* c, double* out4) {
    GVB_ARG(c && out4, "ctx / out");
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    for (int w = 0; w < 2; w++) GVB_CHECK(profile_sum(c, w, out4 + 2 * w, out4 + 2 * w + 1));
    return GVB_OK;
}
// total time and number of the dual X.v sweeps since the last read (call before or after gvb_profile_read)
extern "C" int gvb_profile_read_dual(gvb_ctx* c, double* out2) {
    GVB_ARG(c && out2, "ctx / out");
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    return profile_sum(c, 2, out2, out2 + 1);
}
extern "C" long gvb_dual_sweep_count(gvb_ctx* c) { return c ? c->dual_sweeps : 0; }

int gvb_ax_dev(gvb_ctx* c, const double* v, double* out, bool allreduce) {
    GVB_ARG(c->have_stats, "compute_stats must run before Ax");
    prof_mark(c, 0);
#ifdef GVB_LEGACY_KERNELS
    int rc_mv = c->kernel_gen == 0 ? gvb_ax_simple(c, v, out) : (c->kernel_gen == 1 ? gvb_ax_lut(c, v, out) : gvb_ax_tile(c, v, out));
#else
    int rc_mv = gvb_ax_tile(c, v, out);
#endif
    prof_mark(c, 0);
    GVB_CHECK(rc_mv);
    c->sweeps++;
    if (allreduce && c->nranks > 1) {
        // replaces MPI_Allreduce(Ax_temp, Ax_total, 4*LB, MPI_DOUBLE, MPI_SUM) -- data.cpp:995
        GVB_NCCL(ncclAllReduce(out, out, (size_t)(4 * c->mbytes), ncclDouble, ncclSum, c->comm, c->stream));
    }
    return GVB_OK;
}
int gvb_atx_dev(gvb_ctx* c, const double* u, double* out) {
    GVB_ARG(c->have_stats, "compute_stats must run before ATx");
    prof_mark(c, 1);
#ifdef GVB_LEGACY_KERNELS
    int rc_mv = c->kernel_gen == 0 ? gvb_atx_simple(c, u, out) : (c->kernel_gen == 1 ? gvb_atx_lut(c, u, out) : gvb_atx_tile(c, u, out));
#else
    int rc_mv = gvb_atx_tile(c, u, out);
#endif
    prof_mark(c, 1);
    GVB_CHECK(rc_mv);
    c->sweeps++;
    return GVB_OK;
}

extern "C" int gvb_dAx(gvb_ctx* c, gvb_vec v, gvb_vec out) {
    GVB_ARG(c && v && out && v->cap >= c->Mg_pad * 4 && out->cap >= c->Npad, "Ax needs an M-vector and an N-vector from gvb_vec_alloc_M/_N");
    return gvb_ax_dev(c, v->d, out->d, true);
}
// out0 = X.v0, out1 = X.v1 from one pass over the bed; the two N-vectors are all-reduced like gvb_dAx's
int gvb_ax2_dev(gvb_ctx* c, const double* v0, const double* v1, double* out0, double* out1) {
    GVB_ARG(c->have_stats, "compute_stats must run before Ax");
#ifdef GVB_LEGACY_KERNELS
    if (c->kernel_gen < 2) {
        GVB_CHECK(gvb_ax_dev(c, v0, out0, true));
        return gvb_ax_dev(c, v1, out1, true);
    }
#endif
    prof_mark(c, 2);
    int rc_mv = gvb_ax_tile_dual(c, v0, v1, out0, out1);
    prof_mark(c, 2);
    GVB_CHECK(rc_mv);
    c->sweeps++;
    c->dual_sweeps++;
    if (c->nranks > 1) {
        GVB_NCCL(ncclGroupStart());
        GVB_NCCL(ncclAllReduce(out0, out0, (size_t)(4 * c->mbytes), ncclDouble, ncclSum, c->comm, c->stream));
        GVB_NCCL(ncclAllReduce(out1, out1, (size_t)(4 * c->mbytes), ncclDouble, ncclSum, c->comm, c->stream));
        GVB_NCCL(ncclGroupEnd());
    }
    return GVB_OK;
}
extern "C" int gvb_dAx2(gvb_ctx* c, gvb_vec v0, gvb_vec v1, gvb_vec out0, gvb_vec out1) {
    GVB_ARG(c && v0 && v1 && out0 && out1 && v0->cap >= c->Mg_pad * 4 && v1->cap >= c->Mg_pad * 4 && out0->cap >= c->Npad && out1->cap >= c->Npad,
            "Ax2 needs two M-vectors and two N-vectors from gvb_vec_alloc_M/_N");
    GVB_ARG(out0 != out1, "the two products need two output vectors");
    return gvb_ax2_dev(c, v0->d, v1->d, out0->d, out1->d);
}
// out0 = X^T.u0, out1 = X^T.u1 from one pass over the bed (local products: nothing to all-reduce)
int gvb_atx2_dev(gvb_ctx* c, const double* u0, const double* u1, double* out0, double* out1) {
    GVB_ARG(c->have_stats, "compute_stats must run before ATx");
#ifdef GVB_LEGACY_KERNELS
    if (c->kernel_gen < 2) {
        GVB_CHECK(gvb_atx_dev(c, u0, out0));
        return gvb_atx_dev(c, u1, out1);
    }
#endif
    if (c->total_missing > 0) {   // shards with missing genotypes: two single sweeps (gvb_atx_tile_dual)
        GVB_CHECK(gvb_atx_dev(c, u0, out0));
        return gvb_atx_dev(c, u1, out1);
    }
    prof_mark(c, 2);
    int rc_mv = gvb_atx_tile_dual(c, u0, u1, out0, out1);
    prof_mark(c, 2);
    GVB_CHECK(rc_mv);
    c->sweeps++;
    c->dual_sweeps++;
    return GVB_OK;
}
extern "C" int gvb_dATx2(gvb_ctx* c, gvb_vec u0, gvb_vec u1, gvb_vec out0, gvb_vec out1) {
    GVB_ARG(c && u0 && u1 && out0 && out1 && u0->cap >= c->Npad && u1->cap >= c->Npad && out0->cap >= c->Mg_pad * 4 && out1->cap >= c->Mg_pad * 4,
            "ATx2 needs two N-vectors and two M-vectors from gvb_vec_alloc_N/_M");
    GVB_ARG(out0 != out1, "the two products need two output vectors");
    return gvb_atx2_dev(c, u0->d, u1->d, out0->d, out1->d);
}
extern "C" int gvb_dATx(gvb_ctx* c, gvb_vec u, gvb_vec out) {
    GVB_ARG(c && u && out && u->cap >= c->Npad && out->cap >= c->Mg_pad * 4, "ATx needs an N-vector and an M-vector from gvb_vec_alloc_N/_M");
    return gvb_atx_dev(c, u->d, out->d);
}

__global__ void scale_slice_kernel(const double* __restrict__ in, long off, long n, double f, double* __restrict__ out) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) out[i] = in[off + i] * f;
}
__global__ void scale_inplace_kernel(double* __restrict__ x, long n, double f) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) x[i] *= f;
}

// data::Ax(double*, SB, LB): host in, host out.  A byte sub-range is the full product restricted to
// the individuals [4SB, 4SB+4LB) and rescaled by sqrt(N)/sqrt(4LB) (data.cpp:998-1005).
extern "C" int gvb_Ax(gvb_ctx* c, const double* v, double* out, long SB, long LB) {
    GVB_ARG(c && c->bed && v && out, "ctx / matrix / pointers");
    GVB_ARG(SB >= 0 && LB > 0 && SB + LB <= c->mbytes, "byte range");
    GVB_CUDA(cudaMemcpyAsync(c->tmpM, v, c->M * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    GVB_CHECK(gvb_ax_dev(c, c->tmpM, c->tmpN, true));
    bool full = (SB == 0 && LB == c->mbytes);
    if (full) {
        GVB_CUDA(cudaMemcpyAsync(out, c->tmpN, 4 * LB * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    } else {
        double f = sqrt((double)c->N) / sqrt((double)(4 * LB));
        scale_slice_kernel<<<(unsigned)std::min((4 * LB + 255) / 256, 1184l), 256, 0, c->stream>>>(c->tmpN, 4 * SB, 4 * LB, f, c->tmpN2);
        GVB_LAUNCHED(c);
        GVB_CUDA(cudaMemcpyAsync(out, c->tmpN2, 4 * LB * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    return GVB_OK;
}

// data::ATx(double*, SB, LB): u indexes individuals relative to 4*SB; everything outside the range
// (and any entry that would address an individual >= N) is treated as zero.
extern "C" int gvb_ATx(gvb_ctx* c, const double* u, double* out, long SB, long LB) {
    GVB_ARG(c && c->bed && u && out, "ctx / matrix / pointers");
    GVB_ARG(SB >= 0 && LB > 0 && SB + LB <= c->mbytes, "byte range");
    bool full = (SB == 0 && LB == c->mbytes);
    long n_in = std::min(4 * LB, c->N - 4 * SB);
    GVB_CUDA(cudaMemsetAsync(c->tmpN, 0, c->Npad * sizeof(double), c->stream));
    if (n_in > 0) GVB_CUDA(cudaMemcpyAsync(c->tmpN + 4 * SB, u, n_in * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    GVB_CHECK(gvb_atx_dev(c, c->tmpN, c->tmpM));
    if (!full) {
        double f = sqrt((double)c->N) / sqrt((double)(4 * LB));
        scale_inplace_kernel<<<(unsigned)std::min((c->M + 255) / 256, 1184l), 256, 0, c->stream>>>(c->tmpM, c->M, f);
        GVB_LAUNCHED(c);
    }
    GVB_CUDA(cudaMemcpyAsync(out, c->tmpM, c->M * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    return GVB_OK;
}

// the two accumulators of data::dot_product for one marker (data.cpp:758-779); a utility path: one block
__global__ void marker_dot_kernel(const uint32_t* __restrict__ bed, long Mg_pad, long mloc, const double* __restrict__ u, long p_lo, long p_hi,
                                  double* __restrict__ res) {
    __shared__ double sa[256], sb[256];
    double a = 0.0, b = 0.0;
    long g = mloc >> 2;
    int q = (int)(mloc & 3);
    for (long p = p_lo + threadIdx.x; p < p_hi; p += blockDim.x) {
        uint32_t w = bed[((p >> 5) * Mg_pad + g) * 32 + (p & 31)];
        unsigned byte = (w >> (8 * q)) & 0xFFu;
        for (int k = 0; k < 4; k++) {
            unsigned code = (byte >> (2 * k)) & 3u;
            double x = u[4 * p + k];
            a += (code == 0u ? 2.0 : (code == 2u ? 1.0 : 0.0)) * x;
            b += (code == 1u ? 0.0 : 1.0) * x;
        }
    }
    sa[threadIdx.x] = a;
    sb[threadIdx.x] = b;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) { sa[threadIdx.x] += sa[threadIdx.x + s]; sb[threadIdx.x] += sb[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { res[0] = sa[0]; res[1] = sb[0]; }
}

extern "C" int gvb_marker_dot(gvb_ctx* c, long mloc, const double* u, long SB, long LB, double* dpa, double* dpb) {
    GVB_ARG(c && c->bed && u && dpa && dpb, "ctx / matrix / pointers");
    GVB_ARG(mloc >= 0 && mloc < c->M && SB >= 0 && LB > 0 && SB + LB <= c->mbytes, "marker / byte range");
    long n_in = std::min(4 * LB, c->N - 4 * SB);
    GVB_CUDA(cudaMemsetAsync(c->tmpN, 0, c->Npad * sizeof(double), c->stream));
    if (n_in > 0) GVB_CUDA(cudaMemcpyAsync(c->tmpN + 4 * SB, u, n_in * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    marker_dot_kernel<<<1, 256, 0, c->stream>>>(c->bed, c->Mg_pad, mloc, c->tmpN, SB, SB + LB, c->red_result);
    GVB_LAUNCHED(c);
    GVB_CUDA(cudaMemcpyAsync(c->h_red, c->red_result, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    *dpa = c->h_red[0];
    *dpb = c->h_red[1];
    return GVB_OK;
}

extern "C" int gvb_allreduce_host(gvb_ctx* c, double* buf, int n) {
    GVB_ARG(c && buf && n >= 0 && n <= GVB_RED_MAXK, "buffer of at most GVB_RED_MAXK doubles");
    if (c->nranks <= 1 || n == 0) return GVB_OK;
    for (int i = 0; i < n; i++) c->h_red[i] = buf[i];
    GVB_CUDA(cudaMemcpyAsync(c->red_result, c->h_red, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    GVB_NCCL(ncclAllReduce(c->red_result, c->red_result, n, ncclDouble, ncclSum, c->comm, c->stream));
    GVB_CUDA(cudaMemcpyAsync(c->h_red, c->red_result, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GVB_CUDA(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; i++) buf[i] = c->h_red[i];
    return GVB_OK;
}
