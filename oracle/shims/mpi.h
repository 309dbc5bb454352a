/* oracle/shims/mpi.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Single-rank stand-in for <mpi.h> so the UNMODIFIED gVAMP reference sources under
 * /root/reference compile in an image that has no MPI.  Every collective is
 * the identity on one rank (Allreduce == memcpy), MPI-IO is pread/pwrite with the byte displacement
 * set by MPI_File_set_view, and point-to-point calls abort (never reached with one rank).
 * Only the ~45 MPI identifiers the reference actually uses are provided.
 */
#ifndef GVAMP_ORACLE_MPI_SHIM_H
#define GVAMP_ORACLE_MPI_SHIM_H

#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <time.h>
#include <unistd.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;
typedef int MPI_Request;
typedef long long MPI_Offset;

typedef struct MPI_Status {
    int MPI_SOURCE;
    int MPI_TAG;
    int MPI_ERROR;
    long long shim_count_bytes;
} MPI_Status;

typedef struct shim_mpi_file {
    int fd;
    MPI_Offset disp;   /* byte displacement of the current view */
    int etype_size;    /* size of the view's elementary type */
} *MPI_File;

#define MPI_COMM_WORLD 0
#define MPI_COMM_SELF 1
#define MPI_SUCCESS 0
#define MPI_INFO_NULL 0
#define MPI_THREAD_MULTIPLE 3

/* datatypes are encoded as their size in bytes plus a tag in the high bits */
#define MPI_UNSIGNED_CHAR ((MPI_Datatype)0x0101)
#define MPI_INT ((MPI_Datatype)0x0204)
#define MPI_DOUBLE ((MPI_Datatype)0x0308)
#define MPI_UNSIGNED_LONG_LONG ((MPI_Datatype)0x0408)

#define MPI_SUM 1
#define MPI_MAX 2

#define MPI_MODE_RDONLY 1
#define MPI_MODE_WRONLY 2
#define MPI_MODE_CREATE 4

#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)

static inline int shim_mpi_dtsize(MPI_Datatype dt) { return dt & 0xff; }

static inline int MPI_Init_thread(int* argc, char*** argv, int required, int* provided) {
    (void)argc; (void)argv;
    if (provided) *provided = required;
    return MPI_SUCCESS;
}
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int* rank) { (void)c; *rank = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int* size) { (void)c; *size = 1; return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; fflush(stdout); fflush(stderr); _exit(code ? code : 1); return 0; }
static inline double MPI_Wtime(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static inline int MPI_Type_size(MPI_Datatype dt, int* size) { *size = shim_mpi_dtsize(dt); return MPI_SUCCESS; }

static inline int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm c) {
    (void)op; (void)c;
    if (sendbuf != recvbuf) memmove(recvbuf, sendbuf, (size_t)count * (size_t)shim_mpi_dtsize(dt));
    return MPI_SUCCESS;
}
static inline int MPI_Iallreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm c, MPI_Request* req) {
    if (req) *req = 0;
    return MPI_Allreduce(sendbuf, recvbuf, count, dt, op, c);
}
static inline int MPI_Waitall(int n, MPI_Request* reqs, MPI_Status* st) { (void)n; (void)reqs; (void)st; return MPI_SUCCESS; }

static inline int MPI_Send(const void* buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm c) {
    (void)buf; (void)count; (void)dt; (void)dest; (void)tag; (void)c;
    fprintf(stderr, "mpi shim: MPI_Send called on a single-rank run\n");
    return MPI_Abort(c, 2);
}
static inline int MPI_Recv(void* buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm c, MPI_Status* st) {
    (void)buf; (void)count; (void)dt; (void)src; (void)tag; (void)st;
    fprintf(stderr, "mpi shim: MPI_Recv called on a single-rank run\n");
    return MPI_Abort(c, 2);
}
static inline int MPI_Get_count(const MPI_Status* st, MPI_Datatype dt, int* count) {
    *count = st ? (int)(st->shim_count_bytes / shim_mpi_dtsize(dt)) : 0;
    return MPI_SUCCESS;
}

static inline int MPI_File_open(MPI_Comm c, const char* path, int amode, MPI_Info info, MPI_File* fh) {
    (void)c; (void)info;
    int flags = 0;
    if ((amode & MPI_MODE_RDONLY) && !(amode & MPI_MODE_WRONLY)) flags = O_RDONLY;
    else if (amode & MPI_MODE_WRONLY) flags = O_WRONLY;
    if (amode & MPI_MODE_CREATE) flags |= O_CREAT;
    int fd = open(path, flags, 0644);
    if (fd < 0) { *fh = 0; return 1; }
    MPI_File f = (MPI_File)malloc(sizeof(struct shim_mpi_file));
    f->fd = fd; f->disp = 0; f->etype_size = 1;
    *fh = f;
    return MPI_SUCCESS;
}
static inline int MPI_File_close(MPI_File* fh) {
    if (fh && *fh) { close((*fh)->fd); free(*fh); *fh = 0; }
    return MPI_SUCCESS;
}
static inline int MPI_File_set_view(MPI_File fh, MPI_Offset disp, MPI_Datatype etype, MPI_Datatype filetype, const char* rep, MPI_Info info) {
    (void)filetype; (void)rep; (void)info;
    fh->disp = disp; fh->etype_size = shim_mpi_dtsize(etype);
    return MPI_SUCCESS;
}
static inline int shim_mpi_rw(MPI_File fh, MPI_Offset offset, void* buf, int count, MPI_Datatype dt, MPI_Status* st, int wr) {
    size_t total = (size_t)count * (size_t)shim_mpi_dtsize(dt), done = 0;
    off_t pos = (off_t)(fh->disp + offset * (MPI_Offset)fh->etype_size);
    while (done < total) {
        ssize_t r = wr ? pwrite(fh->fd, (const char*)buf + done, total - done, pos + (off_t)done)
                       : pread(fh->fd, (char*)buf + done, total - done, pos + (off_t)done);
        if (r <= 0) break;
        done += (size_t)r;
    }
    if (st) { st->MPI_SOURCE = 0; st->MPI_TAG = 0; st->MPI_ERROR = MPI_SUCCESS; st->shim_count_bytes = (long long)done; }
    return (done == total || !wr) ? MPI_SUCCESS : 1;
}
static inline int MPI_File_read_at(MPI_File fh, MPI_Offset off, void* buf, int count, MPI_Datatype dt, MPI_Status* st) { return shim_mpi_rw(fh, off, buf, count, dt, st, 0); }
static inline int MPI_File_read_at_all(MPI_File fh, MPI_Offset off, void* buf, int count, MPI_Datatype dt, MPI_Status* st) { return shim_mpi_rw(fh, off, buf, count, dt, st, 0); }
static inline int MPI_File_write_at(MPI_File fh, MPI_Offset off, const void* buf, int count, MPI_Datatype dt, MPI_Status* st) { return shim_mpi_rw(fh, off, (void*)buf, count, dt, st, 1); }
static inline int MPI_File_write_at_all(MPI_File fh, MPI_Offset off, const void* buf, int count, MPI_Datatype dt, MPI_Status* st) { return shim_mpi_rw(fh, off, (void*)buf, count, dt, st, 1); }

#endif
