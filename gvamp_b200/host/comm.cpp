#include "comm.hpp"

#include <fcntl.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "gvamp_b200.h"

namespace gvb_host {

static int env_int(const char* a, const char* b, int dflt) {
    const char* v = getenv(a);
    if (!v && b) v = getenv(b);
    return v ? atoi(v) : dflt;
}

static bool id_from_hex(const char* hex, unsigned char* out) {
    if (!hex || strlen(hex) != 256) return false;
    for (int i = 0; i < 128; i++) {
        unsigned v;
        if (sscanf(hex + 2 * i, "%2x", &v) != 1) return false;
        out[i] = (unsigned char)v;
    }
    return true;
}

static std::string rendezvous_path() {
    const char* dir = getenv("GVB_RDZV_DIR");
    const char* port = getenv("MASTER_PORT");
    const char* run = getenv("TORCHELASTIC_RUN_ID");
    char buf[512];
    snprintf(buf, sizeof(buf), "%s/gvb_nccl_id_%d_%s_%s", dir ? dir : "/tmp", (int)getppid(), port ? port : "0", run ? run : "none");
    return buf;
}

static std::string g_rdzv_file;

// rank 0: the id has been consumed by every rank once the NCCL communicator exists (ncclCommInitRank returns on all ranks together)
void rendezvous_done() {
    if (!g_rdzv_file.empty()) {
        unlink(g_rdzv_file.c_str());
        g_rdzv_file.clear();
    }
}

Comm& world() {
    static Comm c;
    static bool init = false;
    if (init) return c;
    init = true;
    c.rank = env_int("GVB_RANK", "RANK", 0);
    c.nranks = env_int("GVB_NRANKS", "WORLD_SIZE", 1);
    c.local_rank = env_int("GVB_LOCAL_RANK", "LOCAL_RANK", c.rank);
    if (c.nranks <= 1) { c.rank = 0; c.nranks = 1; return c; }
    if (id_from_hex(getenv("GVB_NCCL_ID"), c.nccl_id)) { c.have_id = true; return c; }
    // Rendezvous through a file (SINGLE NODE: the name is built from the launcher's pid and the master port; a multi-node launch passes
    // GVB_NCCL_ID instead).  The payload is the 128-byte id followed by rank 0's wall-clock start time, so that a file left behind by
    // an earlier run under the same launcher can never be taken for this run's; rank 0 replaces any such file (unlink, then O_EXCL |
    // O_NOFOLLOW: no symlink in a shared /tmp is followed) and removes its own once the communicator exists (rendezvous_done()).
    std::string path = rendezvous_path();
    struct timespec now;
    clock_gettime(CLOCK_REALTIME, &now);
    const long long started = (long long)now.tv_sec;
    if (c.rank == 0) {
        if (gvb_nccl_unique_id(c.nccl_id) != GVB_OK) {
            std::cout << "FATAL: cannot create a NCCL unique id: " << gvb_last_error() << std::endl;
            exit(EXIT_FAILURE);
        }
        std::string tmp = path + ".tmp";
        unlink(tmp.c_str());
        unlink(path.c_str());
        int fd = open(tmp.c_str(), O_CREAT | O_EXCL | O_NOFOLLOW | O_WRONLY, 0600);
        if (fd < 0 || write(fd, c.nccl_id, 128) != 128 || write(fd, &started, sizeof(started)) != (ssize_t)sizeof(started)) {
            std::cout << "FATAL: cannot write the rendezvous file " << tmp << std::endl;
            exit(EXIT_FAILURE);
        }
        close(fd);
        rename(tmp.c_str(), path.c_str());
        g_rdzv_file = path;
        atexit(rendezvous_done);
    } else {
        for (int tries = 0;; tries++) {
            int fd = open(path.c_str(), O_RDONLY | O_NOFOLLOW);
            if (fd >= 0) {
                long long stamp = 0;
                const bool ok = read(fd, c.nccl_id, 128) == 128 && read(fd, &stamp, sizeof(stamp)) == (ssize_t)sizeof(stamp);
                close(fd);
                // the ranks of one launch start within seconds of each other; an id published long before this process started
                // belongs to an earlier run whose rank 0 did not get to remove it
                if (ok && stamp + 30 >= started) break;
            }
            if (tries > 6000) {
                std::cout << "FATAL: rank " << c.rank << " timed out waiting for " << path << std::endl;
                exit(EXIT_FAILURE);
            }
            usleep(10000);
        }
    }
    c.have_id = true;
    return c;
}

bool is_root() { return world().rank == 0; }

double wtime() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

}  // namespace gvb_host
