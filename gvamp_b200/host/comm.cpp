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
    // rendezvous through a file next to the launcher: rank 0 publishes, the others wait for a file
    // that is not older than this process
    std::string path = rendezvous_path();
    time_t started = time(nullptr);
    if (c.rank == 0) {
        if (gvb_nccl_unique_id(c.nccl_id) != GVB_OK) {
            std::cout << "FATAL: cannot create a NCCL unique id: " << gvb_last_error() << std::endl;
            exit(EXIT_FAILURE);
        }
        std::string tmp = path + ".tmp";
        int fd = open(tmp.c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0600);
        if (fd < 0 || write(fd, c.nccl_id, 128) != 128) {
            std::cout << "FATAL: cannot write the rendezvous file " << tmp << std::endl;
            exit(EXIT_FAILURE);
        }
        close(fd);
        rename(tmp.c_str(), path.c_str());
    } else {
        for (int tries = 0;; tries++) {
            struct stat st;
            if (stat(path.c_str(), &st) == 0 && st.st_size == 128 && st.st_mtime + 2 >= started) {
                int fd = open(path.c_str(), O_RDONLY);
                if (fd >= 0 && read(fd, c.nccl_id, 128) == 128) { close(fd); break; }
                if (fd >= 0) close(fd);
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
