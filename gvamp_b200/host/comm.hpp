// comm.hpp -- process-level view of the marker-sharded run: which shard (rank) this process owns and
// how the ranks found each other.  Stands where MPI_Init / MPI_Comm_rank / MPI_Comm_size stood in the
// reference (main_real.cpp:17-27): one process per B200, launched by gvamp_b200/bin/gvamp-launch, by
// torchrun (--no-python), or alone.
#pragma once
#include <string>

namespace gvb_host {

struct Comm {
    int rank = 0;
    int nranks = 1;
    int local_rank = 0;       // CUDA ordinal this process drives
    unsigned char nccl_id[128];
    bool have_id = false;
};

// Reads GVB_RANK / GVB_NRANKS / GVB_LOCAL_RANK, falling back to torchrun's RANK / WORLD_SIZE /
// LOCAL_RANK; obtains the NCCL unique id from GVB_NCCL_ID (hex, set by gvamp-launch) or through a
// rendezvous file written by rank 0 (single node only: its name derives from the launcher's pid).  Idempotent.
Comm& world();
void rendezvous_done();   // rank 0 removes the rendezvous file (called once the communicator exists, and at exit)
bool is_root();
double wtime();

}  // namespace gvb_host
