// Multi-GPU plumbing: one process per GPU, NCCL over NVLink.  libnccl is resolved at run time with dlopen (the copy
// torch already loaded when the process uses torch.distributed; else the system library), so the library has no link
// dependency on NCCL and single-GPU use needs none.
#pragma once
#include "sem_common.cuh"

namespace semb {

struct Comm {
    void* nccl;        // ncclComm_t
    int rank, world;
    double* recv;      // device: [2 sides][max_fields][NY] receive staging for the interface lines
    int max_fields;
};

int comm_unique_id(unsigned char out[128]);
int comm_init(Comm& c, const unsigned char id[128], int rank, int world, int NY);
void comm_destroy(Comm& c);
// in-place sum of k doubles over all ranks
int comm_allreduce_sum(const Comm& c, double* buf, int k, cudaStream_t st);
// Interface exchange after a local operator apply: every field's interface line(s) hold this rank's element sums;
// send them to the neighbour(s), receive theirs and add (two-term sum: bitwise identical on both ranks).
int comm_exchange_add(const Comm& c, const MeshDev& g, double* const* fields, int nf, cudaStream_t st);

}  // namespace semb
