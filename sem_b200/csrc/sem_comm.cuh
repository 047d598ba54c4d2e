// Multi-GPU plumbing: one process per GPU, NCCL over NVLink.  libnccl is resolved at run time with dlopen (the copy
// torch already loaded when the process uses torch.distributed; else the system library), so the library has no link
// dependency on NCCL and single-GPU use needs none.
//
// The interface exchange of the partitioned apply does not go through NCCL: every rank owns a mailbox in its own HBM
// (cudaMalloc + cudaIpc handle, opened by its two neighbours at set-up), a push kernel stores the interface lines straight
// into the neighbour's mailbox over NVLink and releases a per-line epoch flag there; the consumer kernel acquires the flag
// and adds.  Epoch counters live in device memory, so the sequence is CUDA-graph replayable.  NCCL send/recv remains as
// the fallback when peer mapping is unavailable (SEM_B200_NO_P2P=1 forces it for A/B runs).
#pragma once
#include "sem_common.cuh"

namespace semb {

struct Comm {
    void* nccl;        // ncclComm_t
    int rank, world;
    double* recv;      // device: [2 sides][max_fields][NY] receive staging for the interface lines (NCCL path)
    int max_fields;
    // peer-memory mailbox (p2p != 0)
    int p2p;
    size_t slot_len;                 // doubles per line slot (NY rounded up to 16)
    void* box;                       // this rank's mailbox: data [2 parities][2 sides][max_fields][slot_len] doubles, then
                                     // 64-bit words: arrived[2][max_fields] (written by the neighbours),
                                     // sent[2][max_fields], consumed[2][max_fields] (local epoch counters)
    void* peer_box[2];               // the left / right neighbour's mailbox mapped into this process (or null)
    // Second region of the same mailbox, used by the exchange that is fused into the operator kernel (XchArgs,
    // sem_march3_kernel): its own data slots [2 parities][2 sides][max_fields][slot_len] and per-STRIP flag words
    // arrived[2 sides][max_fields][slot_len], epoch[2 sides][max_fields][slot_len], so that its epochs never share a
    // parity slot with the per-line epochs of the stand-alone exchange kernels.
    size_t fused_off;                // byte offset of that region in a mailbox
    int loopback;                    // self-test on one GPU: this rank is its own left and right neighbour (no NCCL)
    int min_nex, max_nex;            // narrowest / widest slab of the partition (element columns), agreed at attach time: every
                                     // choice between exchange paths must come out the same on all ranks
};

int comm_unique_id(unsigned char out[128]);
int comm_init(Comm& c, const unsigned char id[128], int rank, int world, int NY, int nex);
// One-GPU self-test of the peer-memory paths: a communicator whose left and right neighbour are this rank itself, so the
// interface lines of a slab context (has_left / has_right) are exchanged with each other (line 0 <-> last line).
int comm_init_loopback(Comm& c, int NY, int nex);
// arguments of the in-kernel exchange for this rank's slab (peer-memory path only)
int comm_fill_xch(const Comm& c, const MeshDev& g, XchArgs& X);
void comm_destroy(Comm& c);
// in-place sum of k doubles over all ranks
int comm_allreduce_sum(const Comm& c, double* buf, int k, cudaStream_t st);
// recv[count_per_rank] = this rank's block of the element-wise sum over ranks of send[world * count_per_rank]
int comm_reduce_scatter_sum(const Comm& c, const double* send, double* recv, size_t count_per_rank, cudaStream_t st);
// recv[world * count_per_rank] = concatenation of every rank's send[count_per_rank]
int comm_allgather(const Comm& c, const double* send, double* recv, size_t count_per_rank, cudaStream_t st);
// Interface exchange after a local operator apply: every field's interface line(s) hold this rank's element sums;
// send them to the neighbour(s), receive theirs and add (two-term sum: bitwise identical on both ranks).
// Two halves so that the caller can run the interior of the operator between them (on another stream):
//   comm_exchange_transfer: ncclSend/ncclRecv of the interface lines into the receive staging;
//   comm_exchange_finish:   line = own + received (lower rank's term first).
int comm_exchange_transfer(const Comm& c, const MeshDev& g, double* const* fields, int nf, cudaStream_t st);
int comm_exchange_finish(const Comm& c, const MeshDev& g, double* const* fields, int nf, cudaStream_t st);
// both halves back to back: ONE kernel on the peer-memory path (push, wait, add per interface line)
int comm_exchange_fused(const Comm& c, const MeshDev& g, double* const* fields, int nf, cudaStream_t st);
inline int comm_exchange_add(const Comm& c, const MeshDev& g, double* const* fields, int nf, cudaStream_t st) {
    return comm_exchange_fused(c, g, fields, nf, st);
}

}  // namespace semb
