#include "sem_comm.cuh"
#include "sem_tma.cuh"

#include <cstdlib>
#include <cstring>
#include <vector>
#include <dlfcn.h>
#include <nccl.h>

namespace semb {

namespace {
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.handle) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { set_error(std::string("cannot dlopen libnccl: ") + dlerror()); return -1; }
#define SEM_SYM(field, sym)                                                    \
    *(void**)(&g_nccl.field) = dlsym(h, sym);                                  \
    if (!g_nccl.field) { set_error(std::string("libnccl lacks ") + sym); return -1; }
    SEM_SYM(GetUniqueId, "ncclGetUniqueId")
    SEM_SYM(CommInitRank, "ncclCommInitRank")
    SEM_SYM(CommDestroy, "ncclCommDestroy")
    SEM_SYM(AllReduce, "ncclAllReduce")
    SEM_SYM(ReduceScatter, "ncclReduceScatter")
    SEM_SYM(AllGather, "ncclAllGather")
    SEM_SYM(Send, "ncclSend")
    SEM_SYM(Recv, "ncclRecv")
    SEM_SYM(GroupStart, "ncclGroupStart")
    SEM_SYM(GroupEnd, "ncclGroupEnd")
    SEM_SYM(GetErrorString, "ncclGetErrorString")
#undef SEM_SYM
    g_nccl.handle = h;
    return 0;
}
}  // namespace

#define SEM_NCCL(call)                                                                                       \
    do {                                                                                                     \
        ncclResult_t _r = (call);                                                                            \
        if (_r != ncclSuccess) {                                                                             \
            set_error(std::string(#call) + " failed: " + g_nccl.GetErrorString(_r));                         \
            return -4;                                                                                       \
        }                                                                                                    \
    } while (0)

static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");

int comm_unique_id(unsigned char out[128]) {
    if (load_nccl()) return -1;
    ncclUniqueId id;
    SEM_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out, &id, 128);
    return 0;
}

// ---- peer-memory mailboxes --------------------------------------------------------------------------------------------------
// Layout of a mailbox (see sem_comm.cuh).  Everything a kernel needs to find its slot is computed from these helpers.
static inline size_t box_data_doubles(const Comm& c) { return (size_t)2 * 2 * c.max_fields * c.slot_len; }
static inline double* box_slot(const Comm& c, void* box, int parity, int side, int f) {
    return reinterpret_cast<double*>(box) + ((size_t)(parity * 2 + side) * c.max_fields + f) * c.slot_len;
}
enum { BOX_ARRIVED = 0, BOX_SENT = 1, BOX_CONSUMED = 2 };
static inline unsigned long long* box_word(const Comm& c, void* box, int kind, int side, int f) {
    unsigned long long* w = reinterpret_cast<unsigned long long*>(reinterpret_cast<double*>(box) + box_data_doubles(c));
    return w + ((size_t)kind * 2 + side) * c.max_fields + f;
}

// total bytes of a mailbox; fixes slot_len and the offset of the fused-exchange region
static size_t box_layout(Comm& c, int NY) {
    c.slot_len = ((size_t)NY + 15) & ~(size_t)15;
    const size_t legacy = box_data_doubles(c) * sizeof(double) + (size_t)3 * 2 * c.max_fields * sizeof(unsigned long long);
    c.fused_off = (legacy + 127) & ~(size_t)127;
    const size_t fused = box_data_doubles(c) * sizeof(double) + (size_t)2 * 2 * c.max_fields * c.slot_len * sizeof(unsigned long long);
    return c.fused_off + fused;
}
static inline double* fused_slot(const Comm& c, void* box, int side) {   // parity 0, field 0
    return reinterpret_cast<double*>(reinterpret_cast<char*>(box) + c.fused_off) + (size_t)side * c.max_fields * c.slot_len;
}
static inline unsigned long long* fused_word(const Comm& c, void* box, int kind, int side) {   // kind 0: arrived, 1: epoch; field 0
    unsigned long long* w = reinterpret_cast<unsigned long long*>(reinterpret_cast<double*>(reinterpret_cast<char*>(box) + c.fused_off) +
                                                                  box_data_doubles(c));
    return w + ((size_t)kind * 2 + side) * c.max_fields * c.slot_len;
}

int comm_fill_xch(const Comm& c, const MeshDev& g, XchArgs& X) {
    std::memset(&X, 0, sizeof(X));
    if (!c.p2p) { set_error("comm_fill_xch: no peer-memory mailboxes"); return -2; }
    X.mask = (g.has_left ? 1 : 0) | (g.has_right ? 2 : 0);
    X.slot_len = c.slot_len;
    X.parity_stride = (unsigned long long)2 * c.max_fields * c.slot_len;
    X.flag_stride = c.slot_len;
    for (int s = 0; s < 2; ++s) {
        if (!(X.mask & (1 << s))) continue;
        // my line 0 is the left neighbour's last line (its side 1); my last line is the right neighbour's line 0 (its side 0)
        X.peer_slot[s] = fused_slot(c, c.peer_box[s], 1 - s);
        X.peer_flag[s] = fused_word(c, c.peer_box[s], 0, 1 - s);
        X.my_slot[s] = fused_slot(c, c.box, s);
        X.my_flag[s] = fused_word(c, c.box, 0, s);
        X.epoch[s] = fused_word(c, c.box, 1, s);
    }
    return 0;
}

int comm_init_loopback(Comm& c, int NY, int nex) {
    std::memset(&c, 0, sizeof(c));
    c.min_nex = c.max_nex = nex;
    c.rank = 0;
    c.world = 1;
    c.max_fields = 4;
    c.loopback = 1;
    const size_t bytes = box_layout(c, NY);
    SEM_CUDA(cudaMalloc(&c.box, bytes));
    SEM_CUDA(cudaMemset(c.box, 0, bytes));
    c.peer_box[0] = c.peer_box[1] = c.box;
    c.p2p = 1;
    return 0;
}

// Allocate the mailbox, spread its IPC handle (ncclAllGather of the 64 handle bytes), map the neighbours' mailboxes.
// All ranks agree on the outcome (all-reduce of a success count): either every rank pushes or every rank uses NCCL.
static int p2p_init(Comm& c, int NY) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
    const bool off = std::getenv("SEM_B200_NO_P2P") != nullptr;
    ncclComm_t comm = (ncclComm_t)c.nccl;
    const size_t bytes = box_layout(c, NY);
    int ok = off ? 0 : 1;
    cudaIpcMemHandle_t mine;
    std::memset(&mine, 0, sizeof(mine));
    if (ok && cudaMalloc(&c.box, bytes) != cudaSuccess) { cudaGetLastError(); c.box = nullptr; ok = 0; }
    // a LOCAL failure must not return before the collectives below (the other ranks would block in them): it only clears `ok`,
    // and the agreement all-reduce then sends every rank to the NCCL fallback together
    if (ok && cudaMemset(c.box, 0, bytes) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    if (ok && cudaIpcGetMemHandle(&mine, c.box) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    // handles of all ranks (64 bytes each) + one double per rank for the agreement
    unsigned char* dh = nullptr;
    double* dok = nullptr;
    SEM_CUDA(cudaMalloc(&dh, (size_t)64 * c.world));
    SEM_CUDA(cudaMalloc(&dok, sizeof(double)));
    SEM_CUDA(cudaMemcpy(dh + (size_t)64 * c.rank, &mine, 64, cudaMemcpyHostToDevice));
    SEM_NCCL(g_nccl.AllGather(dh + (size_t)64 * c.rank, dh, 64, ncclChar, comm, 0));
    SEM_CUDA(cudaStreamSynchronize(0));
    std::vector<cudaIpcMemHandle_t> all(c.world);
    SEM_CUDA(cudaMemcpy(all.data(), dh, (size_t)64 * c.world, cudaMemcpyDeviceToHost));
    if (ok) {
        for (int s = 0; s < 2 && ok; ++s) {
            const int nb = c.rank + (s == 0 ? -1 : 1);
            if (nb < 0 || nb >= c.world) continue;
            if (cudaIpcOpenMemHandle(&c.peer_box[s], all[nb], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                c.peer_box[s] = nullptr;
                ok = 0;
            }
        }
    }
    const double okd = ok ? 1.0 : 0.0;
    SEM_CUDA(cudaMemcpy(dok, &okd, sizeof(double), cudaMemcpyHostToDevice));
    SEM_NCCL(g_nccl.AllReduce(dok, dok, 1, ncclDouble, ncclSum, comm, 0));
    SEM_CUDA(cudaStreamSynchronize(0));
    double sum = 0.0;
    SEM_CUDA(cudaMemcpy(&sum, dok, sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(dh);
    cudaFree(dok);
    c.p2p = (sum > c.world - 0.5) ? 1 : 0;
    if (!c.p2p) {
        for (int s = 0; s < 2; ++s)
            if (c.peer_box[s]) { cudaIpcCloseMemHandle(c.peer_box[s]); c.peer_box[s] = nullptr; }
        if (c.box) { cudaFree(c.box); c.box = nullptr; }
        if (std::getenv("SEM_B200_REQUIRE_P2P")) { set_error("peer-memory mailboxes unavailable (SEM_B200_REQUIRE_P2P set)"); return -5; }
    }
    return 0;
}

int comm_init(Comm& c, const unsigned char idb[128], int rank, int world, int NY, int nex) {
    if (load_nccl()) return -1;
    ncclUniqueId id;
    memcpy(&id, idb, 128);
    ncclComm_t comm;
    SEM_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
    c.nccl = comm;
    c.rank = rank;
    c.world = world;
    c.max_fields = 4;
    SEM_CUDA(cudaMalloc(&c.recv, sizeof(double) * 2 * c.max_fields * (size_t)NY));
    c.p2p = 0;
    c.box = nullptr;
    c.peer_box[0] = c.peer_box[1] = nullptr;
    c.loopback = 0;
    {   // narrowest and widest slab: max over ranks of (-nex, nex)
        double h[2] = {-(double)nex, (double)nex}, *d = nullptr;
        SEM_CUDA(cudaMalloc(&d, sizeof(h)));
        SEM_CUDA(cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice));
        SEM_NCCL(g_nccl.AllReduce(d, d, 2, ncclDouble, ncclMax, comm, 0));
        SEM_CUDA(cudaStreamSynchronize(0));
        SEM_CUDA(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(d);
        c.min_nex = (int)(-h[0] + 0.5);
        c.max_nex = (int)(h[1] + 0.5);
    }
    return p2p_init(c, NY);
}

void comm_destroy(Comm& c) {
    for (int s = 0; s < 2; ++s) {
        if (c.peer_box[s] && !c.loopback) cudaIpcCloseMemHandle(c.peer_box[s]);
        c.peer_box[s] = nullptr;
    }
    if (c.nccl && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)c.nccl);
    if (c.recv) cudaFree(c.recv);
    if (c.box) cudaFree(c.box);
    c.nccl = nullptr;
    c.recv = nullptr;
    c.box = nullptr;
    c.p2p = 0;
}

int comm_allreduce_sum(const Comm& c, double* buf, int k, cudaStream_t st) {
    if (c.loopback) return 0;   // one rank: the sum is the buffer itself
    SEM_NCCL(g_nccl.AllReduce(buf, buf, (size_t)k, ncclDouble, ncclSum, (ncclComm_t)c.nccl, st));
    return 0;
}

int comm_reduce_scatter_sum(const Comm& c, const double* send, double* recv, size_t count_per_rank, cudaStream_t st) {
    if (c.loopback) { SEM_CUDA(cudaMemcpyAsync(recv, send, count_per_rank * sizeof(double), cudaMemcpyDeviceToDevice, st)); return 0; }
    SEM_NCCL(g_nccl.ReduceScatter(send, recv, count_per_rank, ncclDouble, ncclSum, (ncclComm_t)c.nccl, st));
    return 0;
}

int comm_allgather(const Comm& c, const double* send, double* recv, size_t count_per_rank, cudaStream_t st) {
    if (c.loopback) { if (recv != send) SEM_CUDA(cudaMemcpyAsync(recv, send, count_per_rank * sizeof(double), cudaMemcpyDeviceToDevice, st)); return 0; }
    SEM_NCCL(g_nccl.AllGather(send, recv, count_per_rank, ncclDouble, (ncclComm_t)c.nccl, st));
    return 0;
}

// line[f][iy] = a + b with the LOWER rank's partial sum first (a two-term IEEE sum is order independent, but the order
// is fixed anyway so the code reads as what it guarantees)

struct LinePtrs {
    double* line[8];
    const double* recv[8];
    int lower_first[8];
    int n;
};

__global__ void k_halo_add(const LinePtrs p, int NY) {
    const int iy = blockIdx.x * blockDim.x + threadIdx.x;
    const int q = blockIdx.y;
    if (iy >= NY || q >= p.n) return;
    const double mine = p.line[q][iy], other = p.recv[q][iy];
    p.line[q][iy] = p.lower_first[q] ? (mine + other) : (other + mine);
}

// ---- peer-memory exchange ----------------------------------------------------------------------------------------------------
struct PushArgs {
    const double* src[8];            // this rank's interface line
    double* dst[8];                  // parity-0 slot of that line in the neighbour's mailbox
    unsigned long long* arrived[8];  // the neighbour's epoch flag of the slot (remote)
    unsigned long long* sent[8];     // local epoch counter of the slot
    size_t parity_stride;            // doubles between the two parities of a slot
    int n;
};
struct WaitAddArgs {
    double* line[8];
    const double* slot[8];           // parity-0 slot in this rank's mailbox
    unsigned long long* arrived[8];  // local flag, written by the neighbour
    unsigned long long* consumed[8]; // local epoch counter
    int lower_first[8];
    size_t parity_stride;
    int n;
};


// one CTA per interface line: store it into the neighbour's mailbox (NVLink stores), then release the epoch flag there
__global__ void __launch_bounds__(1024) k_halo_push(const PushArgs p, int NY) {
    const int q = blockIdx.x;
    const unsigned long long e = *p.sent[q] + 1ull;        // every thread reads it before thread 0 advances it (barrier below)
    double* dst = p.dst[q] + (size_t)(e & 1ull) * p.parity_stride;
    const double* src = p.src[q];
    for (int i = threadIdx.x; i < NY; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        st_release_sys(p.arrived[q], e);
        *p.sent[q] = e;
    }
}

// one CTA per interface line: wait for the neighbour's epoch, line = own + received (lower rank's term first)
__global__ void __launch_bounds__(1024) k_halo_wait_add(const WaitAddArgs p, int NY) {
    const int q = blockIdx.x;
    __shared__ unsigned long long s_e;
    if (threadIdx.x == 0) {
        const unsigned long long e = *p.consumed[q] + 1ull;
        unsigned long long t0 = 0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (ld_acquire_sys(p.arrived[q]) < e) {
            __nanosleep(64);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 120ull * 1000000000ull) __trap();   // a neighbour died: fail loudly instead of hanging the box
        }
        s_e = e;
    }
    __syncthreads();
    const unsigned long long e = s_e;
    const double* recv = p.slot[q] + (size_t)(e & 1ull) * p.parity_stride;
    double* line = p.line[q];
    const bool lf = p.lower_first[q] != 0;
    for (int i = threadIdx.x; i < NY; i += blockDim.x) {
        const double mine = line[i], other = __ldcg(recv + i);   // the mailbox is written by a peer: read it from L2
        line[i] = lf ? (mine + other) : (other + mine);
    }
    __syncthreads();
    if (threadIdx.x == 0) *p.consumed[q] = e;
}

// Both halves in one launch (one CTA per interface line): push the line, then wait for the neighbour's copy of the same
// line and add it.  Used wherever nothing has to run between the two halves.
struct ExchangeArgs {
    PushArgs push;
    WaitAddArgs add;
};

__global__ void __launch_bounds__(1024) k_halo_exchange(const ExchangeArgs a, int NY) {
    const int q = blockIdx.x;
    __shared__ unsigned long long s_e;
    {   // ---- push (reads the line before the add below changes it)
        const PushArgs& p = a.push;
        const unsigned long long e = *p.sent[q] + 1ull;
        double* dst = p.dst[q] + (size_t)(e & 1ull) * p.parity_stride;
        const double* src = p.src[q];
        for (int i = threadIdx.x; i < NY; i += blockDim.x) dst[i] = src[i];
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            st_release_sys(p.arrived[q], e);
            *p.sent[q] = e;
        }
    }
    {   // ---- wait + add
        const WaitAddArgs& p = a.add;
        if (threadIdx.x == 0) {
            const unsigned long long e = *p.consumed[q] + 1ull;
            unsigned long long t0 = 0, t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            while (ld_acquire_sys(p.arrived[q]) < e) {
                __nanosleep(64);
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 120ull * 1000000000ull) __trap();
            }
            s_e = e;
        }
        __syncthreads();
        const unsigned long long e = s_e;
        const double* recv = p.slot[q] + (size_t)(e & 1ull) * p.parity_stride;
        double* line = p.line[q];
        const bool lf = p.lower_first[q] != 0;
        for (int i = threadIdx.x; i < NY; i += blockDim.x) {
            const double mine = line[i], other = __ldcg(recv + i);
            line[i] = lf ? (mine + other) : (other + mine);
        }
        __syncthreads();
        if (threadIdx.x == 0) *p.consumed[q] = e;
    }
}

static void fill_push(const Comm& c, const MeshDev& g, double* const* fields, int nf, PushArgs& a);
static void fill_wait_add(const Comm& c, const MeshDev& g, double* const* fields, int nf, WaitAddArgs& a);

int comm_exchange_fused(const Comm& c, const MeshDev& g, double* const* fields, int nf, cudaStream_t st) {
    if (nf > c.max_fields) { set_error("comm_exchange: too many fields"); return -2; }
    if (!g.has_left && !g.has_right) return 0;
    if (!c.p2p) {
        if (comm_exchange_transfer(c, g, fields, nf, st)) return -1;
        return comm_exchange_finish(c, g, fields, nf, st);
    }
    ExchangeArgs a;
    fill_push(c, g, fields, nf, a.push);
    fill_wait_add(c, g, fields, nf, a.add);   // same (field, side) order as fill_push: line q is sent and completed by CTA q
    k_halo_exchange<<<a.push.n, 1024, 0, st>>>(a, g.NY);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

static void fill_push(const Comm& c, const MeshDev& g, double* const* fields, int nf, PushArgs& a) {
    const size_t last = (size_t)(g.NX - 1) * g.LD;
    a.n = 0;
    a.parity_stride = (size_t)2 * c.max_fields * c.slot_len;
    for (int f = 0; f < nf; ++f) {
        if (g.has_left) {   // my left line goes to the left neighbour's "from the right" slot (side 1)
            a.src[a.n] = fields[f];
            a.dst[a.n] = box_slot(c, c.peer_box[0], 0, 1, f);
            a.arrived[a.n] = box_word(c, c.peer_box[0], BOX_ARRIVED, 1, f);
            a.sent[a.n] = box_word(c, c.box, BOX_SENT, 0, f);
            a.n++;
        }
        if (g.has_right) {  // my right line goes to the right neighbour's "from the left" slot (side 0)
            a.src[a.n] = fields[f] + last;
            a.dst[a.n] = box_slot(c, c.peer_box[1], 0, 0, f);
            a.arrived[a.n] = box_word(c, c.peer_box[1], BOX_ARRIVED, 0, f);
            a.sent[a.n] = box_word(c, c.box, BOX_SENT, 1, f);
            a.n++;
        }
    }
}

static void fill_wait_add(const Comm& c, const MeshDev& g, double* const* fields, int nf, WaitAddArgs& a) {
    const size_t last = (size_t)(g.NX - 1) * g.LD;
    a.n = 0;
    a.parity_stride = (size_t)2 * c.max_fields * c.slot_len;
    for (int f = 0; f < nf; ++f) {
        if (g.has_left) {
            a.line[a.n] = fields[f];
            a.slot[a.n] = box_slot(c, c.box, 0, 0, f);
            a.arrived[a.n] = box_word(c, c.box, BOX_ARRIVED, 0, f);
            a.consumed[a.n] = box_word(c, c.box, BOX_CONSUMED, 0, f);
            a.lower_first[a.n] = 0;   // the neighbour is the lower rank: its partial goes first
            a.n++;
        }
        if (g.has_right) {
            a.line[a.n] = fields[f] + last;
            a.slot[a.n] = box_slot(c, c.box, 0, 1, f);
            a.arrived[a.n] = box_word(c, c.box, BOX_ARRIVED, 1, f);
            a.consumed[a.n] = box_word(c, c.box, BOX_CONSUMED, 1, f);
            a.lower_first[a.n] = 1;
            a.n++;
        }
    }
}

static int p2p_push(const Comm& c, const MeshDev& g, double* const* fields, int nf, cudaStream_t st) {
    PushArgs a;
    fill_push(c, g, fields, nf, a);
    k_halo_push<<<a.n, 1024, 0, st>>>(a, g.NY);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

static int p2p_wait_add(const Comm& c, const MeshDev& g, double* const* fields, int nf, cudaStream_t st) {
    WaitAddArgs a;
    fill_wait_add(c, g, fields, nf, a);
    k_halo_wait_add<<<a.n, 1024, 0, st>>>(a, g.NY);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

int comm_exchange_transfer(const Comm& c, const MeshDev& g, double* const* fields, int nf, cudaStream_t st) {
    if (c.p2p) {
        if (nf > c.max_fields) { set_error("comm_exchange: too many fields"); return -2; }
        if (!g.has_left && !g.has_right) return 0;
        return p2p_push(c, g, fields, nf, st);
    }
#ifdef SEM_DEBUG_HOOKS   // timing experiments only (wrong results): never part of the shipped library
    static const bool skip = std::getenv("SEM_B200_DEBUG_NO_TRANSFER") != nullptr;
    if (skip) return 0;
#endif
    if (nf > c.max_fields) { set_error("comm_exchange: too many fields"); return -2; }
    if (!g.has_left && !g.has_right) return 0;
    ncclComm_t comm = (ncclComm_t)c.nccl;
    const size_t NY = (size_t)g.NY;
    const size_t last = (size_t)(g.NX - 1) * g.LD;
    SEM_NCCL(g_nccl.GroupStart());
    for (int f = 0; f < nf; ++f) {
        if (g.has_left) {
            SEM_NCCL(g_nccl.Send(fields[f], NY, ncclDouble, c.rank - 1, comm, st));
            SEM_NCCL(g_nccl.Recv(c.recv + (size_t)(0 * c.max_fields + f) * NY, NY, ncclDouble, c.rank - 1, comm, st));
        }
        if (g.has_right) {
            SEM_NCCL(g_nccl.Send(fields[f] + last, NY, ncclDouble, c.rank + 1, comm, st));
            SEM_NCCL(g_nccl.Recv(c.recv + (size_t)(1 * c.max_fields + f) * NY, NY, ncclDouble, c.rank + 1, comm, st));
        }
    }
    SEM_NCCL(g_nccl.GroupEnd());
    return 0;
}

int comm_exchange_finish(const Comm& c, const MeshDev& g, double* const* fields, int nf, cudaStream_t st) {
    if (nf > c.max_fields) { set_error("comm_exchange: too many fields"); return -2; }
    if (!g.has_left && !g.has_right) return 0;
    if (c.p2p) return p2p_wait_add(c, g, fields, nf, st);
    const size_t NY = (size_t)g.NY;
    const size_t last = (size_t)(g.NX - 1) * g.LD;
    LinePtrs lp;
    lp.n = 0;
    for (int f = 0; f < nf; ++f) {
        if (g.has_left) {
            lp.line[lp.n] = fields[f];
            lp.recv[lp.n] = c.recv + (size_t)(0 * c.max_fields + f) * NY;
            lp.lower_first[lp.n] = 0;   // the neighbour is the lower rank: its partial goes first
            lp.n++;
        }
        if (g.has_right) {
            lp.line[lp.n] = fields[f] + last;
            lp.recv[lp.n] = c.recv + (size_t)(1 * c.max_fields + f) * NY;
            lp.lower_first[lp.n] = 1;
            lp.n++;
        }
    }
    dim3 grid((unsigned)((g.NY + 255) / 256), (unsigned)lp.n);
    k_halo_add<<<grid, 256, 0, st>>>(lp, g.NY);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace semb
