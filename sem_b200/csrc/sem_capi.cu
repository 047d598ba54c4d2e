// C ABI of sem_b200 (see include/sem_b200.h): context, host<->device packing, fused operators, Krylov solvers.
#include "../../include/sem_b200.h"
#include "sem_aux.cuh"
#include "sem_comm.cuh"
#include "sem_dispatch.h"
#include "sem_march.cuh"
#include "sem_march3.cuh"
#include <cstdlib>
#include <cublas_v2.h>

#include <cmath>
#include <algorithm>
#include <cstring>
#include <functional>
#include <mutex>
#include <vector>

namespace semb {
static thread_local std::string g_err;
void set_error(const std::string& s) { g_err = s; }
}  // namespace semb

using namespace semb;

typedef int (*march_fn)(int, const MeshDev&, const MarchArgs&, const MarchGeom&, cudaStream_t);
typedef size_t (*smem_fn)(int, int);
typedef int (*upload_fn)(const double*, const double*, const double*);

#define SEM_TAB_ENTRY(P) {march_launch_p##P, march_smem_p##P, upload_tab_p##P, march3_launch_p##P, upload_tab3_p##P, march3_smem_p##P},
static const struct {
    march_fn launch;
    smem_fn smem;
    upload_fn upload;
    march_fn launch3;
    upload_fn upload3;
    size_t (*smem3)(int);
} g_orders[SEM_MAX_P] = {SEM_FOR_EACH_P(SEM_TAB_ENTRY)};

#define SEM_HOST_SEGMENTS 32  // most element-column segments of the host-buffer pipeline (default 16, SEM_B200_HOST_SEGMENTS)
#define SEM_GMRES_LAG 6       // Arnoldi steps enqueued ahead of the host-side Givens / convergence test

struct sem_ctx {
    MeshDev g;
    int device, sm_count, smem_optin, smem_sm;
    int Ty_req, Mx_req;
    int pin_gx, pin_iy;
    double *dD, *dKs, *dw;   // plain device tables
    double* dKdiag;          // diag(K), built on first use
    double* dStage;          // dense [NX][NY] staging buffer of the host<->device boundary, built on first use
    RedScratch rs;
    double* d_small;         // device staging for reduction results
    double* h_small;         // pinned mirror
    int small_len;
    double* h_ring;          // pinned: SEM_GMRES_LAG slots for the Hessenberg columns in flight
    size_t ring_stride;
    cudaEvent_t ev_ring[SEM_GMRES_LAG];
    Comm comm;               // NCCL communicator of the element-column partition (has_comm)
    int has_comm;
    cudaStream_t s_side;     // low-priority stream: the interior of an operator runs here while the interface lines travel
    cudaStream_t s_side2;    // second side stream: right edge and interior run concurrently on small slabs
    cudaEvent_t ev_in, ev_edge, ev_side;
    // host-buffer pipeline (sem_cd_jvp_host): upload / download streams, second staging buffer, per-segment events
    cudaStream_t s_h2d, s_d2h;
    cudaStream_t s_main;     // partitioned applies run (and are captured) here, ordered against the caller's stream by events
    cudaEvent_t ev_g0, ev_g1;
    struct GraphEntry { unsigned char key[sizeof(MarchArgs) + 8 * sizeof(double*) + 16]; int uses; cudaGraphExec_t exec; } gcache[8];
    int gcache_next;
    cudaStream_t s_solve;    // the Krylov solvers run here: CUDA graphs cannot be captured on the legacy default stream
    double* dStageOut;
    cudaEvent_t ev_up[SEM_HOST_SEGMENTS], ev_done[SEM_HOST_SEGMENTS], ev_start, ev_end;
    int streams_ready;
    // fast-diagonalisation preconditioner (sem_ctx_set_fdm): 1-D generalised eigenpairs of the x and y pencils
    cublasHandle_t blas;
    void* blas_ws;
    double *fQx, *fLx, *fQy, *fLy, *fT1, *fT2;
    double *fB1, *fB2;        // two-field transform buffers of the batched apply (allocated on first use)
    // boundary block of the NS pressure rows (sem_ctx_set_pbb, experimental): offsets of the boundary pressure nodes, the
    // dense inverse of K restricted to them (row-major nb x nb), gathered right-hand side and solution
    long long* pbb_idx;
    double *pbb_inv, *pbb_rhs, *pbb_z;
    int pbb_n;
    int fdm_ready, fdm_dir[4];
    int fdm_nxg, fdm_block;  // partitioned mesh: global line count and spectral modes per rank (exact distributed FDM)
    TabDev tab() const { return TabDev{dD, dKs, dw}; }
};

#define SEM_CHECK_CTX(ctx)                              \
    do {                                                \
        if (!(ctx)) {                                   \
            set_error("null context");                  \
            return -2;                                  \
        }                                               \
        SEM_CUDA(cudaSetDevice((ctx)->device));         \
    } while (0)

static const int SEM_MAX_RESTART = 4096;

extern "C" const char* sem_last_error(void) { return g_err.c_str(); }
extern "C" int sem_version(void) { return 100; }

extern "C" int sem_ctx_create(sem_ctx** out, const sem_mesh_desc* d) {
    if (!out || !d) { set_error("sem_ctx_create: null argument"); return -2; }
    if (d->P < 1 || d->P > SEM_MAX_P) { set_error("sem_ctx_create: P must be in 1..16"); return -2; }
    if (d->N_ex < 1 || d->N_ey < 1 || d->m_begin < 0 || d->m_end > d->N_ex || d->m_begin >= d->m_end) {
        set_error("sem_ctx_create: bad element counts / partition");
        return -2;
    }
    SEM_CUDA(cudaSetDevice(d->device));
    sem_ctx* c = new sem_ctx();
    std::memset(c, 0, sizeof(*c));
    c->device = d->device;
    SEM_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, d->device));
    SEM_CUDA(cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, d->device));
    SEM_CUDA(cudaDeviceGetAttribute(&c->smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, d->device));
    MeshDev& g = c->g;
    g.P = d->P;
    g.nex = d->m_end - d->m_begin;
    g.ney = d->N_ey;
    g.NX = g.nex * g.P + 1;
    g.NY = g.ney * g.P + 1;
    g.LD = round_up(g.NY, 16);
    g.gx0 = d->m_begin * g.P;
    g.NXg = d->N_ex * g.P + 1;
    g.has_left = d->m_begin > 0;
    g.has_right = d->m_end < d->N_ex;
    g.dx = d->dx;
    g.dy = d->dy;
    if ((long long)g.NX * g.LD >= (1ll << 31)) { set_error("sem_ctx_create: slab too large for 32-bit offsets"); delete c; return -2; }
    const long long Nglob = (long long)g.NXg * g.NY;
    const long long pin = Nglob / 2;   // int(N/2), NS:89
    c->pin_gx = (int)(pin / g.NY);
    c->pin_iy = (int)(pin % g.NY);
    const int n = g.P + 1;
    SEM_CUDA(cudaMalloc(&c->dD, sizeof(double) * n * n));
    SEM_CUDA(cudaMalloc(&c->dKs, sizeof(double) * n * n));
    SEM_CUDA(cudaMalloc(&c->dw, sizeof(double) * n));
    SEM_CUDA(cudaMemcpy(c->dD, d->D, sizeof(double) * n * n, cudaMemcpyHostToDevice));
    SEM_CUDA(cudaMemcpy(c->dKs, d->Ks, sizeof(double) * n * n, cudaMemcpyHostToDevice));
    SEM_CUDA(cudaMemcpy(c->dw, d->w, sizeof(double) * n, cudaMemcpyHostToDevice));
    if (g_orders[g.P - 1].upload(d->D, d->Ks, d->w)) { delete c; return -1; }
    if (g_orders[g.P - 1].upload3(d->D, d->Ks, d->w)) { delete c; return -1; }
    c->rs.max_blocks = 2 * c->sm_count;
    c->rs.max_k = SEM_MAX_RESTART + 8;
    SEM_CUDA(cudaMalloc(&c->rs.partials, sizeof(double) * (size_t)c->rs.max_blocks * c->rs.max_k));
    SEM_CUDA(cudaMalloc(&c->rs.counter, sizeof(unsigned) * (c->rs.max_k / 8 + 1)));
    SEM_CUDA(cudaMemset(c->rs.counter, 0, sizeof(unsigned) * (c->rs.max_k / 8 + 1)));
    c->rs.sm_count = c->sm_count;
    c->rs.axpy_blocks = 4096;                 // element blocks of 256: slices are used up to n = 1 M doubles
    c->rs.axpy_len = 4ll << 20;               // 4 M doubles of partial sums (32 MB)
    SEM_CUDA(cudaMalloc(&c->rs.axpy_partials, sizeof(double) * c->rs.axpy_len));
    SEM_CUDA(cudaMalloc(&c->rs.axpy_counter, sizeof(unsigned) * c->rs.axpy_blocks));
    SEM_CUDA(cudaMemset(c->rs.axpy_counter, 0, sizeof(unsigned) * c->rs.axpy_blocks));
    c->small_len = 2 * SEM_MAX_RESTART + 64;
    SEM_CUDA(cudaMalloc(&c->d_small, sizeof(double) * c->small_len));
    SEM_CUDA(cudaMallocHost(&c->h_small, sizeof(double) * c->small_len));
    c->ring_stride = (size_t)c->small_len;
    SEM_CUDA(cudaMallocHost(&c->h_ring, sizeof(double) * c->ring_stride * SEM_GMRES_LAG));
    for (int i = 0; i < SEM_GMRES_LAG; ++i) SEM_CUDA(cudaEventCreateWithFlags(&c->ev_ring[i], cudaEventDisableTiming));
    *out = c;
    return 0;
}

extern "C" void sem_ctx_destroy(sem_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaFree(c->dD); cudaFree(c->dKs); cudaFree(c->dw);
    if (c->dKdiag) cudaFree(c->dKdiag);
    if (c->dStage) cudaFree(c->dStage);
    for (auto& e : c->gcache)
        if (e.exec) cudaGraphExecDestroy(e.exec);
    if (c->has_comm) comm_destroy(c->comm);
    if (c->dStageOut) cudaFree(c->dStageOut);
    if (c->fQx) { cudaFree(c->fQx); cudaFree(c->fLx); cudaFree(c->fQy); cudaFree(c->fLy); cudaFree(c->fT1); cudaFree(c->fT2); }
    if (c->fB1) { cudaFree(c->fB1); cudaFree(c->fB2); }
    if (c->pbb_idx) { cudaFree(c->pbb_idx); cudaFree(c->pbb_inv); cudaFree(c->pbb_rhs); cudaFree(c->pbb_z); }
    if (c->blas) { cublasDestroy(c->blas); cudaFree(c->blas_ws); }
    if (c->streams_ready) {
        cudaStreamDestroy(c->s_side); cudaStreamDestroy(c->s_side2); cudaStreamDestroy(c->s_h2d); cudaStreamDestroy(c->s_d2h); cudaStreamDestroy(c->s_solve); cudaStreamDestroy(c->s_main);
        cudaEventDestroy(c->ev_g0); cudaEventDestroy(c->ev_g1);
        cudaEventDestroy(c->ev_in); cudaEventDestroy(c->ev_side); cudaEventDestroy(c->ev_edge); cudaEventDestroy(c->ev_start); cudaEventDestroy(c->ev_end);
        for (int i = 0; i < SEM_HOST_SEGMENTS; ++i) { cudaEventDestroy(c->ev_up[i]); cudaEventDestroy(c->ev_done[i]); }
    }
    cudaFree(c->rs.partials); cudaFree(c->rs.counter); cudaFree(c->rs.axpy_partials); cudaFree(c->rs.axpy_counter);
    cudaFree(c->d_small); cudaFreeHost(c->h_small); cudaFreeHost(c->h_ring);
    for (int i = 0; i < SEM_GMRES_LAG; ++i) cudaEventDestroy(c->ev_ring[i]);
    delete c;
}

extern "C" int sem_ctx_ld(const sem_ctx* c) { return c ? c->g.LD : -1; }
extern "C" int sem_ctx_nx(const sem_ctx* c) { return c ? c->g.NX : -1; }
extern "C" int sem_ctx_ny(const sem_ctx* c) { return c ? c->g.NY : -1; }
extern "C" long long sem_ctx_vec_len(const sem_ctx* c) { return c ? (long long)c->g.NX * c->g.LD : -1; }
extern "C" int sem_ctx_set_tiling(sem_ctx* c, int Ty, int Mx) {
    if (!c) return -2;
    c->Ty_req = Ty;
    c->Mx_req = Mx;
    return 0;
}

extern "C" int sem_nccl_unique_id(unsigned char* out128) { return comm_unique_id(out128); }

extern "C" int sem_ctx_attach_comm(sem_ctx* c, const unsigned char* id128, int rank, int world) {
    SEM_CHECK_CTX(c);
    if (c->has_comm) { set_error("sem_ctx_attach_comm: communicator already attached"); return -2; }
    if (comm_init(c->comm, id128, rank, world, c->g.NY)) return -1;
    c->has_comm = 1;
    return 0;
}

extern "C" int sem_ctx_comm_mode(const sem_ctx* c) {
    if (!c || !c->has_comm) return 0;
    return c->comm.p2p ? 2 : 1;
}

// interface exchange of freshly applied operator outputs (no-op on one GPU)
static int exchange(sem_ctx* c, std::initializer_list<double*> fields, cudaStream_t st) {
    if (!c->has_comm) return 0;
    double* f[8];
    int n = 0;
    for (double* p : fields)
        if (p) f[n++] = p;
    if (n == 0) return 0;
    return comm_exchange_add(c->comm, c->g, f, n, st);
}

// k dot products over the owned nodes of this rank, summed over all ranks (deterministic two-stage local sums)
static int ctx_multi_dot(sem_ctx* c, const double* V, long long n, int k, const double* w, double* h, int nf,
                         long long vlen, cudaStream_t st) {
    const long long skip = c->g.has_left ? c->g.LD : 0;   // the interface line is counted by the left rank
    if (aux_multi_dot(V, n, k, w, h, nf, vlen, skip, c->rs, st)) return -1;
    if (c->has_comm) return comm_allreduce_sum(c->comm, h, k, st);
    return 0;
}

static int ensure_stage(sem_ctx* c) {
    if (c->dStage) return 0;
    SEM_CUDA(cudaMalloc(&c->dStage, sizeof(double) * (size_t)c->g.NX * c->g.NY));
    return 0;
}

// One contiguous copy over PCIe (full rate from pinned memory) + a device repack kernel; a strided 2-D copy of
// 8193-double rows runs at a fraction of the link rate.
extern "C" int sem_h2d(sem_ctx* c, const double* host, double* vec, void* stream) {
    SEM_CHECK_CTX(c);
    const MeshDev& g = c->g;
    cudaStream_t st = (cudaStream_t)stream;
    if (ensure_stage(c)) return -1;
    SEM_CUDA(cudaMemcpyAsync(c->dStage, host, sizeof(double) * (size_t)g.NX * g.NY, cudaMemcpyHostToDevice, st));
    return aux_pad(g, c->dStage, vec, st);
}

extern "C" int sem_d2h(sem_ctx* c, const double* vec, double* host, void* stream) {
    SEM_CHECK_CTX(c);
    const MeshDev& g = c->g;
    cudaStream_t st = (cudaStream_t)stream;
    if (ensure_stage(c)) return -1;
    if (aux_unpad(g, vec, c->dStage, st)) return -1;
    SEM_CUDA(cudaMemcpyAsync(host, c->dStage, sizeof(double) * (size_t)g.NX * g.NY, cudaMemcpyDeviceToHost, st));
    SEM_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel generation: v3 (one warp per strip, TMA-staged, folded tables) for every order.  SEM_B200_MARCH=1 selects the
// round-1 v1 kernel (one column per thread, CTA barriers) for A/B comparisons.
static int march_generation() {
    static const int v = [] {
        const char* e = std::getenv("SEM_B200_MARCH");
        return (e && e[0] == '1') ? 1 : 3;
    }();
    return v;
}

// one fused operator launch over the element columns [m_lo, m_hi) of the slab (m_hi < 0: all)
static int march(sem_ctx* c, int mode, MarchArgs& A, cudaStream_t st, int m_lo = 0, int m_hi = -1) {
    A.zero = 0;
    if (m_hi < 0) m_hi = c->g.nex;
    if (m_lo >= m_hi) return 0;
    const auto& ord = g_orders[c->g.P - 1];
    const int gen = march_generation();
    if (gen == 3) {
        const MarchGeom q = march3_geometry(c->g, mode, c->Mx_req, c->sm_count, ord.smem3(mode), (size_t)c->smem_sm, m_lo, m_hi);
        return ord.launch3(mode, c->g, A, q, st);
    }
    MarchGeom q = march_geometry(c->g, c->Ty_req, c->Mx_req, c->sm_count, m_lo, m_hi);
    // shrink the strip until the tile fits the opt-in shared memory of the device
    while (g_orders[c->g.P - 1].smem(mode, q.pitch) > (size_t)c->smem_optin && q.Ty > 1) {
        q = march_geometry(c->g, q.Ty / 2 > 0 ? q.Ty / 2 : 1, c->Mx_req, c->sm_count, m_lo, m_hi);
    }
    return g_orders[c->g.P - 1].launch(mode, c->g, A, q, st);
}

static int ensure_streams(sem_ctx* c) {
    if (c->streams_ready) return 0;
    int lo = 0, hi = 0;
    SEM_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // lo = least priority
    SEM_CUDA(cudaStreamCreateWithPriority(&c->s_side, cudaStreamNonBlocking, lo));
    SEM_CUDA(cudaStreamCreateWithPriority(&c->s_side2, cudaStreamNonBlocking, hi));
    SEM_CUDA(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
    SEM_CUDA(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
    SEM_CUDA(cudaStreamCreateWithFlags(&c->s_solve, cudaStreamNonBlocking));
    SEM_CUDA(cudaStreamCreateWithFlags(&c->s_main, cudaStreamNonBlocking));
    SEM_CUDA(cudaEventCreateWithFlags(&c->ev_g0, cudaEventDisableTiming));
    SEM_CUDA(cudaEventCreateWithFlags(&c->ev_g1, cudaEventDisableTiming));
    SEM_CUDA(cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming));
    SEM_CUDA(cudaEventCreateWithFlags(&c->ev_side, cudaEventDisableTiming));
    SEM_CUDA(cudaEventCreateWithFlags(&c->ev_edge, cudaEventDisableTiming));
    SEM_CUDA(cudaEventCreateWithFlags(&c->ev_start, cudaEventDisableTiming));
    SEM_CUDA(cudaEventCreateWithFlags(&c->ev_end, cudaEventDisableTiming));
    for (int i = 0; i < SEM_HOST_SEGMENTS; ++i) {
        SEM_CUDA(cudaEventCreateWithFlags(&c->ev_up[i], cudaEventDisableTiming));
        SEM_CUDA(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
    }
    c->streams_ready = 1;
    return 0;
}

// A fused operator apply followed by the interface exchange of its outputs (`fields`; null entries are skipped).
// `post` (optional) runs after the operator and before the exchange (the O(sqrt N) pressure-Neumann rows of NS).
// One GPU: operator, post.  Partitioned: the element columns next to the interfaces are applied first, their interface
// lines travel (NCCL send/recv on the caller's stream) while the interior columns are applied on a low-priority side
// stream, then the received partial sums are added.  `post` is idempotent and cheap, so it simply runs before the
// transfer (interface lines final) and again after the interior (interior lines final).
// a launch this narrow is pure latency (one warp per strip, 4 marching steps); SEM_B200_EDGE_COLUMNS=0 disables the split
// (whole slab in one launch, then the exchange) for A/B runs
static const int SEM_EDGE_COLUMNS = [] {
    const char* e = std::getenv("SEM_B200_EDGE_COLUMNS");
    return e ? std::max(0, std::atoi(e)) : 4;
}();

typedef std::function<int(cudaStream_t)> postop;

// the partitioned sequence itself, everything ordered on `st` (and the side stream forked from / joined to it)
static int partitioned_apply(sem_ctx* c, int mode, MarchArgs& A, double* const* f, int n, int el, int er, cudaStream_t st,
                             const postop& post) {
    const int nex = c->g.nex;
    SEM_CUDA(cudaEventRecord(c->ev_in, st));                 // inputs are ready here
    SEM_CUDA(cudaStreamWaitEvent(c->s_side, c->ev_in, 0));
    // A large interior (several resident rounds of one-warp CTAs) goes first: it starts at once, and the edge launches on
    // the higher-priority stream slip in as soon as its first CTAs retire -- long before it ends.  A small interior
    // (8 GPUs on config 5: about one round) would delay the edges past its own end, so there the edges go first.
    const long long interior_ctas = (long long)(c->g.ney / 8 + 1) * ((nex - el - er + 15) / 16);
    const bool interior_first = interior_ctas >= 3ll * 6 * c->sm_count;
    if (interior_first) {
        if (march(c, mode, A, c->s_side, el, nex - er)) return -1;
        SEM_CUDA(cudaEventRecord(c->ev_side, c->s_side));
        if (march(c, mode, A, st, 0, el)) return -1;
        if (march(c, mode, A, st, nex - er, nex)) return -1;
    } else {
        // three concurrent launches, the edges submitted first so that their CTAs are placed first
        SEM_CUDA(cudaStreamWaitEvent(c->s_side2, c->ev_in, 0));
        if (march(c, mode, A, st, 0, el)) return -1;                   // left edge
        if (march(c, mode, A, c->s_side2, nex - er, nex)) return -1;   // right edge
        SEM_CUDA(cudaEventRecord(c->ev_edge, c->s_side2));
        if (march(c, mode, A, c->s_side, el, nex - er)) return -1;     // interior (low priority)
        SEM_CUDA(cudaEventRecord(c->ev_side, c->s_side));
        SEM_CUDA(cudaStreamWaitEvent(st, c->ev_edge, 0));
    }
    if (!post) {
        // The interface lines are written by the edge launches only: the whole exchange (push, wait for the neighbour, add)
        // is one kernel that runs while the interior columns are still being applied on the side stream.
        if (comm_exchange_fused(c->comm, c->g, f, n, st)) return -1;
        SEM_CUDA(cudaStreamWaitEvent(st, c->ev_side, 0));
        return 0;
    }
    // with a post-operator (NS: boundary rows recomputed after the interior) the add has to come last
    if (post(st)) return -1;
    if (comm_exchange_transfer(c->comm, c->g, f, n, st)) return -1;
    SEM_CUDA(cudaStreamWaitEvent(st, c->ev_side, 0));
    if (post(st)) return -1;
    return comm_exchange_finish(c->comm, c->g, f, n, st);
}

static int apply_and_exchange(sem_ctx* c, int mode, MarchArgs& A, std::initializer_list<double*> fields, cudaStream_t st,
                              const postop& post = nullptr) {
    if (!c->has_comm) {
        if (march(c, mode, A, st)) return -1;
        return post ? post(st) : 0;
    }
    double* f[8];
    int n = 0;
    for (double* p : fields)
        if (p) f[n++] = p;
    const int nex = c->g.nex;
    const int el = c->g.has_left ? std::min(SEM_EDGE_COLUMNS, nex) : 0;
    const int er = c->g.has_right ? std::min(SEM_EDGE_COLUMNS, nex - el) : 0;
    if (el + er >= nex || el + er == 0 || n == 0) {   // nothing (left) to overlap with
        if (march(c, mode, A, st)) return -1;
        if (post && post(st)) return -1;
        return n ? comm_exchange_add(c->comm, c->g, f, n, st) : 0;
    }
    if (ensure_streams(c)) return -1;
    static const bool no_graph = std::getenv("SEM_B200_NO_GRAPH") != nullptr;
    if (no_graph) return partitioned_apply(c, mode, A, f, n, el, er, st, post);

    // The sequence is ~12 host calls (3 launches, 6 event operations, a grouped NCCL send/recv, the add): at 8 GPUs the
    // host's enqueue time exceeds the device time of the apply.  The second time the same apply (mode, arguments, fields)
    // comes along it is captured -- NCCL operations included, every rank captures in lock step -- and from then on it is
    // replayed with one cudaGraphLaunch.  Capture is not allowed on the legacy default stream a caller may hand in, so the
    // partitioned applies run on a context-owned stream ordered against the caller's by two events.
    A.zero = 0;
    unsigned char key[sizeof(c->gcache[0].key)];
    std::memset(key, 0, sizeof(key));
    std::memcpy(key, &A, sizeof(A));
    std::memcpy(key + sizeof(A), f, sizeof(double*) * n);
    const int tail[4] = {mode, n, c->Mx_req, c->Ty_req};
    std::memcpy(key + sizeof(A) + 8 * sizeof(double*), tail, sizeof(tail));
    sem_ctx::GraphEntry* e = nullptr;
    for (auto& g : c->gcache)
        if (g.uses > 0 && std::memcmp(g.key, key, sizeof(key)) == 0) e = &g;
    cudaStream_t sm = c->s_main;
    SEM_CUDA(cudaEventRecord(c->ev_g0, st));
    SEM_CUDA(cudaStreamWaitEvent(sm, c->ev_g0, 0));
    int rc = 0;
    if (e && e->exec) {
        SEM_CUDA(cudaGraphLaunch(e->exec, sm));
    } else if (!e) {                      // first sighting: run eagerly (also establishes the NCCL connections)
        e = &c->gcache[c->gcache_next];
        c->gcache_next = (c->gcache_next + 1) % 8;
        if (e->exec) cudaGraphExecDestroy(e->exec);
        e->exec = nullptr;
        std::memcpy(e->key, key, sizeof(key));
        e->uses = 1;
        rc = partitioned_apply(c, mode, A, f, n, el, er, sm, post);
    } else {                              // second sighting: capture, instantiate, launch
        e->uses++;
        cudaGraph_t graph = nullptr;
        SEM_CUDA(cudaStreamBeginCapture(sm, cudaStreamCaptureModeThreadLocal));
        rc = partitioned_apply(c, mode, A, f, n, el, er, sm, post);
        const cudaError_t ce = cudaStreamEndCapture(sm, &graph);
        if (rc || ce != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            set_error(std::string("apply_and_exchange: graph capture failed: ") + cudaGetErrorString(ce));
            return -1;
        }
        SEM_CUDA(cudaGraphInstantiate(&e->exec, graph, 0));
        cudaGraphDestroy(graph);
        SEM_CUDA(cudaGraphLaunch(e->exec, sm));
    }
    SEM_CUDA(cudaEventRecord(c->ev_g1, sm));
    SEM_CUDA(cudaStreamWaitEvent(st, c->ev_g1, 0));
    return rc;
}

static MarchArgs zero_args() {
    MarchArgs A;
    std::memset(&A, 0, sizeof(A));
    A.bc.pin_gx = -1;
    A.bc.pin_iy = -1;
    return A;
}

static void fill_cd_bc(BCSpec& bc, const sem_cd_bc& in, int residual) {
    for (int s = 0; s < 4; ++s) {
        bc.active[s] = in.active[s];
        bc.val0[s] = in.value[s];
        bc.val1[s] = 0.0;
    }
    bc.residual = residual;
    bc.pin_gx = bc.pin_iy = -1;
}

static void fill_ns_bc(const sem_ctx* c, BCSpec& bc, const sem_ns_bc& in, int residual) {
    // NS:78-88: W, E, S, N in that order, later wins: (u,v) = (0,v_W), (0,v_E), (u_S,0), (u_N,0)
    const double u[4] = {0.0, 0.0, in.u_S, in.u_N};
    const double v[4] = {in.v_W, in.v_E, 0.0, 0.0};
    for (int s = 0; s < 4; ++s) {
        bc.active[s] = 1;
        bc.val0[s] = u[s];
        bc.val1[s] = v[s];
    }
    bc.residual = residual;
    bc.pin_gx = c->pin_gx;
    bc.pin_iy = c->pin_iy;
}

extern "C" int sem_apply_stiffness(sem_ctx* c, const double* x, double* y, void* stream) {
    SEM_CHECK_CTX(c);
    MarchArgs A = zero_args();
    A.a = x;
    A.y0 = y;
    return apply_and_exchange(c, MODE_K, A, {y}, (cudaStream_t)stream);
}

extern "C" int sem_apply_gradient(sem_ctx* c, const double* x, double scale, double* gx, double* gy, void* stream) {
    SEM_CHECK_CTX(c);
    MarchArgs A = zero_args();
    A.a = x;
    A.y0 = gx;
    A.y1 = gy;
    A.cconv = scale;
    return apply_and_exchange(c, MODE_G, A, {gx, gy}, (cudaStream_t)stream);
}

extern "C" int sem_apply_mass(sem_ctx* c, const double* x, double* y, void* stream) {
    SEM_CHECK_CTX(c);
    if (aux_mass_apply(c->g, c->tab(), x, y, (cudaStream_t)stream)) return -1;
    return exchange(c, {y}, (cudaStream_t)stream);
}

extern "C" int sem_mass_diag(sem_ctx* c, double* m, void* stream) {
    SEM_CHECK_CTX(c);
    if (aux_mass_apply(c->g, c->tab(), nullptr, m, (cudaStream_t)stream)) return -1;
    return exchange(c, {m}, (cudaStream_t)stream);
}

extern "C" int sem_gather_scatter(sem_ctx* c, const double* elem, double* y, void* stream) {
    SEM_CHECK_CTX(c);
    if (aux_gather_scatter(c->g, elem, y, (cudaStream_t)stream)) return -1;
    return exchange(c, {y}, (cudaStream_t)stream);
}

extern "C" int sem_scatter(sem_ctx* c, const double* x, double* elem, void* stream) {
    SEM_CHECK_CTX(c);
    return aux_scatter(c->g, x, elem, (cudaStream_t)stream);
}

// ---- convection-diffusion ---------------------------------------------------------------------------------------
extern "C" int sem_cd_residual(sem_ctx* c, const sem_cd_state* s, const double* T, double* res, void* stream) {
    SEM_CHECK_CTX(c);
    MarchArgs A = zero_args();
    A.a = T; A.U = s->u; A.V = s->v; A.cconv = s->Pe; A.y0 = res;
    fill_cd_bc(A.bc, s->bc, 1);
    return apply_and_exchange(c, MODE_CD, A, {res}, (cudaStream_t)stream);
}

extern "C" int sem_cd_jacobians(sem_ctx* c, double Pe, const double* T, double* gxT, double* gyT, void* stream) {
    return sem_apply_gradient(c, T, Pe, gxT, gyT, stream);
}

extern "C" int sem_cd_jvp(sem_ctx* c, const sem_cd_state* s, const double* dT, const double* du, const double* dv,
                          double* dres, void* stream) {
    SEM_CHECK_CTX(c);
    if ((du || dv) && (!s->gxT || !s->gyT)) { set_error("sem_cd_jvp: du/dv given but no Jacobians in the state"); return -2; }
    MarchArgs A = zero_args();
    A.a = dT; A.U = s->u; A.V = s->v; A.cconv = s->Pe; A.y0 = dres;
    A.d0 = s->gxT; A.e0 = du; A.d1 = s->gyT; A.e1 = dv;
    fill_cd_bc(A.bc, s->bc, 0);
    return apply_and_exchange(c, MODE_CD, A, {dres}, (cudaStream_t)stream);
}

// Host-buffer variant of sem_cd_jvp: three streams, SEM_HOST_SEGMENTS segments of element columns.
//   s_h2d : dense upload of the node lines of segment s            -> ev_up[s]
//   stream: wait ev_up[s]; pad lines; fused apply of the segment's element columns; unpad the finished lines -> ev_done[s]
//   s_d2h : wait ev_done[s]; dense download of the finished lines
// A segment [ma, mb) reads the lines (ma-1)*P .. mb*P (the first P+1 of them were uploaded with the previous segment)
// and finishes the lines ma*P .. mb*P-1, plus the last line of the slab when mb is the last column.
extern "C" int sem_cd_jvp_host(sem_ctx* c, const sem_cd_state* s, const double* host_dT, double* host_dres,
                               double* dT_vec, double* dres_vec, void* stream) {
    SEM_CHECK_CTX(c);
    const MeshDev& g = c->g;
    cudaStream_t st = (cudaStream_t)stream;
    if (ensure_stage(c) || ensure_streams(c)) return -1;
    if (!c->dStageOut) SEM_CUDA(cudaMalloc(&c->dStageOut, sizeof(double) * (size_t)g.NX * g.NY));
    MarchArgs A = zero_args();
    A.a = dT_vec; A.U = s->u; A.V = s->v; A.cconv = s->Pe; A.y0 = dres_vec;
    fill_cd_bc(A.bc, s->bc, 0);
    // segments of whole 16-column chunks (the chunk size of the kernel), at most SEM_HOST_SEGMENTS
    const int chunk = 16;
    const int nchunks = (g.nex + chunk - 1) / chunk;
    static const int want = [] {
        const char* e = std::getenv("SEM_B200_HOST_SEGMENTS");
        const int v = e ? std::atoi(e) : 16;
        return std::max(1, std::min(SEM_HOST_SEGMENTS, v));
    }();
    const int nseg = std::max(1, std::min(want, nchunks));
    SEM_CUDA(cudaEventRecord(c->ev_start, st));              // earlier work on the caller's stream (state vectors ...)
    SEM_CUDA(cudaStreamWaitEvent(c->s_h2d, c->ev_start, 0));
    SEM_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_start, 0));
    const size_t NY = (size_t)g.NY;
    for (int k = 0; k < nseg; ++k) {
        const int ma = (int)((long long)nchunks * k / nseg) * chunk;
        const int mb = std::min(g.nex, (int)((long long)nchunks * (k + 1) / nseg) * chunk);
        const int up0 = (k == 0) ? 0 : ma * g.P + 1, up1 = mb * g.P;                  // lines uploaded for this segment
        const int dn0 = ma * g.P, dn1 = (mb == g.nex) ? g.nex * g.P : mb * g.P - 1;   // lines this segment finishes
        SEM_CUDA(cudaMemcpyAsync(c->dStage + (size_t)up0 * NY, host_dT + (size_t)up0 * NY, sizeof(double) * (up1 - up0 + 1) * NY,
                                 cudaMemcpyHostToDevice, c->s_h2d));
        SEM_CUDA(cudaEventRecord(c->ev_up[k], c->s_h2d));
        SEM_CUDA(cudaStreamWaitEvent(st, c->ev_up[k], 0));
        if (aux_pad_lines(g, c->dStage, dT_vec, up0, up1 - up0 + 1, st)) return -1;
        if (march(c, MODE_CD, A, st, ma, mb)) return -1;
        if (aux_unpad_lines(g, dres_vec, c->dStageOut, dn0, dn1 - dn0 + 1, st)) return -1;
        SEM_CUDA(cudaEventRecord(c->ev_done[k], st));
        SEM_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_done[k], 0));
        SEM_CUDA(cudaMemcpyAsync(host_dres + (size_t)dn0 * NY, c->dStageOut + (size_t)dn0 * NY,
                                 sizeof(double) * (dn1 - dn0 + 1) * NY, cudaMemcpyDeviceToHost, c->s_d2h));
    }
    if (c->has_comm && (g.has_left || g.has_right)) {
        // Partitioned mesh: the interface line(s) downloaded above hold this rank's element sums only.  Exchange them with the
        // neighbour(s) (peer-memory mailboxes), then download the completed line(s) again -- the download stream is in order,
        // so the second copy of a line lands after the first.
        double* f[1] = {dres_vec};
        if (comm_exchange_add(c->comm, g, f, 1, st)) return -1;
        const int lines[2] = {g.has_left ? 0 : -1, g.has_right ? g.NX - 1 : -1};
        for (int q = 0; q < 2; ++q)
            if (lines[q] >= 0 && aux_unpad_lines(g, dres_vec, c->dStageOut, lines[q], 1, st)) return -1;
        SEM_CUDA(cudaEventRecord(c->ev_edge, st));
        SEM_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_edge, 0));
        for (int q = 0; q < 2; ++q)
            if (lines[q] >= 0)
                SEM_CUDA(cudaMemcpyAsync(host_dres + (size_t)lines[q] * NY, c->dStageOut + (size_t)lines[q] * NY,
                                         sizeof(double) * NY, cudaMemcpyDeviceToHost, c->s_d2h));
    }
    SEM_CUDA(cudaEventRecord(c->ev_end, c->s_d2h));
    SEM_CUDA(cudaStreamWaitEvent(st, c->ev_end, 0));         // the caller's stream stays the single point of ordering
    SEM_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// ---- Navier-Stokes ------------------------------------------------------------------------------------------------
extern "C" int sem_ns_residual(sem_ctx* c, const sem_ns_state* s, const double* u, const double* v, const double* p,
                               const double* T, double* ru, double* rv, double* rc, void* stream) {
    SEM_CHECK_CTX(c);
    MarchArgs A = zero_args();
    A.a = u; A.b = v; A.c = p; A.U = u; A.V = v; A.cconv = s->Re;
    A.e0 = T; A.cbuoy = -s->Gr_over_Re;
    A.y0 = ru; A.y1 = rv; A.y2 = rc;
    fill_ns_bc(c, A.bc, s->bc, 1);
    // NS:116-119: pin first, then the Neumann rows (they win if the pin sits on the boundary)
    return apply_and_exchange(c, MODE_NS, A, {ru, rv, rc}, (cudaStream_t)stream, [&](cudaStream_t s) {
        return aux_neumann_rows(c->g, c->tab(), p, rc, c->pin_gx, c->pin_iy, 0, s);
    });
}

extern "C" int sem_ns_jacobians(sem_ctx* c, double Re, const double* u, const double* v, double* gxu, double* gyu,
                                double* gxv, double* gyv, void* stream) {
    int r = sem_apply_gradient(c, u, Re, gxu, gyu, stream);
    if (r) return r;
    return sem_apply_gradient(c, v, Re, gxv, gyv, stream);
}

extern "C" int sem_ns_jvp(sem_ctx* c, const sem_ns_state* s, const double* du, const double* dv, const double* dp,
                          const double* dT, double* ou, double* ov, double* oc, void* stream) {
    SEM_CHECK_CTX(c);
    if (!s->gxu || !s->gyu || !s->gxv || !s->gyv) { set_error("sem_ns_jvp: Jacobians missing in the state"); return -2; }
    MarchArgs A = zero_args();
    A.a = du; A.b = dv; A.c = dp; A.U = s->u; A.V = s->v; A.cconv = s->Re;
    A.d0 = s->gxu; A.d1 = s->gyu; A.d2 = s->gxv; A.d3 = s->gyv;
    A.e0 = dT; A.cbuoy = -s->Gr_over_Re;
    A.y0 = ou; A.y1 = ov; A.y2 = oc;
    fill_ns_bc(c, A.bc, s->bc, 0);
    // NS:157-158: Neumann rows first, then the pin (the pin wins)
    return apply_and_exchange(c, MODE_NS, A, {ou, ov, oc}, (cudaStream_t)stream, [&](cudaStream_t s) {
        return aux_neumann_rows(c->g, c->tab(), dp, oc, c->pin_gx, c->pin_iy, 1, s);
    });
}

// ---- reductions ------------------------------------------------------------------------------------------------------
extern "C" int sem_dot(sem_ctx* c, const double* x, const double* y, long long n, double* host_out, void* stream) {
    SEM_CHECK_CTX(c);
    cudaStream_t st = (cudaStream_t)stream;
    const long long vlen = (long long)c->g.NX * c->g.LD;
    if (n % vlen != 0) { set_error("sem_dot: n must be a multiple of the field length"); return -2; }
    if (ctx_multi_dot(c, x, n, 1, y, c->d_small, (int)(n / vlen), vlen, st)) return -1;
    SEM_CUDA(cudaMemcpyAsync(c->h_small, c->d_small, sizeof(double), cudaMemcpyDeviceToHost, st));
    SEM_CUDA(cudaStreamSynchronize(st));
    *host_out = c->h_small[0];
    return 0;
}

extern "C" int sem_axpby(sem_ctx* c, double a, const double* x, double b, double* y, long long n, void* stream) {
    SEM_CHECK_CTX(c);
    return aux_axpby(a, x, b, y, n, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------------------
// Fast-diagonalisation (FDM) preconditioner.  On the reference's meshes (uniform rectangular element grid, SEM.py:170-203)
// the assembled stiffness matrix is a Kronecker sum of 1-D operators, K = K1x (x) M1y + M1x (x) K1y, and every side is
// either all Dirichlet or all Neumann, so the Laplacian with the Dirichlet rows eliminated is inverted exactly by the
// generalised eigenpairs of the two 1-D pencils:   K^-1 = (Qx (x) Qy) diag(1/(lx_i + ly_j)) (Qx (x) Qy)^T,
// Q^T M1 Q = I, Q^T K1 Q = diag(l).  Four dense fp64 GEMMs (cuBLAS: plain library GEMMs on the tensor cores) + one
// scaling pass per application; the iteration count of the Krylov solver no longer grows with the mesh.
// Q is stored row-major [n][n] with zero rows at Dirichlet end nodes, so the result is zero there and the identity rows
// of the operator are restored by a boundary pass (z = r).
// ---------------------------------------------------------------------------------------------------------------
#define SEM_CUBLAS(call)                                                                                   \
    do {                                                                                                   \
        cublasStatus_t _s = (call);                                                                        \
        if (_s != CUBLAS_STATUS_SUCCESS) {                                                                 \
            set_error(std::string(#call) + " failed with cuBLAS status " + std::to_string((int)_s));      \
            return -1;                                                                                     \
        }                                                                                                  \
    } while (0)

extern "C" int sem_ctx_set_fdm(sem_ctx* c, const double* Qx, const double* lamx, const double* Qy, const double* lamy,
                               const int* dirichlet_wesn) {
    SEM_CHECK_CTX(c);
    const size_t nx = (size_t)c->g.NX, ny = (size_t)c->g.NY, ld = (size_t)c->g.LD;
    const size_t nxg = (size_t)c->g.NXg;                       // == nx on one GPU
    const int world = c->has_comm ? c->comm.world : 1;
    if (!c->has_comm && (c->g.has_left || c->g.has_right)) {
        set_error("sem_ctx_set_fdm: a partitioned context needs its communicator first (sem_ctx_attach_comm)");
        return -2;
    }
    const size_t block = (nxg + world - 1) / world;            // spectral modes per rank
    const size_t t_rows = (world > 1) ? block * world : nx;    // rows of the transform buffers
    if (c->fQx) {
        cudaFree(c->fQx); cudaFree(c->fLx); cudaFree(c->fQy); cudaFree(c->fLy); cudaFree(c->fT1); cudaFree(c->fT2);
        c->fQx = nullptr;
    }
    SEM_CUDA(cudaMalloc(&c->fQx, sizeof(double) * nx * nxg));
    SEM_CUDA(cudaMalloc(&c->fLx, sizeof(double) * t_rows));
    SEM_CUDA(cudaMalloc(&c->fQy, sizeof(double) * ny * ny));
    SEM_CUDA(cudaMalloc(&c->fLy, sizeof(double) * ny));
    SEM_CUDA(cudaMalloc(&c->fT1, sizeof(double) * t_rows * ld));
    SEM_CUDA(cudaMalloc(&c->fT2, sizeof(double) * t_rows * ld));
    SEM_CUDA(cudaMemset(c->fT1, 0, sizeof(double) * t_rows * ld));   // the GEMMs never touch the pad columns
    SEM_CUDA(cudaMemset(c->fT2, 0, sizeof(double) * t_rows * ld));
    if (!c->blas) {
        SEM_CUBLAS(cublasCreate(&c->blas));
        SEM_CUDA(cudaMalloc(&c->blas_ws, (size_t)32 << 20));   // fixed workspace: the GEMMs are captured into CUDA graphs
        SEM_CUBLAS(cublasSetWorkspace(c->blas, c->blas_ws, (size_t)32 << 20));
    }
    SEM_CUDA(cudaMemcpy(c->fQx, Qx, sizeof(double) * nx * nxg, cudaMemcpyDeviceToDevice));
    SEM_CUDA(cudaMemset(c->fLx, 0, sizeof(double) * t_rows));
    SEM_CUDA(cudaMemcpy(c->fLx, lamx, sizeof(double) * nxg, cudaMemcpyDeviceToDevice));
    SEM_CUDA(cudaMemcpy(c->fQy, Qy, sizeof(double) * ny * ny, cudaMemcpyDeviceToDevice));
    SEM_CUDA(cudaMemcpy(c->fLy, lamy, sizeof(double) * ny, cudaMemcpyDeviceToDevice));
    for (int k = 0; k < 4; ++k) c->fdm_dir[k] = dirichlet_wesn[k];
    c->fdm_nxg = (int)nxg;
    c->fdm_block = (int)block;
    c->fdm_ready = 1;
    return 0;
}

// z = K_II^-1 r on the nodes that carry no Dirichlet row, z = r on the Dirichlet nodes.  r and z may not alias.
// Partitioned mesh: the x transform couples all slabs, so it is distributed -- every rank forms its contribution
// Qx_r^T R_r to all modes (one GEMM over its own lines; the duplicated interface line is taken from the left rank only), a
// reduce-scatter over NVLink leaves each rank with a block of modes, the y transforms and the spectral scaling are local to
// that block, an all-gather returns all modes and one GEMM with the slab's rows of Qx gives the slab of the result.  The
// preconditioner is the exact inverse on any number of GPUs: the Krylov iteration count does not depend on the partition.
static int fdm_apply(sem_ctx* c, const double* r, double* z, cudaStream_t st) {
    if (!c->fdm_ready) { set_error("fdm_apply: sem_ctx_set_fdm has not been called"); return -2; }
    const int nx = c->g.NX, ny = c->g.NY, ld = c->g.LD, nxg = c->fdm_nxg;
    const double one = 1.0, zero = 0.0;
    SEM_CUBLAS(cublasSetStream(c->blas, st));
    // a row-major [NX][LD] vec is the column-major matrix R^T (ny x nx, leading dimension LD); a row-major Q is the
    // column-major Q^T.  Steps: T1 = Qx^T R, Z = T1 Qy, Z /= (lx + ly), T2 = Z Qy^T, X = Qx T2 -- written for the transposes.
    if (!c->has_comm) {
        SEM_CUBLAS(cublasDgemm(c->blas, CUBLAS_OP_N, CUBLAS_OP_T, ny, nx, nx, &one, r, ld, c->fQx, nx, &zero, c->fT1, ld));
        SEM_CUBLAS(cublasDgemm(c->blas, CUBLAS_OP_N, CUBLAS_OP_N, ny, nx, ny, &one, c->fQy, ny, c->fT1, ld, &zero, c->fT2, ld));
        if (aux_fdm_scale(c->g, c->fLx, c->fLy, c->fT2, nx, st)) return -1;
        SEM_CUBLAS(cublasDgemm(c->blas, CUBLAS_OP_T, CUBLAS_OP_N, ny, nx, ny, &one, c->fQy, ny, c->fT2, ld, &zero, c->fT1, ld));
        SEM_CUBLAS(cublasDgemm(c->blas, CUBLAS_OP_N, CUBLAS_OP_N, ny, nx, nx, &one, c->fT1, ld, c->fQx, nx, &zero, z, ld));
        return aux_fdm_boundary(c->g, c->fdm_dir, r, z, st);
    }
    const int B = c->fdm_block, me = c->comm.rank;
    const int skip = c->g.has_left ? 1 : 0;                    // the interface line is summed from the left rank's copy
    double* mine1 = c->fT2;                                    // this rank's block of modes (B x LD), then Z
    double* mine2 = c->fT2 + (size_t)B * ld;                   // T2 block
    // contribution of this slab to all modes: T1p (nxg x ny) = Qx_r[skip:, :]^T R_r[skip:, :]
    SEM_CUBLAS(cublasDgemm(c->blas, CUBLAS_OP_N, CUBLAS_OP_T, ny, nxg, nx - skip, &one, r + (size_t)skip * ld, ld,
                           c->fQx + (size_t)skip * nxg, nxg, &zero, c->fT1, ld));
    if (comm_reduce_scatter_sum(c->comm, c->fT1, mine1, (size_t)B * ld, st)) return -1;
    SEM_CUBLAS(cublasDgemm(c->blas, CUBLAS_OP_N, CUBLAS_OP_N, ny, B, ny, &one, c->fQy, ny, mine1, ld, &zero, mine2, ld));
    if (aux_fdm_scale(c->g, c->fLx + (size_t)me * B, c->fLy, mine2, B, st)) return -1;
    SEM_CUBLAS(cublasDgemm(c->blas, CUBLAS_OP_T, CUBLAS_OP_N, ny, B, ny, &one, c->fQy, ny, mine2, ld, &zero, mine1, ld));
    if (comm_allgather(c->comm, mine1, c->fT1, (size_t)B * ld, st)) return -1;
    SEM_CUBLAS(cublasDgemm(c->blas, CUBLAS_OP_N, CUBLAS_OP_N, ny, nx, nxg, &one, c->fT1, ld, c->fQx, nxg, &zero, z, ld));
    return aux_fdm_boundary(c->g, c->fdm_dir, r, z, st);
}

// ---------------------------------------------------------------------------------------------------------------
// EXPERIMENTAL (not validated on a GPU in round 1, off unless sem_krylov.precond == 3): boundary block of the NS pressure
// rows.  The pressure-Neumann rows K[mask,:] p (NS:119,157) are scaled by 1/M in the reference's Schur preconditioner, which
// leaves eigenvalues ~ 2e4 and costs two thirds of the Krylov iterations (scratch/schur/p12.py).  Block
// elimination instead:  z_I = mass sweep as before,  z_B = K_BB^-1 (r_B - K_BI z_I)  with the dense inverse of K restricted
// to the boundary pressure nodes (a ring of 4(n-1) nodes; the pin node is excluded and keeps its identity row).  The member
// of the singular system's solution set is unchanged: l_c vanishes on the boundary, so S~^T l_c = M_p l_c still holds.
// idx_host[nb]: offsets ix*LD + iy of the boundary nodes;  inv_dev: DEVICE, row-major nb x nb.  One GPU only.
// ---------------------------------------------------------------------------------------------------------------
extern "C" int sem_ctx_set_pbb(sem_ctx* c, const long long* idx_host, int nb, const double* inv_dev) {
    SEM_CHECK_CTX(c);
    if (c->has_comm) { set_error("sem_ctx_set_pbb: not available on a partitioned context"); return -2; }
    if (nb <= 0 || !idx_host || !inv_dev) { set_error("sem_ctx_set_pbb: bad arguments"); return -2; }
    const long long vlen = (long long)c->g.NX * c->g.LD;
    for (int a = 0; a < nb; ++a)
        if (idx_host[a] < 0 || idx_host[a] >= vlen) { set_error("sem_ctx_set_pbb: node offset out of range"); return -2; }
    if (c->pbb_idx) { cudaFree(c->pbb_idx); cudaFree(c->pbb_inv); cudaFree(c->pbb_rhs); cudaFree(c->pbb_z); c->pbb_idx = nullptr; }
    c->pbb_n = 0;
    SEM_CUDA(cudaMalloc(&c->pbb_idx, sizeof(long long) * nb));
    SEM_CUDA(cudaMalloc(&c->pbb_inv, sizeof(double) * (size_t)nb * nb));
    SEM_CUDA(cudaMalloc(&c->pbb_rhs, sizeof(double) * nb));
    SEM_CUDA(cudaMalloc(&c->pbb_z, sizeof(double) * nb));
    SEM_CUDA(cudaMemcpy(c->pbb_idx, idx_host, sizeof(long long) * nb, cudaMemcpyHostToDevice));
    SEM_CUDA(cudaMemcpy(c->pbb_inv, inv_dev, sizeof(double) * (size_t)nb * nb, cudaMemcpyDeviceToDevice));
    c->pbb_n = nb;
    return 0;
}

// z_p (already holding the mass sweep) -> boundary entries replaced by K_BB^-1 (r_B - K_BI z_I); q: scratch vec
static int pbb_apply(sem_ctx* c, const double* rc, double* zp, double* q, cudaStream_t st) {
    const int nb = c->pbb_n;
    if (nb <= 0) { set_error("pbb_apply: sem_ctx_set_pbb has not been called"); return -2; }
    if (aux_pbb_zero(zp, c->pbb_idx, nb, st)) return -1;
    // (K z)[boundary nodes] with z_B = 0 is K_BI z_I (the pin node, if it lies on the boundary, is skipped and keeps r_pin)
    if (aux_neumann_rows(c->g, c->tab(), zp, q, c->pin_gx, c->pin_iy, 1, st)) return -1;
    if (aux_pbb_gather(rc, q, c->pbb_idx, nb, c->pbb_rhs, st)) return -1;
    const double one = 1.0, zero = 0.0;
    SEM_CUBLAS(cublasSetStream(c->blas, st));
    // row-major INV is the column-major INV^T: y = (INV^T)^T x
    SEM_CUBLAS(cublasDgemv(c->blas, CUBLAS_OP_T, nb, nb, &one, c->pbb_inv, nb, c->pbb_rhs, 1, &zero, c->pbb_z, 1));
    return aux_pbb_scatter(c->pbb_z, c->pbb_idx, nb, zp, st);
}

// Two fields at once (the velocity components of the NS preconditioner, `stride` doubles apart): the same four GEMMs as
// strided-batched calls, one scaling and one boundary launch -- 6 launches instead of 12.  On the reference's meshes
// (65 x 65 nodes) every one of them is pure launch latency (4.3 us per GEMM, profiles/README.md).
static int fdm_apply2(sem_ctx* c, const double* r, double* z, long long stride, cudaStream_t st) {
    if (c->has_comm) return (fdm_apply(c, r, z, st) || fdm_apply(c, r + stride, z + stride, st)) ? -1 : 0;
    if (!c->fdm_ready) { set_error("fdm_apply: sem_ctx_set_fdm has not been called"); return -2; }
    const int nx = c->g.NX, ny = c->g.NY, ld = c->g.LD;
    const long long fs = (long long)nx * ld;
    if (!c->fB1) {
        SEM_CUDA(cudaMalloc(&c->fB1, sizeof(double) * 2 * fs));
        SEM_CUDA(cudaMalloc(&c->fB2, sizeof(double) * 2 * fs));
        SEM_CUDA(cudaMemset(c->fB1, 0, sizeof(double) * 2 * fs));   // the GEMMs never touch the pad columns
        SEM_CUDA(cudaMemset(c->fB2, 0, sizeof(double) * 2 * fs));
    }
    const double one = 1.0, zero = 0.0;
    SEM_CUBLAS(cublasSetStream(c->blas, st));
    SEM_CUBLAS(cublasDgemmStridedBatched(c->blas, CUBLAS_OP_N, CUBLAS_OP_T, ny, nx, nx, &one, r, ld, stride, c->fQx, nx, 0, &zero,
                                         c->fB1, ld, fs, 2));
    SEM_CUBLAS(cublasDgemmStridedBatched(c->blas, CUBLAS_OP_N, CUBLAS_OP_N, ny, nx, ny, &one, c->fQy, ny, 0, c->fB1, ld, fs, &zero,
                                         c->fB2, ld, fs, 2));
    if (aux_fdm_scale(c->g, c->fLx, c->fLy, c->fB2, nx, st, 2, fs)) return -1;
    SEM_CUBLAS(cublasDgemmStridedBatched(c->blas, CUBLAS_OP_T, CUBLAS_OP_N, ny, nx, ny, &one, c->fQy, ny, 0, c->fB2, ld, fs, &zero,
                                         c->fB1, ld, fs, 2));
    SEM_CUBLAS(cublasDgemmStridedBatched(c->blas, CUBLAS_OP_N, CUBLAS_OP_N, ny, nx, nx, &one, c->fB1, ld, fs, c->fQx, nx, 0, &zero,
                                         z, ld, stride, 2));
    return aux_fdm_boundary(c->g, c->fdm_dir, r, z, st, 2, stride);
}

// ---------------------------------------------------------------------------------------------------------------
// Right-preconditioned restarted GMRES with CGS2 orthogonalisation (two classical Gram-Schmidt passes, each one
// batched dot kernel + one batched update kernel).  Replaces scipy.sparse.linalg.lgmres of CD:146-148 / NS:222-224.
// The stopping rule is the reference's: true residual 2-norm <= atol (right preconditioning keeps the Arnoldi
// residual equal to the true residual; it is re-evaluated explicitly before returning).
// ---------------------------------------------------------------------------------------------------------------
typedef std::function<int(const double*, double*)> vecop;

struct GmresLayout {
    long long n;     // doubles per multi-vector
    int nf;          // fields per multi-vector
    long long vlen;  // doubles per field
};

static int gmres(sem_ctx* c, const GmresLayout& L, const vecop& Aop, const vecop& Pinv, const double* b, double* x,
                 sem_krylov* kr, double* V, double* w, double* t, double* vin, cudaStream_t st, bool use_graph) {
    const long long n = L.n;
    int m = kr->restart;
    if (m < 1) m = 1;
    if (m > SEM_MAX_RESTART) m = SEM_MAX_RESTART;
    std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m), sn(m), gv(m + 1), yv(m);
    kr->iters = 0;
    kr->resnorm = -1.0;
    auto norm2 = [&](const double* v, double* out) -> int {
        if (ctx_multi_dot(c, v, n, 1, v, c->d_small, L.nf, L.vlen, st)) return -1;
        SEM_CUDA(cudaMemcpyAsync(c->h_small, c->d_small, sizeof(double), cudaMemcpyDeviceToHost, st));
        SEM_CUDA(cudaStreamSynchronize(st));
        *out = std::sqrt(c->h_small[0]);
        return 0;
    };
    // r0 = b - A x  -> V0
    if (Aop(x, w)) return -1;
    if (aux_axpby(1.0, b, 0.0, V, n, st)) return -1;
    if (aux_axpby(-1.0, w, 1.0, V, n, st)) return -1;
    double beta;
    if (norm2(V, &beta)) return -1;
    if (kr->verbose) fprintf(stderr, "[sem gmres] start |r| = %.6e  atol = %.3e  n = %lld restart = %d\n", beta, kr->atol, n, m);
    // The fixed part of an iteration, w = A Pinv(vin) (~17 launches for NS), is captured once into a CUDA graph and
    // replayed: on the reference's meshes an iteration is launch bound.  vin is a second copy of the newest basis vector
    // at a fixed address (the graph's kernel arguments never change).
    cudaGraphExec_t gexec = nullptr;
    int eager_applies = 0;
    struct GraphGuard { cudaGraphExec_t* g; ~GraphGuard() { if (*g) cudaGraphExecDestroy(*g); } } guard{&gexec};
    auto apply_fixed = [&]() -> int {
        if (gexec) { SEM_CUDA(cudaGraphLaunch(gexec, st)); return 0; }
        if (!use_graph || eager_applies < 1) {   // the first application also sets every lazily configured kernel attribute
            ++eager_applies;
            if (Pinv(vin, t)) return -1;
            return Aop(t, w);
        }
        cudaGraph_t graph = nullptr;
        SEM_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        const int rc = Pinv(vin, t) || Aop(t, w);
        const cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (rc || ce != cudaSuccess || !graph) {   // not capturable here: carry on without a graph
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            use_graph = false;
            if (Pinv(vin, t)) return -1;
            return Aop(t, w);
        }
        const cudaError_t ie = cudaGraphInstantiate(&gexec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) { gexec = nullptr; cudaGetLastError(); use_graph = false; if (Pinv(vin, t)) return -1; return Aop(t, w); }
        SEM_CUDA(cudaGraphLaunch(gexec, st));
        return 0;
    };
    while (true) {
        kr->resnorm = beta;
        if (!(beta > kr->atol)) return 0;
        if (kr->iters >= kr->max_iters) return kr->iters > 0 ? kr->iters : 1;
        if (aux_axpby(1.0 / beta, V, 0.0, w, n, st)) return -1;
        if (aux_axpby(1.0, w, 0.0, V, n, st)) return -1;
        if (aux_axpby(1.0, w, 0.0, vin, n, st)) return -1;
        gv.assign(m + 1, 0.0);
        gv[0] = beta;
        double est = beta;
        // Arnoldi steps are ENQUEUED up to SEM_GMRES_LAG ahead of the host: the Hessenberg column of step j travels to a
        // pinned ring slot behind an event and is rotated / tested when it arrives, so the GPU never waits for the host
        // between steps.  A step enqueued past the converged one only writes basis vectors that are not used.
        int enq = 0, done = 0;
        bool stop = false;
        auto enqueue_step = [&](int j) -> int {
            if (apply_fixed()) return -1;   // w = A Pinv(V_j), V_j read through its copy vin
            double* h1 = c->d_small;
            double* h2 = c->d_small + (j + 1);
            double* nr = c->d_small + 2 * (j + 1);
            if (ctx_multi_dot(c, V, n, j + 1, w, h1, L.nf, L.vlen, st)) return -1;
            if (aux_multi_axpy(V, n, j + 1, h1, -1.0, w, c->rs, st)) return -1;
            if (ctx_multi_dot(c, V, n, j + 1, w, h2, L.nf, L.vlen, st)) return -1;
            if (aux_multi_axpy(V, n, j + 1, h2, -1.0, w, c->rs, st)) return -1;
            if (ctx_multi_dot(c, w, n, 1, w, nr, L.nf, L.vlen, st)) return -1;
            if (aux_scale_inv_norm(w, nr, V + (long long)(j + 1) * n, vin, n, st)) return -1;
            const int slot = j % SEM_GMRES_LAG;
            SEM_CUDA(cudaMemcpyAsync(c->h_ring + (size_t)slot * c->ring_stride, c->d_small, sizeof(double) * (2 * (j + 1) + 1),
                                     cudaMemcpyDeviceToHost, st));
            SEM_CUDA(cudaEventRecord(c->ev_ring[slot], st));
            return 0;
        };
        while (!stop) {
            const bool can_enqueue = enq < m && enq - done < SEM_GMRES_LAG && kr->iters + enq < kr->max_iters;
            if (can_enqueue) {
                if (enqueue_step(enq)) return -1;
                ++enq;
                if (enq - done < SEM_GMRES_LAG && enq < m && kr->iters + enq < kr->max_iters) continue;
            }
            if (done == enq) break;   // basis full or iteration cap reached, everything processed
            const int j = done;
            const int slot = j % SEM_GMRES_LAG;
            SEM_CUDA(cudaEventSynchronize(c->ev_ring[slot]));
            const double* hs = c->h_ring + (size_t)slot * c->ring_stride;
            double* Hj = &H[(size_t)j * (m + 1)];   // column j
            for (int i = 0; i <= j; ++i) Hj[i] = hs[i] + hs[j + 1 + i];
            const double hn = std::sqrt(hs[2 * (j + 1)]);
            Hj[j + 1] = hn;
            for (int i = 0; i < j; ++i) {
                const double a = cs[i] * Hj[i] + sn[i] * Hj[i + 1];
                Hj[i + 1] = -sn[i] * Hj[i] + cs[i] * Hj[i + 1];
                Hj[i] = a;
            }
            const double d = std::hypot(Hj[j], Hj[j + 1]);
            cs[j] = d > 0 ? Hj[j] / d : 1.0;
            sn[j] = d > 0 ? Hj[j + 1] / d : 0.0;
            Hj[j] = d;
            Hj[j + 1] = 0.0;
            gv[j + 1] = -sn[j] * gv[j];
            gv[j] = cs[j] * gv[j];
            est = std::fabs(gv[j + 1]);
            ++done;
            if (kr->verbose > 1 || (kr->verbose && (kr->iters + done) % 100 == 0))
                fprintf(stderr, "[sem gmres] it %d  |r| ~ %.6e\n", kr->iters + done, est);
            if (!std::isfinite(est)) { set_error("gmres: non-finite residual"); return -3; }
            if (est <= kr->atol || !(hn > 0.0)) stop = true;
        }
        kr->iters += done;   // Arnoldi steps that enter the solution
        const int j = done;
        // y = H^-1 g (back substitution on the rotated upper-triangular H), x += Pinv(V y)
        const int k = j;
        for (int i = k - 1; i >= 0; --i) {
            double s = gv[i];
            for (int q = i + 1; q < k; ++q) s -= H[(size_t)q * (m + 1) + i] * yv[q];
            yv[i] = s / H[(size_t)i * (m + 1) + i];
        }
        std::memcpy(c->h_small, yv.data(), sizeof(double) * k);
        SEM_CUDA(cudaMemcpyAsync(c->d_small, c->h_small, sizeof(double) * k, cudaMemcpyHostToDevice, st));
        if (aux_multi_comb(V, n, k, c->d_small, w, c->rs, st)) return -1;
        if (Pinv(w, t)) return -1;
        if (aux_axpby(1.0, t, 1.0, x, n, st)) return -1;
        // true residual
        if (Aop(x, w)) return -1;
        if (aux_axpby(1.0, b, 0.0, V, n, st)) return -1;
        if (aux_axpby(-1.0, w, 1.0, V, n, st)) return -1;
        if (norm2(V, &beta)) return -1;
        if (kr->verbose) fprintf(stderr, "[sem gmres] cycle end: its %d  true |r| = %.6e (estimate %.3e)\n", kr->iters, beta, est);
    }
}

static int ensure_kdiag(sem_ctx* c, cudaStream_t st) {
    if (c->dKdiag) return 0;
    SEM_CUDA(cudaMalloc(&c->dKdiag, sizeof(double) * (size_t)c->g.NX * c->g.LD));
    return aux_stiffness_diag(c->g, c->tab(), c->dKdiag, st);
}

// The Krylov solvers run on a context-owned non-blocking stream (a CUDA graph cannot be captured on the legacy default
// stream a caller may hand in); it is ordered after the caller's stream on entry and drained before returning.  With a
// communicator (NCCL calls, side streams) the caller's stream is used directly and no graph is built.
static int solve_stream(sem_ctx* c, cudaStream_t user, cudaStream_t* st, bool* use_graph) {
    static const bool no_graph = std::getenv("SEM_B200_NO_GRAPH") != nullptr;
    *use_graph = !c->has_comm && !no_graph;
    *st = user;
    if (!*use_graph) return 0;
    if (ensure_streams(c)) return -1;
    SEM_CUDA(cudaEventRecord(c->ev_start, user));
    SEM_CUDA(cudaStreamWaitEvent(c->s_solve, c->ev_start, 0));
    *st = c->s_solve;
    return 0;
}

extern "C" long long sem_cd_work_len(const sem_ctx* c, int restart) {
    if (!c) return -1;
    return (long long)(restart + 1 + 3) * c->g.NX * c->g.LD;
}

extern "C" int sem_cd_solve(sem_ctx* c, const sem_cd_state* s, const double* rhs, double* dT, sem_krylov* kr,
                            double* work, long long work_len, void* stream) {
    SEM_CHECK_CTX(c);
    cudaStream_t st;
    bool use_graph;
    if (solve_stream(c, (cudaStream_t)stream, &st, &use_graph)) return -1;
    const long long vlen = (long long)c->g.NX * c->g.LD;
    if (kr->restart > SEM_MAX_RESTART) kr->restart = SEM_MAX_RESTART;
    if (work_len < sem_cd_work_len(c, kr->restart)) { set_error("sem_cd_solve: work buffer too small"); return -2; }
    if (ensure_kdiag(c, st)) return -1;
    double* V = work;
    double* w = work + (long long)(kr->restart + 1) * vlen;
    double* t = w + vlen;
    double* vin = t + vlen;
    SEM_CUDA(cudaMemsetAsync(w, 0, sizeof(double) * 3 * vlen, st));
    sem_cd_state lin = *s;
    BCSpec bc;
    fill_cd_bc(bc, s->bc, 0);
    vecop Aop = [&](const double* xx, double* yy) { return sem_cd_jvp(c, &lin, xx, nullptr, nullptr, yy, (void*)st); };
    vecop Pinv = [&](const double* r, double* z) {
        if (kr->precond == 0) return aux_axpby(1.0, r, 0.0, z, vlen, st);
        if (kr->precond == 2) return fdm_apply(c, r, z, st);
        return aux_cd_jacobi(c->g, bc, c->dKdiag, r, z, st);
    };
    GmresLayout L{vlen, 1, vlen};
    const int rc = gmres(c, L, Aop, Pinv, rhs, dT, kr, V, w, t, vin, st, use_graph);
    SEM_CUDA(cudaStreamSynchronize(st));
    return rc;
}

extern "C" long long sem_ns_work_len(const sem_ctx* c, int restart) {
    if (!c) return -1;
    return ((long long)(restart + 1 + 3) * 3 + 1) * c->g.NX * c->g.LD;
}

extern "C" int sem_ns_solve(sem_ctx* c, const sem_ns_state* s, const double* rhs3, double* x3, sem_krylov* kr,
                            double* work, long long work_len, void* stream) {
    SEM_CHECK_CTX(c);
    cudaStream_t st;
    bool use_graph;
    if (solve_stream(c, (cudaStream_t)stream, &st, &use_graph)) return -1;
    const long long vlen = (long long)c->g.NX * c->g.LD;
    const long long n = 3 * vlen;
    if (kr->restart > SEM_MAX_RESTART) kr->restart = SEM_MAX_RESTART;
    if (work_len < sem_ns_work_len(c, kr->restart)) { set_error("sem_ns_solve: work buffer too small"); return -2; }
    if (ensure_kdiag(c, st)) return -1;
    double* V = work;
    double* w = work + (long long)(kr->restart + 1) * n;
    double* t = w + n;
    double* vin = t + n;
    double* tmp = vin + n;
    SEM_CUDA(cudaMemsetAsync(w, 0, sizeof(double) * (3 * n + vlen), st));
    sem_ns_state lin = *s;
    vecop Aop = [&](const double* xx, double* yy) {
        return sem_ns_jvp(c, &lin, xx, xx + vlen, xx + 2 * vlen, nullptr, yy, yy + vlen, yy + 2 * vlen, (void*)st);
    };
    // Block lower-triangular right preconditioner  [[P_a, 0], [C, M_p]]^-1  with P_a = Jacobi on the velocity block,
    // C = continuity rows, M_p = diagonal mass with the pin row passed through (the reference's Schur preconditioner,
    // NS:208-212).  With this structure GMRES converges to the same member of the (singular, consistent) system's
    // solution family as the reference's Schur-complement iteration -- see DESIGN.md.
    vecop Pinv = [&](const double* r, double* z) {
        if (kr->precond == 0) return aux_axpby(1.0, r, 0.0, z, n, st);
        if (kr->precond == 2 || kr->precond == 3) {
            if (fdm_apply2(c, r, z, vlen, st)) return -1;
        } else if (aux_ns_jacobi(c->g, c->dKdiag, lin.gxu, lin.gyv, r, r + vlen, z, z + vlen, st)) {
            return -1;
        }
        MarchArgs A = zero_args();
        A.a = z; A.b = z + vlen; A.y0 = tmp;
        if (apply_and_exchange(c, MODE_DIV, A, {tmp}, st)) return -1;
        if (aux_ns_schur_mass(c->g, c->tab(), r + 2 * vlen, tmp, z + 2 * vlen, c->pin_gx, c->pin_iy, st)) return -1;
        if (kr->precond == 3) return pbb_apply(c, r + 2 * vlen, z + 2 * vlen, tmp, st);   // experimental, see sem_ctx_set_pbb
        return 0;
    };
    GmresLayout L{n, 3, vlen};
    const int rc = gmres(c, L, Aop, Pinv, rhs3, x3, kr, V, w, t, vin, st, use_graph);
    SEM_CUDA(cudaStreamSynchronize(st));
    return rc;
}
