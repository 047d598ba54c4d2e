// C ABI of sem_b200 (see include/sem_b200.h): context, host<->device packing, fused operators, Krylov solvers.
#include "../../include/sem_b200.h"
#include "sem_aux.cuh"
#include "sem_comm.cuh"
#include "sem_dispatch.h"
#include "sem_march.cuh"
#include "sem_march3.cuh"
#include "sem_gemm.cuh"
#include "sem_schur.cuh"
#include <cstdlib>

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

#define SEM_FDM_SLOTS 3
enum { SEM_FDM_MAIN = 0, SEM_FDM_NEUMANN = 1, SEM_FDM_COARSE = 2 };
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
    long long n_fused, n_split;   // partitioned applies so far: one-launch (in-kernel exchange) / three launches + exchange kernel
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
    // fast-diagonalisation plans (sem_ctx_set_fdm), see sem_gemm.cuh.  Slots: SEM_FDM_MAIN the Laplacian with the solver's
    // Dirichlet sides (CD operator / NS velocity block), SEM_FDM_NEUMANN the all-Neumann pressure Laplacian (pseudo-inverse),
    // SEM_FDM_COARSE the structured coarse operator of the NS Schur complement.  One GPU: FdmPlan; partitioned: FdmDist.
    FdmPlan plan[SEM_FDM_SLOTS];
    struct FdmDist {
        int ready, outside;
        int nmodes, B, modesP;     // global x modes, modes per rank (multiple of 128), B * world
        int s1, k1, k1p;           // lines [s1, s1 + k1) of the slab enter the forward transform (k1p: padded to 16)
        int s4, m4, m4p;           // lines [s4, s4 + m4) of the slab are produced by the backward transform (m4p: padded to 128)
        double *QxT, *Qx, *lamx;   // [modesP][k1p], [m4p][modesP], [modesP]
        FdmDir y;
        int cols;
        double *R, *T, *mA, *mB, *X;
        double den_floor;
    } dist[SEM_FDM_SLOTS];
    // Navier-Stokes Schur-complement preconditioner (sem_ctx_set_ns_schur): projector tables, null-vector factors, work vectors
    struct NsSchur {
        int ready, singular, two_level;
        double inv_den;                  // 1 / (m_c . l_c)
        int cheb_steps;
        double cheb_inv_theta, cheb_a[8], cheb_b[8];
        PwDir px, py;
        double* tables;                  // one device block holding wl, wr and the null-vector factors
        double *TinvX, *TinvY;           // padded inverses of the projector Gram matrices: [kxp][kxp], [kyp][kyp]
        int tile, kxp, kyp, ldx, rowsy;  // GEMM tile; padded vertex counts; pitch of cx; padded rows of cy
        double *lc, *mc;                 // l_c = lfx (x) lfy and m_c = M_p l_c as vectors
        double* w[10];                   // work vectors (zero initialised): Y, Z1, R1, G0, G1, V0, V1, NR, RHO, E
        double *cx, *cx2, *cy, *cy2;     // projector coefficients before / after T^-1: [kxp][ldx] and [rowsy][kyp]
        double* sums;                    // device [2]
    } sch;
    TabDev tab() const { return TabDev{dD, dKs, dw}; }
};

static void schur_free(sem_ctx::NsSchur& q) {
    if (q.tables) cudaFree(q.tables);
    if (q.lc) cudaFree(q.lc);
    if (q.mc) cudaFree(q.mc);
    for (double* p : q.w)
        if (p) cudaFree(p);
    for (double* p : {q.cx, q.cx2, q.cy, q.cy2, q.TinvX, q.TinvY})
        if (p) cudaFree(p);
    if (q.sums) cudaFree(q.sums);
    std::memset(&q, 0, sizeof(q));
}

static void fdm_dist_free(sem_ctx::FdmDist& d) {
    if (d.QxT) { cudaFree(d.QxT); cudaFree(d.Qx); cudaFree(d.lamx); cudaFree(d.R); cudaFree(d.T); cudaFree(d.mA); cudaFree(d.mB); cudaFree(d.X); }
    fdm_dir_free(d.y);
    std::memset(&d, 0, sizeof(d));
}

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
    for (int k = 0; k < SEM_FDM_SLOTS; ++k) {
        fdm_plan_free(c->plan[k]);
        fdm_dist_free(c->dist[k]);
    }
    schur_free(c->sch);
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
    if (comm_init(c->comm, id128, rank, world, c->g.NY, c->g.nex)) return -1;
    c->has_comm = 1;
    return 0;
}

// One-GPU self-test of the peer-memory exchange paths: the context must describe an inner slab (m_begin > 0, m_end < N_ex);
// its two interface lines are then exchanged with each other through the context's own mailbox.
extern "C" int sem_ctx_attach_loopback(sem_ctx* c) {
    SEM_CHECK_CTX(c);
    if (c->has_comm) { set_error("sem_ctx_attach_loopback: communicator already attached"); return -2; }
    if (!c->g.has_left || !c->g.has_right) { set_error("sem_ctx_attach_loopback: the context must be an inner slab"); return -2; }
    if (comm_init_loopback(c->comm, c->g.NY, c->g.nex)) return -1;
    c->has_comm = 1;
    return 0;
}

extern "C" long long sem_ctx_partitioned_applies(const sem_ctx* c, int fused) {
    if (!c) return -1;
    return fused ? c->n_fused : c->n_split;
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

// The whole partitioned apply as ONE launch (no post-operator, peer-memory mailboxes, v3 kernel): the chunks that finish the
// interface lines come first in the grid (two element columns each: a priming phase and two marching steps), push their
// segments of those lines into the neighbours' mailboxes when their chunk is done, and then add the neighbour's
// contribution, which its own edge CTAs -- first in its grid too -- have normally delivered by then.  Every
// push of a CTA precedes its waits and a pushing CTA waits for nothing, so the scheme cannot deadlock as long as the edge
// CTAs of a launch are co-resident (checked: 2 x strips <= half the resident slots; else the three-launch path is used).
// SEM_B200_FUSED_XCH=0 switches the path off for A/B runs.
static const bool SEM_FUSED_XCH = [] {
    const char* e = std::getenv("SEM_B200_FUSED_XCH");
    return !(e && e[0] == '0');
}();

static bool large_interior(const sem_ctx* c, int interior_columns);

static bool fused_apply_possible(const sem_ctx* c, int mode, int n_fields) {
    if (!SEM_FUSED_XCH || !c->has_comm || !c->comm.p2p || march_generation() != 3 || mode == MODE_NS) return false;
    if (n_fields < 1 || n_fields > c->comm.max_fields) return false;
    const auto& ord = g_orders[c->g.P - 1];
    const MarchGeom q = march3_geometry(c->g, mode, c->Mx_req, c->sm_count, ord.smem3(mode), (size_t)c->smem_sm, 0, c->g.nex);
    const int strips = (int)q.grid.x;
    int resident = (int)((size_t)c->smem_sm / (ord.smem3(mode) + 1024));
    if (resident > 32) resident = 32;
    if (4 * strips > c->sm_count * resident) return false;
    // Neighbours must take the same path (the one-launch exchange and the exchange kernels use different mailbox regions and
    // flags): the choice uses only numbers that are the same on every rank -- the narrowest and the widest slab of the
    // partition, and two edge chunks of two columns whether this rank has one neighbour or two.
    if (c->comm.min_nex <= 4) return false;   // some rank would have no interior
    // A large interior (several resident rounds of CTAs) hides the whole exchange of the three-launch path, and the edge CTAs
    // of the one-launch path then only lose time waiting for the neighbour while they hold 2 x strips resident slots:
    // measured at 2 GPUs on config 5 (512-column slabs) 0.205 ms fused against 0.199 ms split; 4 GPUs (256 columns) 0.109
    // against 0.112 ms; the 128-column slab of 8 GPUs takes 66.5 us fused against 71.3 us split (one-GPU loopback).
    // SEM_B200_FUSED_XCH=2 forces the one-launch path.
    static const bool force = [] { const char* e = std::getenv("SEM_B200_FUSED_XCH"); return e && e[0] == '2'; }();
    return force || !large_interior(c, c->comm.max_nex - 4);
}

static bool large_interior(const sem_ctx* c, int interior_columns) {
    const long long interior_ctas = (long long)(c->g.ney / 8 + 1) * ((interior_columns + 15) / 16);
    return interior_ctas >= 3ll * 6 * c->sm_count;
}

static int fused_apply(sem_ctx* c, int mode, MarchArgs& A, cudaStream_t st) {
    A.zero = 0;
    const auto& ord = g_orders[c->g.P - 1];
    const int nex = c->g.nex;
    MarchGeom q = march3_geometry(c->g, mode, c->Mx_req, c->sm_count, ord.smem3(mode), (size_t)c->smem_sm, 0, nex);
    XchArgs X;
    if (comm_fill_xch(c->comm, c->g, X)) return -1;
    const int el = c->g.has_left ? 2 : 0, er = c->g.has_right ? 2 : 0;
    if (el) { X.e_lo[X.nedge] = 0; X.e_hi[X.nedge] = el; X.nedge++; }
    if (er) { X.e_lo[X.nedge] = nex - er; X.e_hi[X.nedge] = nex; X.nedge++; }
    q.m_lo = el;
    q.m_hi = nex - er;
    q.grid.y = (unsigned)(X.nedge + (q.m_hi - q.m_lo + q.Mx - 1) / q.Mx);
    q.xch = &X;
    c->n_fused++;
    return ord.launch3(mode, c->g, A, q, st);
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
    const bool interior_first = large_interior(c, nex - el - er);
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
    if (!post && fused_apply_possible(c, mode, n)) return fused_apply(c, mode, A, st);   // the fields ARE the kernel's non-null outputs
    const int nex = c->g.nex;
    const int el = c->g.has_left ? std::min(SEM_EDGE_COLUMNS, nex) : 0;
    const int er = c->g.has_right ? std::min(SEM_EDGE_COLUMNS, nex - el) : 0;
    if (el + er >= nex || el + er == 0 || n == 0) {   // nothing (left) to overlap with
        if (march(c, mode, A, st)) return -1;
        if (post && post(st)) return -1;
        return n ? comm_exchange_add(c->comm, c->g, f, n, st) : 0;
    }
    if (ensure_streams(c)) return -1;
    c->n_split++;
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

// SEM.eval_interpolation (SEM.py:248-273) of a device vector on an ij-meshgrid: mx[nxp] / ny[nyp] = element of every plot
// column / row (x2xi, SEM.py:23-36), Sx [nxp][P+1] / Sy [nyp][P+1] = GLL.standard_evaluation_matrix (GLL.py:105-116), all
// DEVICE arrays; out [nxp][nyp] DEVICE.  On a partitioned mesh every rank evaluates the plot columns inside its slab and the
// arrays are summed over the ranks (ncclAllReduce): every rank returns the whole result.
extern "C" int sem_interpolate(sem_ctx* c, const double* vec, int nxp, const int* mx, const double* Sx, int nyp, const int* ny,
                               const double* Sy, double* out, void* stream) {
    SEM_CHECK_CTX(c);
    cudaStream_t st = (cudaStream_t)stream;
    if (aux_interpolate(c->g, vec, nxp, mx, Sx, nyp, ny, Sy, out, nyp, st)) return -1;
    if (c->has_comm) return comm_allreduce_sum(c->comm, out, nxp * nyp, st);
    return 0;
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
// Q^T M1 Q = I, Q^T K1 Q = diag(l).  The four transforms are dense fp64 products on the DMMA kernel of sem_gemm.cu; a
// direction whose pencil is symmetric about the domain centre (both ends Dirichlet or both Neumann) is folded into an even
// and an odd half-size product (half the flops); the spectral scaling is the epilogue of the second product.
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_pad_rows(const double* __restrict__ src, int src_ld, int row0, int nrows, int ncols, double* __restrict__ dst,
                           int dst_ld, int transpose) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (i >= nrows || j >= ncols) return;
    const double v = src[(size_t)(row0 + i) * src_ld + j];
    if (transpose) dst[(size_t)j * dst_ld + i] = v;
    else dst[(size_t)i * dst_ld + j] = v;
}

static double fdm_den_floor(const MeshDev& g) {
    // lx + ly = 0 only for the constant mode of an all-Neumann problem; the largest eigenvalue is ~ P^4 (1/dx^2 + 1/dy^2)
    // and the eigen-solver's absolute error ~ 1e-16 of that
    const double p4 = (double)g.P * g.P * g.P * g.P;
    return 1e-13 * p4 * (1.0 / (g.dx * g.dx) + 1.0 / (g.dy * g.dy));
}

extern "C" int sem_ctx_set_fdm(sem_ctx* c, int slot, const sem_fdm_dir* x, const sem_fdm_dir* y, int outside, double den_floor) {
    SEM_CHECK_CTX(c);
    if (slot < 0 || slot >= SEM_FDM_SLOTS || !x || !y) { set_error("sem_ctx_set_fdm: bad slot / null direction"); return -2; }
    if (!c->has_comm && (c->g.has_left || c->g.has_right)) {
        set_error("sem_ctx_set_fdm: a partitioned context needs its communicator first (sem_ctx_attach_comm)");
        return -2;
    }
    if (den_floor < 0) den_floor = fdm_den_floor(c->g);
    const MeshDev& g = c->g;
    if (!c->has_comm) return fdm_plan_build(c->plan[slot], g, x->lo, x->cnt, x->fold, x->Qe, x->Qo, x->lam, y->lo, y->cnt, y->fold,
                                            y->Qe, y->Qo, y->lam, outside, 2, den_floor);
    // partitioned: x is the GLOBAL pencil, unfolded; x->Qe = the slab's rows [NX][nmodes] of the global Q (zero rows at
    // eliminated end nodes), x->lam[nmodes]; x->lo / x->cnt = the lines of the SLAB the operator acts on.
    if (x->fold) { set_error("sem_ctx_set_fdm: the x direction of a partitioned mesh is not folded"); return -2; }
    sem_ctx::FdmDist& d = c->dist[slot];
    fdm_dist_free(d);
    const int world = c->comm.world;
    d.outside = outside;
    d.den_floor = den_floor;
    d.nmodes = x->nmodes;
    d.B = round_up((d.nmodes + world - 1) / world, 128);
    d.modesP = d.B * world;
    d.s4 = x->lo;
    d.m4 = x->cnt;
    d.m4p = round_up(d.m4, 128);
    d.s1 = std::max(x->lo, g.has_left ? 1 : 0);               // the interface line is summed from the left rank's copy
    d.k1 = x->lo + x->cnt - d.s1;
    d.k1p = round_up(d.k1, 16);
    if (fdm_dir_build(d.y, y->lo, y->cnt, y->fold, y->Qe, y->Qo, y->lam, 128)) return -1;
    d.cols = std::max(round_up(y->cnt, 128), d.y.nep + d.y.nop);
    auto zalloc = [](double** p, size_t n) -> int {
        SEM_CUDA(cudaMalloc(p, sizeof(double) * n));
        SEM_CUDA(cudaMemset(*p, 0, sizeof(double) * n));
        return 0;
    };
    if (zalloc(&d.QxT, (size_t)d.modesP * d.k1p) || zalloc(&d.Qx, (size_t)d.m4p * d.modesP) || zalloc(&d.lamx, d.modesP) ||
        zalloc(&d.R, (size_t)d.k1p * d.cols) || zalloc(&d.T, (size_t)d.modesP * d.cols) || zalloc(&d.mA, (size_t)d.B * d.cols) ||
        zalloc(&d.mB, (size_t)d.B * d.cols) || zalloc(&d.X, (size_t)d.m4p * d.cols))
        return -1;
    const dim3 blk(128);
    k_pad_rows<<<dim3((unsigned)((d.nmodes + 127) / 128), (unsigned)d.k1), blk>>>(x->Qe, d.nmodes, d.s1, d.k1, d.nmodes, d.QxT, d.k1p, 1);
    k_pad_rows<<<dim3((unsigned)((d.nmodes + 127) / 128), (unsigned)d.m4), blk>>>(x->Qe, d.nmodes, d.s4, d.m4, d.nmodes, d.Qx, d.modesP, 0);
    SEM_CUDA(cudaGetLastError());
    SEM_CUDA(cudaMemcpy(d.lamx, x->lam, sizeof(double) * d.nmodes, cudaMemcpyDeviceToDevice));
    SEM_CUDA(cudaDeviceSynchronize());
    d.ready = 1;
    return 0;
}

// Partitioned mesh: the x transform couples all slabs, so it is distributed -- every rank forms its contribution
// Qx_r^T R_r to all modes (one product over its own lines; the duplicated interface line is taken from the left rank only), a
// reduce-scatter over NVLink leaves each rank with a block of modes, the y transforms and the spectral scaling are local to
// that block, an all-gather returns all modes and one product with the slab's rows of Qx gives the slab of the result.  The
// preconditioner is the exact inverse on any number of GPUs: the Krylov iteration count does not depend on the partition.
static int fdm_dist_apply(sem_ctx* c, sem_ctx::FdmDist& d, const double* r, double* z, cudaStream_t st) {
    const MeshDev& g = c->g;
    const int ycols = round_up(d.y.cnt, 128);
    FdmDir x1, x4;
    std::memset(&x1, 0, sizeof(x1));
    x1.lo = d.s1; x1.cnt = d.k1; x1.ne = d.k1;
    x4 = x1;
    x4.lo = d.s4; x4.cnt = d.m4; x4.ne = d.m4;
    if (fdm_fold_x(x1, d.y.lo, d.y.cnt, r, 0, g.LD, d.R, d.cols, 0, 1, st)) return -1;
    GemmArgs a;
    std::memset(&a, 0, sizeof(a));
    a.nprob = 1; a.batch = 1; a.tile = 128;
    a.p[0] = GemmProblem{d.QxT, d.R, d.T, d.k1p, d.cols, d.cols, 0, 0, 0, d.modesP, ycols, d.k1p, nullptr, nullptr};
    if (gemm_launch(a, EPI_NONE, st)) return -1;
    if (comm_reduce_scatter_sum(c->comm, d.T, d.mA, (size_t)d.B * d.cols, st)) return -1;
    if (fdm_fold_y(d.y, d.mA, d.mB, d.B, d.cols, 0, 1, st)) return -1;
    if (fdm_step_y(d.y, false, d.mB, d.mA, d.B, d.cols, 0, d.lamx + (size_t)c->comm.rank * d.B, d.den_floor, 1, 128, st)) return -1;
    if (fdm_step_y(d.y, true, d.mA, d.mB, d.B, d.cols, 0, nullptr, 0.0, 1, 128, st)) return -1;
    if (fdm_unfold_y(d.y, d.mB, d.mA, d.B, d.cols, 0, 1, st)) return -1;
    if (comm_allgather(c->comm, d.mA, d.T, (size_t)d.B * d.cols, st)) return -1;
    a.p[0] = GemmProblem{d.Qx, d.T, d.X, d.modesP, d.cols, d.cols, 0, 0, 0, d.m4p, ycols, d.modesP, nullptr, nullptr};
    if (gemm_launch(a, EPI_NONE, st)) return -1;
    return fdm_unfold_x(x4, d.y.lo, d.y.cnt, d.X, d.cols, 0, r, z, 0, g, d.outside, 1, st);
}

// nf fields `stride` doubles apart; r and z may alias
static int fdm_apply(sem_ctx* c, int slot, const double* r, double* z, int nf, long long stride, cudaStream_t st) {
    if (!c->has_comm) {
        if (!c->plan[slot].ready) { set_error("fdm_apply: sem_ctx_set_fdm has not been called for this slot"); return -2; }
        return fdm_plan_apply(c->plan[slot], c->g, r, z, nf, stride, st);
    }
    if (!c->dist[slot].ready) { set_error("fdm_apply: sem_ctx_set_fdm has not been called for this slot"); return -2; }
    for (int f = 0; f < nf; ++f)
        if (fdm_dist_apply(c, c->dist[slot], r + f * stride, z + f * stride, st)) return -1;
    return 0;
}

extern "C" int sem_fdm_apply(sem_ctx* c, int slot, const double* r, double* z, int nf, void* stream) {
    SEM_CHECK_CTX(c);
    if (slot < 0 || slot >= SEM_FDM_SLOTS || nf < 1 || nf > 2) { set_error("sem_fdm_apply: bad slot / field count"); return -2; }
    return fdm_apply(c, slot, r, z, nf, (long long)c->g.NX * c->g.LD, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------------------
// Navier-Stokes preconditioner: block lower-triangular  [[P_a, 0], [C, S~]]^-1  for the 3-field Jacobian of NS:138-160.
//   P_a  = Laplacian inverse per velocity component (fast diagonalisation; identity rows on the walls), or Jacobi
//   C    = continuity rows,  S~ = approximation of the pressure Schur complement, by `level` (sem_krylov.precond):
//     1, 2  z_p = y / M_p                        the reference's own diagonal-mass preconditioner (NS:208-212)
//     3     + boundary ring by block elimination  z_B = K_BB^-1 (y_B - K_BI z_I), K_BB^-1 = fixed Chebyshev polynomial
//     4     two-level:  z1 = Pi Shat^+ Pi^T y  (structured coarse operator on the near-null space of the equal-order Schur
//           complement),  r1 = y - S_0 z1,  z2 = M^-1 F_p K_N^+ r1  (pressure convection-diffusion),
//           both followed by the ring elimination; then the rank-one correction that keeps GMRES on the reference's member
//           of the singular system's solution set.
// Every stage is restated on the CPU by oracle/ns_precond.py (test infrastructure) and compared stage by stage in
// tests/test_gpu_parity.py.
// ---------------------------------------------------------------------------------------------------------------
static int ensure_kdiag(sem_ctx* c, cudaStream_t st);

enum { SW_Y = 0, SW_Z1, SW_R1, SW_G0, SW_G1, SW_V0, SW_V1, SW_NR, SW_RHO, SW_E };

extern "C" int sem_ctx_set_ns_schur(sem_ctx* c, const sem_ns_schur_desc* d) {
    SEM_CHECK_CTX(c);
    if (!d) { set_error("sem_ctx_set_ns_schur: null descriptor"); return -2; }
    if (d->cheb_steps < 1 || d->cheb_steps > 8) { set_error("sem_ctx_set_ns_schur: 1..8 Chebyshev steps"); return -2; }
    sem_ctx::NsSchur& q = c->sch;
    schur_free(q);
    const MeshDev& g = c->g;
    const int n1 = g.P + 1, nexg = (g.NXg - 1) / g.P, ney = g.ney;
    const size_t vlen = (size_t)g.NX * g.LD;
    // tables: wl[n1] wr[n1] | lfx[NXg] | lfy[NY]
    const size_t ntab = 2 * n1 + g.NXg + g.NY;
    std::vector<double> h(ntab);
    size_t o = 0;
    auto put = [&](const double* src, size_t n) { size_t at = o; std::memcpy(h.data() + o, src, sizeof(double) * n); o += n; return at; };
    const size_t o_wl = put(d->wl, n1), o_wr = put(d->wr, n1);
    const size_t o_lx = put(d->lfx, g.NXg), o_ly = put(d->lfy, g.NY);
    SEM_CUDA(cudaMalloc(&q.tables, sizeof(double) * ntab));
    SEM_CUDA(cudaMemcpy(q.tables, h.data(), sizeof(double) * ntab, cudaMemcpyHostToDevice));
    q.px = PwDir{g.NXg, q.tables + o_wl, q.tables + o_wr};
    q.py = PwDir{g.NY, q.tables + o_wl, q.tables + o_wr};
    auto zalloc = [](double** p, size_t n) -> int {
        SEM_CUDA(cudaMalloc(p, sizeof(double) * n));
        SEM_CUDA(cudaMemset(*p, 0, sizeof(double) * n));
        return 0;
    };
    // G0|G1 and V0|V1 are pairs of adjacent vectors (two-field FDM applies)
    if (zalloc(&q.w[SW_G0], 2 * vlen) || zalloc(&q.w[SW_V0], 2 * vlen)) return -1;
    for (int k : {SW_Y, SW_Z1, SW_R1, SW_NR, SW_RHO, SW_E})
        if (zalloc(&q.w[k], vlen)) return -1;
    if (zalloc(&q.lc, vlen) || zalloc(&q.mc, vlen) || zalloc(&q.sums, 2)) return -1;
    // projector coefficients c = W^T (M v) -> T^-1 c: the small dense inverse is applied by the DMMA GEMM (a Thomas sweep
    // with one thread per line is a latency-bound chain: 9 % of an iteration at a million nodes in the round-2 launch list)
    q.tile = fdm_tile_for(std::min(g.NXg, g.NY) - 2);
    q.kxp = round_up(nexg + 1, q.tile);
    q.kyp = round_up(ney + 1, q.tile);
    q.ldx = round_up(g.NY, q.tile);
    q.rowsy = round_up(g.NX, q.tile);
    if (zalloc(&q.cx, (size_t)q.kxp * q.ldx) || zalloc(&q.cx2, (size_t)q.kxp * q.ldx) || zalloc(&q.cy, (size_t)q.rowsy * q.kyp) ||
        zalloc(&q.cy2, (size_t)q.rowsy * q.kyp) || zalloc(&q.TinvX, (size_t)q.kxp * q.kxp) || zalloc(&q.TinvY, (size_t)q.kyp * q.kyp))
        return -1;
    SEM_CUDA(cudaMemcpy2D(q.TinvX, sizeof(double) * q.kxp, d->Tinv_x, sizeof(double) * (nexg + 1), sizeof(double) * (nexg + 1),
                          nexg + 1, cudaMemcpyHostToDevice));
    SEM_CUDA(cudaMemcpy2D(q.TinvY, sizeof(double) * q.kyp, d->Tinv_y, sizeof(double) * (ney + 1), sizeof(double) * (ney + 1), ney + 1,
                          cudaMemcpyHostToDevice));
    const Regions rg{c->pin_gx, c->pin_iy};
    if (member_vectors(g, c->tab(), rg, q.tables + o_lx, q.tables + o_ly, q.lc, q.mc, 0)) return -1;
    SEM_CUDA(cudaDeviceSynchronize());
    q.singular = d->singular;
    q.two_level = d->two_level;
    q.inv_den = d->inv_den;
    // Chebyshev iteration for the diagonally scaled ring block with spectrum in [lo, hi]
    const double theta = 0.5 * (d->cheb_hi + d->cheb_lo), delta = 0.5 * (d->cheb_hi - d->cheb_lo);
    const double sigma = theta / delta;
    double rho = 1.0 / sigma;
    q.cheb_steps = d->cheb_steps;
    q.cheb_inv_theta = 1.0 / theta;
    for (int k = 0; k < d->cheb_steps; ++k) {
        const double rho_new = 1.0 / (2.0 * sigma - rho);
        q.cheb_a[k] = rho_new * rho;
        q.cheb_b[k] = 2.0 * rho_new / delta;
        rho = rho_new;
    }
    // w[SW_G1], w[SW_V1]: second halves of the pairs (not separately owned)
    q.w[SW_G1] = nullptr;
    q.w[SW_V1] = nullptr;
    q.ready = 1;
    return 0;
}

struct SchurWork {
    double *Y, *Z1, *R1, *G0, *G1, *V0, *V1, *NR, *RHO, *E;
};
static SchurWork schur_work(const sem_ctx* c) {
    const sem_ctx::NsSchur& q = c->sch;
    const size_t vlen = (size_t)c->g.NX * c->g.LD;
    return SchurWork{q.w[SW_Y], q.w[SW_Z1], q.w[SW_R1], q.w[SW_G0], q.w[SW_G0] + vlen, q.w[SW_V0], q.w[SW_V0] + vlen,
                     q.w[SW_NR], q.w[SW_RHO], q.w[SW_E]};
}

// rows K[ring,:] x of a vector, completed across partition interfaces (the interface lines are cleared first: the kernel
// writes ring nodes only and the exchange adds whole lines)
static int ring_rows(sem_ctx* c, const double* x, double* q, cudaStream_t st) {
    const MeshDev& g = c->g;
    if (c->has_comm) {
        if (g.has_left) SEM_CUDA(cudaMemsetAsync(q, 0, sizeof(double) * g.LD, st));
        if (g.has_right) SEM_CUDA(cudaMemsetAsync(q + (size_t)(g.NX - 1) * g.LD, 0, sizeof(double) * g.LD, st));
    }
    if (aux_neumann_rows(g, c->tab(), x, q, c->pin_gx, c->pin_iy, 1, st)) return -1;
    return exchange(c, {q}, st);
}

// z (given on the inner nodes and the pin, zero on the ring)  ->  z_B = K_BB^-1 (y_B - K_BI z_I) on the ring
static int schur_add_ring(sem_ctx* c, double* z, const double* y, cudaStream_t st) {
    sem_ctx::NsSchur& q = c->sch;
    const SchurWork W = schur_work(c);
    const Regions rg{c->pin_gx, c->pin_iy};
    if (ring_rows(c, z, W.NR, st)) return -1;
    if (ring_init(c->g, rg, y, W.NR, c->dKdiag, q.cheb_inv_theta, W.RHO, W.E, st)) return -1;
    for (int k = 0; k < q.cheb_steps; ++k) {
        if (ring_rows(c, W.E, W.NR, st)) return -1;
        if (ring_step(c->g, rg, W.NR, c->dKdiag, q.cheb_a[k], q.cheb_b[k], z, W.RHO, W.E, st)) return -1;
    }
    return 0;
}

// out = v - Q v Q^T restricted to the interior grid, Q = I - P_W (transposed: Q^T):  the separable projector Pi (Pi^T).
// u: work vector.  v and out may alias.
static int schur_project(sem_ctx* c, int transposed, const double* v, double* u, double* out, cudaStream_t st) {
    sem_ctx::NsSchur& q = c->sch;
    const MeshDev& g = c->g;
    GemmArgs a;
    std::memset(&a, 0, sizeof(a));
    a.nprob = 1; a.batch = 1; a.tile = q.tile;
    if (pw_restrict(g, c->tab(), q.px, 0, transposed, v, q.cx, q.ldx, st)) return -1;
    if (c->has_comm && comm_allreduce_sum(c->comm, q.cx, q.kxp * q.ldx, st)) return -1;
    a.p[0] = GemmProblem{q.TinvX, q.cx, q.cx2, q.kxp, q.ldx, q.ldx, 0, 0, 0, q.kxp, q.ldx, q.kxp, nullptr, nullptr};   // cx2 = T^-1 cx
    if (gemm_launch(a, EPI_NONE, st)) return -1;
    if (pw_prolong(g, c->tab(), q.px, 0, transposed, q.cx2, q.ldx, v, u, st)) return -1;        // u = (I - P_x) v
    if (pw_restrict(g, c->tab(), q.py, 1, transposed, u, q.cy, q.kyp, st)) return -1;
    a.p[0] = GemmProblem{q.cy, q.TinvY, q.cy2, q.kyp, q.kyp, q.kyp, 0, 0, 0, q.rowsy, q.kyp, q.kyp, nullptr, nullptr};  // cy2 = cy T^-1
    if (gemm_launch(a, EPI_NONE, st)) return -1;
    if (pw_prolong(g, c->tab(), q.py, 1, transposed, q.cy2, q.kyp, u, u, st)) return -1;        // u = (I - P_y) u
    return aux_sub(v, u, out, (long long)g.NX * g.LD, st);                                      // out = v - u
}

// z1 = Pi Shat^+ Pi^T y on the inner nodes, z1[pin] = y[pin], zero on the ring
static int schur_coarse(sem_ctx* c, const double* y, double* z1, cudaStream_t st) {
    const SchurWork W = schur_work(c);
    const Regions rg{c->pin_gx, c->pin_iy};
    const long long vlen = (long long)c->g.NX * c->g.LD;
    if (schur_inner(c->g, c->tab(), rg, y, nullptr, 0, W.G0, st)) return -1;
    if (schur_project(c, 1, W.G0, W.G1, W.V0, st)) return -1;
    if (fdm_apply(c, SEM_FDM_COARSE, W.V0, W.V1, 1, vlen, st)) return -1;
    if (schur_project(c, 0, W.V1, W.G1, W.G0, st)) return -1;
    return schur_inner(c->g, c->tab(), rg, W.G0, y, 0, z1, st);
}

// r1 = y - S_0 z1,  S_0 = D - C A_0^-1 G with the Stokes velocity block A_0 (what the fast diagonalisation inverts exactly)
static int schur_stokes_res(sem_ctx* c, const double* y, const double* z1, double* tmp, double* r1, cudaStream_t st) {
    const SchurWork W = schur_work(c);
    const Regions rg{c->pin_gx, c->pin_iy};
    const long long vlen = (long long)c->g.NX * c->g.LD;
    MarchArgs A = zero_args();
    A.a = z1; A.y0 = W.G0; A.y1 = W.G1; A.cconv = 1.0;
    if (apply_and_exchange(c, MODE_G, A, {W.G0, W.G1}, st)) return -1;
    if (schur_zero_boundary2(c->g, W.G0, W.G1, st)) return -1;
    if (fdm_apply(c, SEM_FDM_MAIN, W.G0, W.V0, 2, vlen, st)) return -1;
    MarchArgs B = zero_args();
    B.a = W.V0; B.b = W.V1; B.y0 = tmp;
    if (apply_and_exchange(c, MODE_DIV, B, {tmp}, st)) return -1;
    if (ring_rows(c, z1, W.NR, st)) return -1;
    return schur_stokes_residual(c->g, rg, y, W.NR, tmp, z1, r1, st);
}

// z2 = M^-1 F_p K_N^+ r1 on the inner nodes (F_p = K + Re (diag(u) G_x + diag(v) G_y) on the pressure nodes, K_N^+ = the
// pseudo-inverse of the all-Neumann Laplacian by fast diagonalisation), z2[pin] = r1[pin], zero on the ring
static int schur_pcd(sem_ctx* c, const sem_ns_state& lin, const double* r1, double* z2, cudaStream_t st) {
    const SchurWork W = schur_work(c);
    const Regions rg{c->pin_gx, c->pin_iy};
    const long long vlen = (long long)c->g.NX * c->g.LD;
    if (schur_inner(c->g, c->tab(), rg, r1, nullptr, 0, W.G0, st)) return -1;
    if (fdm_apply(c, SEM_FDM_NEUMANN, W.G0, W.G1, 1, vlen, st)) return -1;
    MarchArgs A = zero_args();
    A.a = W.G1; A.U = lin.u; A.V = lin.v; A.cconv = lin.Re; A.y0 = W.V0;
    if (apply_and_exchange(c, MODE_CD, A, {W.V0}, st)) return -1;
    return schur_inner(c->g, c->tab(), rg, W.V0, r1, 1, z2, st);
}

// z = P^-1 r for r = (r_u | r_v | r_c), three vecs `vlen` apart; tmp: one scratch vec
static int ns_precond_apply(sem_ctx* c, const sem_ns_state& lin, int level, const double* r, double* z, double* tmp,
                            cudaStream_t st) {
    const long long vlen = (long long)c->g.NX * c->g.LD;
    if (level == 0) return aux_axpby(1.0, r, 0.0, z, 3 * vlen, st);
    if (level >= 2) {
        if (fdm_apply(c, SEM_FDM_MAIN, r, z, 2, vlen, st)) return -1;
    } else if (aux_ns_jacobi(c->g, c->dKdiag, lin.gxu, lin.gyv, r, r + vlen, z, z + vlen, st)) {
        return -1;
    }
    MarchArgs A = zero_args();
    A.a = z; A.b = z + vlen; A.y0 = tmp;
    if (apply_and_exchange(c, MODE_DIV, A, {tmp}, st)) return -1;
    const double* rc = r + 2 * vlen;
    double* zp = z + 2 * vlen;
    if (level <= 2) return aux_ns_schur_mass(c->g, c->tab(), rc, tmp, zp, c->pin_gx, c->pin_iy, st);
    sem_ctx::NsSchur& q = c->sch;
    if (!q.ready) { set_error("ns_precond_apply: sem_ctx_set_ns_schur has not been called"); return -2; }
    const SchurWork W = schur_work(c);
    const Regions rg{c->pin_gx, c->pin_iy};
    if (schur_rhs(c->g, rg, rc, tmp, W.Y, st)) return -1;
    if (level == 3 || !q.two_level) {
        if (schur_inner(c->g, c->tab(), rg, W.Y, W.Y, 1, zp, st)) return -1;
        return schur_add_ring(c, zp, W.Y, st);
    }
    if (schur_coarse(c, W.Y, W.Z1, st)) return -1;
    if (schur_add_ring(c, W.Z1, W.Y, st)) return -1;
    if (schur_stokes_res(c, W.Y, W.Z1, tmp, W.R1, st)) return -1;
    if (schur_pcd(c, lin, W.R1, zp, st)) return -1;
    if (schur_add_ring(c, zp, W.R1, st)) return -1;
    if (aux_axpby(1.0, W.Z1, 1.0, zp, vlen, st)) return -1;
    if (q.singular) {
        if (ctx_multi_dot(c, q.mc, vlen, 1, zp, q.sums, 1, vlen, st)) return -1;
        if (ctx_multi_dot(c, q.lc, vlen, 1, W.Y, q.sums + 1, 1, vlen, st)) return -1;
        if (member_update(c->g, q.lc, q.sums, q.inv_den, zp, st)) return -1;
    }
    return 0;
}

// One application of the preconditioner / of one of its stages (tests compare them with oracle/ns_precond.py).
//   what 0: z3 = P^-1 r3 at `level`            (in = 3 vecs, out = 3 vecs)
//        1: out = coarse(in)   2: out = in2 with the ring added for right-hand side in (add_ring)
//        3: out = in - S_0 in2  (r1 for y = in, z1 = in2)   4: out = pcd(in)      (single vecs)
extern "C" int sem_ns_precond_debug(sem_ctx* c, const sem_ns_state* s, int what, int level, const double* in, const double* in2,
                                    double* out, void* stream) {
    SEM_CHECK_CTX(c);
    cudaStream_t st = (cudaStream_t)stream;
    const long long vlen = (long long)c->g.NX * c->g.LD;
    if (ensure_kdiag(c, st)) return -1;
    double* tmp = nullptr;
    SEM_CUDA(cudaMalloc(&tmp, sizeof(double) * vlen));
    SEM_CUDA(cudaMemsetAsync(tmp, 0, sizeof(double) * vlen, st));
    int rc = 0;
    if (what != 0 && !c->sch.ready) { set_error("sem_ns_precond_debug: sem_ctx_set_ns_schur has not been called"); rc = -2; }
    else if (what == 0) rc = ns_precond_apply(c, *s, level, in, out, tmp, st);
    else if (what == 1) rc = schur_coarse(c, in, out, st);
    else if (what == 2) { rc = aux_axpby(1.0, in2, 0.0, out, vlen, st); if (!rc) rc = schur_add_ring(c, out, in, st); }
    else if (what == 3) rc = schur_stokes_res(c, in, in2, tmp, out, st);
    else if (what == 4) rc = schur_pcd(c, *s, in, out, st);
    else { set_error("sem_ns_precond_debug: unknown stage"); rc = -2; }
    cudaStreamSynchronize(st);
    cudaFree(tmp);
    return rc;
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

// Z != nullptr: FLEXIBLE GMRES -- the preconditioned vectors Z_j = Pinv(V_j) are kept (restart vectors of n doubles) and the
// solution is updated as x += Z y instead of x += Pinv(V y).  Needed when Pinv is only approximately linear (an inner Krylov
// solve to a tolerance): otherwise the true residual stalls at the inner tolerance while the Arnoldi estimate keeps falling.
static int gmres(sem_ctx* c, const GmresLayout& L, const vecop& Aop, const vecop& Pinv, const double* b, double* x,
                 sem_krylov* kr, double* V, double* w, double* t, double* vin, cudaStream_t st, bool use_graph,
                 double* Z = nullptr) {
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
        // a NaN / Inf residual (non-finite input, overflowing operator, broken-down singular system) is a failure, not
        // convergence: the reference's lgmres reports info != 0 and the solvers raise (CD:149-150, NS:225-226)
        if (!std::isfinite(beta)) { set_error("gmres: non-finite residual"); return -3; }
        if (beta <= kr->atol) return 0;
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
            if (Z && aux_axpby(1.0, t, 0.0, Z + (long long)j * n, n, st)) return -1;
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
        int k = j;
        for (int i = 0; i < k; ++i)   // a zero pivot (exact breakdown of a singular system): use the columns before it
            if (!(std::fabs(H[(size_t)i * (m + 1) + i]) > 0.0)) { k = i; break; }
        for (int i = k - 1; i >= 0; --i) {
            double s = gv[i];
            for (int q = i + 1; q < k; ++q) s -= H[(size_t)q * (m + 1) + i] * yv[q];
            yv[i] = s / H[(size_t)i * (m + 1) + i];
        }
        std::memcpy(c->h_small, yv.data(), sizeof(double) * k);
        SEM_CUDA(cudaMemcpyAsync(c->d_small, c->h_small, sizeof(double) * k, cudaMemcpyHostToDevice, st));
        if (Z) {
            if (aux_multi_comb(Z, n, k, c->d_small, t, c->rs, st)) return -1;
        } else {
            if (aux_multi_comb(V, n, k, c->d_small, w, c->rs, st)) return -1;
            if (Pinv(w, t)) return -1;
        }
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
        if (kr->precond == 2) return fdm_apply(c, SEM_FDM_MAIN, r, z, 1, vlen, st);
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
    vecop Pinv = [&](const double* r, double* z) { return ns_precond_apply(c, lin, kr->precond, r, z, tmp, st); };
    GmresLayout L{n, 3, vlen};
    const int rc = gmres(c, L, Aop, Pinv, rhs3, x3, kr, V, w, t, vin, st, use_graph);
    SEM_CUDA(cudaStreamSynchronize(st));
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------
// Coupled Boussinesq system on the device (reference: the OpenMDAO Newton-Krylov of
// OpenMDAO/Boussinesq_SequentialCoupler.py:75-94 -- ScipyKrylov GMRES(restart) on apply_linear with one block-Jacobi sweep of
// solve_linear as preconditioner).  The coupled vector is [dT (CD mesh) | du | dv | dp (NS mesh)]; the Jacobian-vector
// product is the two fused operator kernels with the fields moved between the two meshes by the interpolation kernel
// (change_inputs, CD_Component.py:23-36 / NS_Component.py:23-33); the preconditioner is one CD solve and one NS solve (right-hand
// side projected onto the range of the singular NS Jacobian).  Everything stays in HBM: per Krylov iteration only the small
// Hessenberg columns travel to the host.  One GPU (both solvers in this process); `outer` is a context on the NS mesh whose
// reduction scratch serves the outer iteration (the inner solves use the scratch of their own contexts).
// ---------------------------------------------------------------------------------------------------------------
static int transfer(sem_ctx* from, sem_ctx* to, const sem_transfer& t, const double* src, double* dst, cudaStream_t st) {
    if (t.nxp != to->g.NX || t.nyp != to->g.NY) { set_error("sem_coupled: transfer table does not match the target mesh"); return -2; }
    return aux_interpolate(from->g, src, t.nxp, t.mx, t.Sx, t.nyp, t.ny, t.Sy, dst, to->g.LD, st);
}

extern "C" long long sem_coupled_vec_len(const sem_coupled* q) {
    if (!q || !q->ns || !q->cd) return -1;
    return (long long)q->cd->g.NX * q->cd->g.LD + 3ll * q->ns->g.NX * q->ns->g.LD;
}
extern "C" long long sem_coupled_work_len(const sem_coupled* q, int restart) {
    const long long n = sem_coupled_vec_len(q);
    if (n < 0) return -1;
    // basis + preconditioned basis (flexible GMRES) + w, t, vin; scratch: 2 CD vecs (u, v on the CD mesh), 1 NS vec (T on the NS
    // mesh), 3 NS vecs (projected right-hand side)
    return (long long)(2 * restart + 1 + 3) * n + 2ll * q->cd->g.NX * q->cd->g.LD + 4ll * q->ns->g.NX * q->ns->g.LD;
}

static int coupled_jvp(const sem_coupled* q, double* scratch, const double* x, double* y, cudaStream_t st) {
    const long long vc = (long long)q->cd->g.NX * q->cd->g.LD, vn = (long long)q->ns->g.NX * q->ns->g.LD;
    const double *dT = x, *du = x + vc, *dv = du + vn, *dp = dv + vn;
    double *uc = scratch, *vcd = scratch + vc, *Tn = scratch + 2 * vc;
    if (transfer(q->ns, q->cd, q->ns_to_cd, du, uc, st) || transfer(q->ns, q->cd, q->ns_to_cd, dv, vcd, st)) return -1;
    if (transfer(q->cd, q->ns, q->cd_to_ns, dT, Tn, st)) return -1;
    if (sem_cd_jvp(q->cd, q->cd_state, dT, uc, vcd, y, (void*)st)) return -1;
    return sem_ns_jvp(q->ns, q->ns_state, du, dv, dp, Tn, y + vc, y + vc + vn, y + vc + 2 * vn, (void*)st);
}

extern "C" int sem_coupled_jvp(const sem_coupled* q, const double* x, double* y, double* scratch, void* stream) {
    if (!q || !q->ns || !q->cd) { set_error("sem_coupled_jvp: null argument"); return -2; }
    SEM_CUDA(cudaSetDevice(q->ns->device));
    return coupled_jvp(q, scratch, x, y, (cudaStream_t)stream);
}

extern "C" int sem_coupled_solve(sem_ctx* outer, sem_coupled* q, const double* rhs, double* x, sem_krylov* kr, double* work,
                                 long long work_len, void* stream) {
    SEM_CHECK_CTX(outer);
    if (!q || !q->ns || !q->cd || !q->kr_ns || !q->kr_cd) { set_error("sem_coupled_solve: null argument"); return -2; }
    if (outer->has_comm || q->ns->has_comm || q->cd->has_comm) { set_error("sem_coupled_solve: one GPU only"); return -2; }
    cudaStream_t st = (cudaStream_t)stream;
    const long long vc = (long long)q->cd->g.NX * q->cd->g.LD, vn = (long long)q->ns->g.NX * q->ns->g.LD, n = vc + 3 * vn;
    if (kr->restart > SEM_MAX_RESTART) kr->restart = SEM_MAX_RESTART;
    if (work_len < sem_coupled_work_len(q, kr->restart)) { set_error("sem_coupled_solve: work buffer too small"); return -2; }
    double* V = work;
    double* Z = work + (long long)(kr->restart + 1) * n;   // flexible GMRES: the block solves are Krylov solves to a tolerance
    double* w = Z + (long long)kr->restart * n;
    double* t = w + n;
    double* vin = t + n;
    double* scratch = vin + n;                 // 2 CD vecs + 1 NS vec
    double* proj = scratch + 2 * vc + vn;      // 3 NS vecs
    SEM_CUDA(cudaMemsetAsync(w, 0, sizeof(double) * (3 * n + 2 * vc + 4 * vn), st));
    q->iters_cd = q->iters_ns = q->solves = 0;
    vecop Aop = [&](const double* xx, double* yy) { return coupled_jvp(q, scratch, xx, yy, st); };
    vecop Pinv = [&](const double* r, double* z) {
        // LinearBlockJac, one sweep, zero initial guess (BSC:89,94): solve_linear of each component
        SEM_CUDA(cudaMemsetAsync(z, 0, sizeof(double) * n, st));
        sem_krylov kc = *q->kr_cd;
        int rc = sem_cd_solve(q->cd, q->cd_state, r, z, &kc, q->cd_work, q->cd_work_len, (void*)st);
        q->iters_cd += kc.iters;
        if (rc) { if (rc > 0) set_error("sem_coupled_solve: the CD block solve did not converge"); return rc > 0 ? -6 : rc; }
        SEM_CUDA(cudaMemcpyAsync(proj, r + vc, sizeof(double) * 3 * vn, cudaMemcpyDeviceToDevice, st));
        if (q->ns_null) {   // b <- b - l (l.b) / (l.l): the NS block is singular, a Krylov vector need not lie in its range
            if (ctx_multi_dot(q->ns, q->ns_null, 3 * vn, 1, proj, q->ns->d_small, 3, vn, st)) return -1;
            if (aux_axpy_dev(3 * vn, q->ns_null, q->ns->d_small, -1.0 / q->ns_null_nrm2, proj, st)) return -1;
        }
        sem_krylov kn = *q->kr_ns;
        rc = sem_ns_solve(q->ns, q->ns_state, proj, z + vc, &kn, q->ns_work, q->ns_work_len, (void*)st);
        q->iters_ns += kn.iters;
        q->solves += 1;
        if (rc) { if (rc > 0) set_error("sem_coupled_solve: the NS block solve did not converge"); return rc > 0 ? -6 : rc; }
        return 0;
    };
    GmresLayout L{n, 1, n};
    const int rc = gmres(outer, L, Aop, Pinv, rhs, x, kr, V, w, t, vin, st, false, Z);
    SEM_CUDA(cudaStreamSynchronize(st));
    return rc;
}
