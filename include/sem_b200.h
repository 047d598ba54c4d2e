/*
 * sem_b200 -- C ABI of the B200-native (sm_100a) hot path of the 2-D spectral element solver Tangxiaotian11/SEM.
 *
 * The reference has no FFI: its boundary is the Python class API of Solvers/ConvectionDiffusion_Solver.py and
 * Solvers/NavierStokes_Solver.py.  The Python classes in sem_b200/ keep that API and call the entry points below
 * through ctypes; each entry point names the reference code (file:line under /root/reference) it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, a negative value on a CUDA/NCCL/argument error (text via
 *     sem_last_error()) and, for the solvers, a positive value = "not converged" (iterations spent).
 *   - "vec" arguments are DEVICE pointers to fp64 field vectors in the padded layout of this context:
 *     node (ix, iy) of the local slab at  vec[ix * LD + iy],  0 <= ix < NX_local, 0 <= iy < NY, LD >= NY
 *     (x slow / y fast like SEM.py:110, rows padded to LD = sem_ctx_ld()).  Pad entries are kept at zero.
 *   - "host" arguments are HOST pointers to the reference's dense layout (length N = NX*NY, SEM.py:110).
 *   - no memory ownership crosses the ABI: device storage is allocated by the caller (torch tensors).
 *   - stream is a cudaStream_t passed as void* (NULL = legacy default stream).
 */
#ifndef SEM_B200_H
#define SEM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sem_ctx sem_ctx;

/* Mesh + partition description (replaces the mesh set-up of CD:38-50 / NS:53-63 and the tables of GLL.py). */
typedef struct {
    int P;                 /* polynomial order (1..16) */
    int N_ex, N_ey;        /* GLOBAL element counts */
    double dx, dy;         /* element widths L_x/N_ex, L_y/N_ey */
    const double *D;       /* host, (P+1)^2 row-major: GLL.standard_differentiation_matrix (GLL.py:45-59) */
    const double *Ks;      /* host, (P+1)^2 row-major: GLL.standard_stiffness_matrix (GLL.py:73-81) */
    const double *w;       /* host, P+1: GLL weights (GLL.py:30) */
    int device;            /* CUDA device ordinal */
    int m_begin, m_end;    /* element columns [m_begin, m_end) owned by this rank (0, N_ex on one GPU) */
} sem_mesh_desc;

/* Dirichlet description of the CD solver (CD:62-71): sides W, E, S, N; later sides win at corners. */
typedef struct {
    int active[4];
    double value[4];
} sem_cd_bc;

/* Boundary data of the NS solver (NS:78-91): tangential wall velocities; the pressure pin is node int(N/2). */
typedef struct {
    double v_W, v_E, u_S, u_N;
} sem_ns_bc;

/* Linearisation state of the CD solver == what CD:73-102 caches in self._Sys / self._Jac_T_u / self._Jac_T_v. */
typedef struct {
    sem_cd_bc bc;
    double Pe;
    const double *u, *v;        /* vec: advecting velocity of the last _get_residuals */
    const double *gxT, *gyT;    /* vec: Pe*G_x T, Pe*G_y T of the last _calc_jacobians (may be NULL) */
} sem_cd_state;

/* Linearisation state of the NS solver == self._Sys / self._Jac_* of NS:93-136. */
typedef struct {
    sem_ns_bc bc;
    double Re, Gr_over_Re;
    const double *u, *v;                      /* vec: advecting velocity of the last _get_residuals */
    const double *gxu, *gyu, *gxv, *gyv;      /* vec: Re*G_x u, Re*G_y u, Re*G_x v, Re*G_y v (NS:131-136) */
} sem_ns_state;

/* Krylov controls and statistics (replaces the lgmres arguments of CD:146-148 / NS:222-224). */
typedef struct {
    double atol;           /* stop when the true residual 2-norm <= atol */
    int restart;           /* Krylov basis size per cycle */
    int max_iters;         /* total iteration cap */
    int precond;           /* 0 none, 1 Jacobi, 2 fast diagonalisation (needs sem_ctx_set_fdm); NS: block lower-
                              triangular with that velocity block; NS only, EXPERIMENTAL: 3 = 2 + block
                              elimination of the boundary pressure rows (needs sem_ctx_set_pbb) */
    int verbose;
    /* outputs */
    int iters;             /* operator applications spent */
    double resnorm;        /* final residual norm */
} sem_krylov;

const char *sem_last_error(void);
int sem_version(void);

/* ---- context ------------------------------------------------------------------------------------------------ */
int sem_ctx_create(sem_ctx **out, const sem_mesh_desc *desc);
void sem_ctx_destroy(sem_ctx *ctx);
int sem_ctx_ld(const sem_ctx *ctx);            /* padded row pitch LD (doubles) */
int sem_ctx_nx(const sem_ctx *ctx);            /* local node lines (m_end-m_begin)*P+1 */
int sem_ctx_ny(const sem_ctx *ctx);            /* N_ey*P+1 */
long long sem_ctx_vec_len(const sem_ctx *ctx); /* doubles per vec = NX_local*LD */
/* launch tuning: elements per strip in y and per marching chunk in x (0 = automatic) */
int sem_ctx_set_tiling(sem_ctx *ctx, int Ty, int Mx);

/* ---- fast-diagonalisation (FDM) preconditioner of the Krylov solvers (sem_krylov.precond = 2) ----------------------
 * Qx: DEVICE, row-major [NX][NX], column k = k-th generalised eigenvector of the assembled 1-D pencil (K1x, M1x) with the
 * Dirichlet end nodes eliminated (their rows, and the unused trailing columns, are zero), normalised Qx^T M1x Qx = I;
 * lamx[NX]: the eigenvalues (1 for unused columns); Qy, lamy likewise for y.  dirichlet_wesn[4]: which sides carry
 * Dirichlet rows.  The arrays are copied.  On a partitioned context (after sem_ctx_attach_comm) Qx is the slab's row range
 * of the GLOBAL matrix, row-major [NX_local][NX_global], lamx[NX_global]; the x transform is then distributed over the ranks
 * (GEMM + ncclReduceScatter, ncclAllGather + GEMM) and the preconditioner stays the exact inverse. */
int sem_ctx_set_fdm(sem_ctx *ctx, const double *Qx, const double *lamx, const double *Qy, const double *lamy,
                    const int *dirichlet_wesn);

/* EXPERIMENTAL (written at the end of round 1, compiled but not yet validated on a GPU; nothing uses it unless
 * sem_krylov.precond == 3): boundary block of the NS pressure rows.  idx_host[nb]: offsets ix*LD + iy of the boundary pressure
 * nodes (the pin node left out); inv_dev: DEVICE, row-major nb x nb inverse of the stiffness matrix restricted to those nodes
 * (rows K[mask,:] of NS:119,157).  The NS preconditioner then solves the boundary rows exactly, z_B = K_BB^-1 (r_B - K_BI z_I),
 * instead of scaling them by 1/M.  One GPU only.  The arrays are copied. */
int sem_ctx_set_pbb(sem_ctx *ctx, const long long *idx_host, int nb, const double *inv_dev);

/* ---- multi-GPU: one process per GPU, element columns [m_begin, m_end) per rank (sem_mesh_desc).  Rank 0 creates a
 * 128-byte NCCL unique id, the caller distributes it (torch.distributed), every rank attaches.  Afterwards every operator
 * apply ends with the exchange of the interface node line(s) and every dot product is all-reduced (NCCL).  The interface
 * exchange itself goes through peer memory: each rank owns a mailbox in its HBM that its two neighbours map with
 * cudaIpcOpenMemHandle at attach time; a push kernel stores the lines into the neighbour's mailbox over NVLink and
 * releases an epoch flag, the consumer acquires it and adds.  sem_ctx_comm_mode: 0 = no communicator, 1 = NCCL send/recv
 * fallback (peer mapping unavailable or SEM_B200_NO_P2P set), 2 = peer-memory mailboxes. */
int sem_nccl_unique_id(unsigned char *out128);
int sem_ctx_attach_comm(sem_ctx *ctx, const unsigned char *id128, int rank, int world);
int sem_ctx_comm_mode(const sem_ctx *ctx);

/* ---- host <-> device packing of the reference's dense vectors (the numpy <-> device boundary) ----------------- */
int sem_h2d(sem_ctx *ctx, const double *host_local, double *vec, void *stream);
int sem_d2h(sem_ctx *ctx, const double *vec, double *host_local, void *stream);

/* ---- single operators: matrix-free counterparts of SEM.global_*_matrix (SEM.py:170-223) applied to a vector ---- */
int sem_apply_stiffness(sem_ctx *ctx, const double *x, double *y, void *stream);              /* y = K x        */
int sem_apply_gradient(sem_ctx *ctx, const double *x, double scale, double *gx, double *gy,
                       void *stream);                                       /* gx = s G_x x, gy = s G_y x        */
int sem_apply_mass(sem_ctx *ctx, const double *x, double *y, void *stream);                   /* y = M x        */
int sem_mass_diag(sem_ctx *ctx, double *m, void *stream);                                     /* m = diag(M)    */

/* ---- gather-scatter: SEM.assemble for 4-index arrays (SEM.py:126-131) and SEM.scatter (SEM.py:149-167) --------
 * elem is a DEVICE array [m][n][i][j] (local element columns), contiguous, j fastest.
 * The assembly is colour-ordered (4 colours = element parity in x and y), atomic-free and bitwise reproducible. */
int sem_gather_scatter(sem_ctx *ctx, const double *elem, double *y, void *stream);
int sem_scatter(sem_ctx *ctx, const double *x, double *elem, void *stream);

/* ---- convection-diffusion: CD._get_residuals / _calc_jacobians / _get_dresiduals (CD:73-121) ------------------- */
int sem_cd_residual(sem_ctx *ctx, const sem_cd_state *st, const double *T, double *res, void *stream);
int sem_cd_jacobians(sem_ctx *ctx, double Pe, const double *T, double *gxT, double *gyT, void *stream);
int sem_cd_jvp(sem_ctx *ctx, const sem_cd_state *st, const double *dT, const double *du, const double *dv,
               double *dres, void *stream);
/* The same product with HOST vectors (the reference's call: numpy in, numpy out, CD:104-121 with du = dv = None):
 * host_dT and host_dres are dense local vectors of NX_local*NY doubles (page-locked memory gives the full link rate).
 * The element columns are processed in segments: the upload of segment s+1, the kernel of segment s and the download of
 * segment s-1 run concurrently (PCIe is full duplex), so the call costs about one direction of the transfer instead of
 * upload + kernel + download.  dT_vec / dres_vec: device vecs that receive the padded input / output (scratch of the
 * caller).  Returns after the result is complete in host_dres.  On a partitioned context the interface line(s) are
 * exchanged after the last segment and downloaded once more (the download stream is in order). */
int sem_cd_jvp_host(sem_ctx *ctx, const sem_cd_state *st, const double *host_dT, double *host_dres,
                    double *dT_vec, double *dres_vec, void *stream);
/* CD._get_update (CD:123-156): solve J dT = rhs; dT holds the initial guess on entry. work >= sem_cd_work_len(). */
long long sem_cd_work_len(const sem_ctx *ctx, int restart);
int sem_cd_solve(sem_ctx *ctx, const sem_cd_state *st, const double *rhs, double *dT, sem_krylov *kr,
                 double *work, long long work_len, void *stream);

/* ---- Navier-Stokes: NS._get_residuals / _calc_jacobians / _get_dresiduals (NS:93-160) ------------------------- */
int sem_ns_residual(sem_ctx *ctx, const sem_ns_state *st, const double *u, const double *v, const double *p,
                    const double *T, double *res_u, double *res_v, double *res_c, void *stream);
int sem_ns_jacobians(sem_ctx *ctx, double Re, const double *u, const double *v, double *gxu, double *gyu,
                     double *gxv, double *gyv, void *stream);
int sem_ns_jvp(sem_ctx *ctx, const sem_ns_state *st, const double *du, const double *dv, const double *dp,
               const double *dT, double *dres_u, double *dres_v, double *dres_c, void *stream);
/* NS._get_update (NS:162-236): solve the 3-field linearised system; x = [du|dv|dp] (3 vecs, guess on entry). */
long long sem_ns_work_len(const sem_ctx *ctx, int restart);
int sem_ns_solve(sem_ctx *ctx, const sem_ns_state *st, const double *rhs3, double *x3, sem_krylov *kr,
                 double *work, long long work_len, void *stream);

/* ---- reductions used by the Python layer (deterministic two-stage sums) ---------------------------------------- */
/* n = (number of fields) * sem_ctx_vec_len(); interface lines are counted once and the sum is global over ranks */
int sem_dot(sem_ctx *ctx, const double *x, const double *y, long long n, double *host_out, void *stream);
/* y = a*x + b*y over n doubles (b == 0: y = a*x); the Newton update u += du of NS:265-267 / T + dT of CD:170 */
int sem_axpby(sem_ctx *ctx, double a, const double *x, double b, double *y, long long n, void *stream);

#ifdef __cplusplus
}
#endif
#endif
