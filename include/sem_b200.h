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
    int precond;           /* 0 none, 1 Jacobi, 2 fast diagonalisation (needs sem_ctx_set_fdm slot 0); NS: block lower-
                              triangular with that velocity block and the reference's mass preconditioner on the Schur block
                              (NS:208-212); NS only (need sem_ctx_set_ns_schur): 3 = 2 + block elimination of
                              the boundary pressure rows, 4 = 3 + two-level Schur preconditioner with a pressure convection-
                              diffusion stage (needs fdm slots 1 and 2) */
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

/* ---- fast-diagonalisation (FDM) plans: z = (Qx (x) Qy) diag(1/(lx_k + ly_l))^+ (Qx (x) Qy)^T r --------------------------------
 * Replaces the SuperLU factorisation / solves of NS:178-191 (velocity block) as the preconditioner of the Krylov solvers
 * (sem_krylov.precond >= 2) and serves the pressure stages of the NS preconditioner.  One direction of a plan: the operator
 * acts on the nodes [lo, lo + cnt) of the direction; Q^T M1 Q = I, Q^T K1 Q = diag(lam) are the generalised eigenpairs of the
 * assembled 1-D pencil restricted to those nodes.  fold = 1 (pencil symmetric about the domain centre): Qe [ne][ne] holds the
 * even eigenvectors on the first ne = ceil(cnt/2) nodes, Qo [no][no] the odd ones on the first no = floor(cnt/2) nodes, lam
 * the ne even then the no odd eigenvalues -- every transform then costs half the flops.  fold = 0: Qe [cnt][cnt], Qo unused.
 * All arrays are DEVICE pointers, row-major [node][mode], and are copied.  On a partitioned context the x direction is the
 * GLOBAL pencil, unfolded: Qe = the slab's rows [NX_local][nmodes] of the global matrix (zero rows at eliminated end nodes),
 * lam[nmodes], lo / cnt = the lines of the slab the operator acts on; the x transform is then distributed over the ranks
 * (product + ncclReduceScatter, ncclAllGather + product) and stays the exact inverse.
 * slot: 0 = Laplacian with the solver's Dirichlet sides (CD operator, NS velocity block), 1 = all-Neumann pressure
 * Laplacian, 2 = coarse operator of the NS Schur complement.  outside: 1 = z = r outside the node range (identity rows),
 * 0 = z = 0.  den_floor: denominators <= den_floor are treated as zero modes (pseudo-inverse); < 0 = automatic. */
typedef struct {
    int lo, cnt, fold;
    const double *Qe, *Qo, *lam;
    int nmodes;            /* partitioned x direction only: number of global modes */
} sem_fdm_dir;
int sem_ctx_set_fdm(sem_ctx *ctx, int slot, const sem_fdm_dir *x, const sem_fdm_dir *y, int outside, double den_floor);
/* z = plan(r) for nf (1 or 2) consecutive vecs; r and z may alias */
int sem_fdm_apply(sem_ctx *ctx, int slot, const double *r, double *z, int nf, void *stream);

/* ---- Schur-complement stages of the NS preconditioner (sem_krylov.precond = 3, 4); replaces the diagonal-mass
 * preconditioner of NS:208-212.  All arrays are HOST pointers and are copied.
 *   wl, wr [P+1]        element-local values (1 - xn_j) L_P(xi_j), xn_j L_P(xi_j) of the coarse-space functions
 *   Tinv_x [(N_ex+1)^2], Tinv_y [(N_ey+1)^2]   inverses (dense, row-major) of the tridiagonal Gram matrices W^T M W
 *   lfx [NX_global], lfy [NY]             1-D factors of the pressure part l_c = lfx (x) lfy of the Jacobian's left null vector
 *   singular            the Jacobian is singular (l_c vanishes on the boundary and at the pin): apply the member correction
 *   inv_den             1 / (m_c . l_c), m_c = M_p l_c
 *   two_level           0: only the ring elimination is available (level 4 falls back to 3)
 *   cheb_lo, cheb_hi, cheb_steps           spectrum bounds of the diagonally scaled ring block and polynomial degree */
typedef struct {
    const double *wl, *wr;
    const double *Tinv_x, *Tinv_y;
    const double *lfx, *lfy;
    int singular, two_level;
    double inv_den;
    double cheb_lo, cheb_hi;
    int cheb_steps;
} sem_ns_schur_desc;
int sem_ctx_set_ns_schur(sem_ctx *ctx, const sem_ns_schur_desc *desc);

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
/* Operator applies without a post-operator (K, G, DIV, CD) on the peer-memory path are ONE kernel launch: the CTAs that
 * finish an interface line store their segment of it straight into the neighbour's mailbox (per-strip epoch flags) and add
 * the neighbour's contribution at the end of their chunk (SEM_B200_FUSED_XCH=0: the three-launch path + exchange kernel).
 * sem_ctx_attach_loopback: one-GPU self-test of these paths -- the context must be an inner slab (0 < m_begin,
 * m_end < N_ex); it becomes its own left and right neighbour, so line 0 and the last line are exchanged with each other. */
int sem_ctx_attach_loopback(sem_ctx *ctx);
/* partitioned operator applies issued so far: fused != 0 -> as one launch with the in-kernel exchange, else as edge /
 * interior launches + exchange kernel */
long long sem_ctx_partitioned_applies(const sem_ctx *ctx, int fused);

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

/* ---- interpolation: SEM.eval_interpolation (SEM.py:248-273) on an ij-meshgrid of nxp x nyp points; also the mesh-to-mesh
 * transfer `change_inputs` of the OpenMDAO components (CD_Component.py:23-36, NS_Component.py:23-33).  mx[nxp] / ny[nyp]:
 * GLOBAL element column / row of every plot column / row (SEM.x2xi, SEM.py:23-36); Sx [nxp][P+1] / Sy [nyp][P+1]: the Lagrange
 * basis values there (GLL.standard_evaluation_matrix, GLL.py:105-116); out [nxp][nyp].  All DEVICE pointers.  On a partitioned
 * context every rank receives the whole array (the slabs' contributions are summed with ncclAllReduce). */
int sem_interpolate(sem_ctx *ctx, const double *vec, int nxp, const int *mx, const double *Sx, int nyp, const int *ny,
                    const double *Sy, double *out, void *stream);

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

/* One application of the NS preconditioner (what = 0: z3 = P^-1 r3 at `level`) or of one of its stages (1 coarse, 2 ring
 * elimination, 3 Stokes-Schur residual, 4 pressure convection-diffusion) -- for the stage-by-stage parity tests. */
int sem_ns_precond_debug(sem_ctx *ctx, const sem_ns_state *st, int what, int level, const double *in, const double *in2,
                         double *out, void *stream);

/* ---- coupled Boussinesq system on the device: the linear solve of the OpenMDAO Newton-Krylov coupling
 * (OpenMDAO/Boussinesq_SequentialCoupler.py:75-94: ScipyKrylov GMRES(restart) on apply_linear, one block-Jacobi sweep of
 * solve_linear as preconditioner).  Coupled vectors are [dT (CD mesh vec) | du | dv | dp (NS mesh vecs)], device resident.
 * sem_transfer: tables of the mesh-to-mesh interpolation (change_inputs, CD_Component.py:23-36, NS_Component.py:23-33) onto
 * EVERY node of the target mesh -- element column / row of each target line (mx[nxp], ny[nyp]) and Lagrange basis values
 * (Sx [nxp][P+1], Sy [nyp][P+1]) of the SOURCE mesh, DEVICE arrays as for sem_interpolate.  One GPU. */
typedef struct {
    int nxp, nyp;
    const int *mx, *ny;
    const double *Sx, *Sy;
} sem_transfer;
typedef struct {
    sem_ctx *ns, *cd;
    const sem_ns_state *ns_state;      /* linearisation point incl. Jacobian diagonals (after _get_residuals + _calc_jacobians) */
    const sem_cd_state *cd_state;      /* incl. gxT, gyT */
    sem_transfer ns_to_cd, cd_to_ns;
    const sem_krylov *kr_ns, *kr_cd;   /* controls of the inner (block) solves */
    double *ns_work; long long ns_work_len;     /* work buffers of the inner solves (sem_ns_work_len / sem_cd_work_len) */
    double *cd_work; long long cd_work_len;
    const double *ns_null;             /* 3 NS vecs: left null vector of the NS Jacobian, or NULL (regular Jacobian) */
    double ns_null_nrm2;
    /* outputs of sem_coupled_solve */
    int iters_cd, iters_ns, solves;
} sem_coupled;
long long sem_coupled_vec_len(const sem_coupled *q);
long long sem_coupled_work_len(const sem_coupled *q, int restart);
/* y = J x; scratch: 2 CD vecs + 1 NS vec */
int sem_coupled_jvp(const sem_coupled *q, const double *x, double *y, double *scratch, void *stream);
/* solve J x = rhs (x holds the guess); `outer`: a context on the NS mesh whose reduction scratch serves the outer iteration */
int sem_coupled_solve(sem_ctx *outer, sem_coupled *q, const double *rhs, double *x, sem_krylov *kr, double *work,
                      long long work_len, void *stream);

/* ---- reductions used by the Python layer (deterministic two-stage sums) ---------------------------------------- */
/* n = (number of fields) * sem_ctx_vec_len(); interface lines are counted once and the sum is global over ranks */
int sem_dot(sem_ctx *ctx, const double *x, const double *y, long long n, double *host_out, void *stream);
/* y = a*x + b*y over n doubles (b == 0: y = a*x); the Newton update u += du of NS:265-267 / T + dT of CD:170 */
int sem_axpby(sem_ctx *ctx, double a, const double *x, double b, double *y, long long n, void *stream);

#ifdef __cplusplus
}
#endif
#endif
