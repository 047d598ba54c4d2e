// Kernels of the Navier-Stokes Schur-complement preconditioner (sem_schur.cu): region masks of the pressure rows, the
// boundary-ring block elimination (Chebyshev), the separable coarse-space projectors and the member-selection correction.
// Reference context: the reference solves the pressure Schur complement by LGMRES with a diagonal-mass preconditioner
// (NS:199-224); these stages replace that preconditioner (CPU mirror: oracle/ns_precond.py).
#pragma once
#include "sem_aux.cuh"

namespace semb {

// Pressure-row regions (global indices): boundary ring (the pressure-Neumann rows K[mask,:] p of NS:119,157), the pin node
// int(N/2) (NS:89; identity row, wins over the ring in JVP form NS:157-158) and the inner nodes (continuity rows).
struct Regions {
    int pin_gx, pin_iy;
};

// y = rc - (inner ? div : 0)                                    right-hand side of the Schur block
int schur_rhs(const MeshDev& g, Regions rg, const double* rc, const double* div, double* y, cudaStream_t st);
// out = inner ? src * (scale_mass ? 1 / M : 1) : (pin && pin_src ? pin_src[pin] : 0)
int schur_inner(const MeshDev& g, TabDev t, Regions rg, const double* src, const double* pin_src, int scale_mass, double* out,
                cudaStream_t st);
// gx, gy = 0 on the boundary nodes (the velocity rows there are identity rows: no pressure gradient)
int schur_zero_boundary2(const MeshDev& g, double* gx, double* gy, cudaStream_t st);
// r1 = y - (ring ? nr : inner ? -div : z1[pin])                 r1 = y - S_0 z1
int schur_stokes_residual(const MeshDev& g, Regions rg, const double* y, const double* nr, const double* div, const double* z1,
                          double* r1, cudaStream_t st);

// ---- ring block: K_BB^-1 by a fixed Chebyshev polynomial of the diagonally scaled block ---------------------------------
// ring vectors are full-layout vectors that are zero away from the ring.
// rho = (y - q) on the ring, e = rho / (theta * Kdiag)
int ring_init(const MeshDev& g, Regions rg, const double* y, const double* q, const double* kdiag, double inv_theta, double* rho,
              double* e, cudaStream_t st);
// z += e; rho -= q; e = a * e + b * rho / Kdiag
int ring_step(const MeshDev& g, Regions rg, const double* q, const double* kdiag, double a, double b, double* z, double* rho,
              double* e, cudaStream_t st);

// ---- coarse-space projector P_W = W T^-1 W^T M along one direction (dir 0: x, 1: y) on the INTERIOR nodes -----------------
struct PwDir {
    int n;                 // GLOBAL nodes of the direction, ne = (n - 1) / P elements
    const double* wl;      // [P+1] device: (1 - xn_j) L_P(xi_j)  (value of the left-vertex function at local node j)
    const double* wr;      // [P+1] device: xn_j L_P(xi_j)
};
// c[k][iy] (dir 0, k = GLOBAL vertex index) or c[ix][k] (dir 1), pitch ldc: restriction W^T (M v) (transposed: W^T v); the
// coefficients of the projection are T^-1 c with the (dense, small) inverse of the tridiagonal Gram matrix T = W^T M W,
// applied by the DMMA GEMM (sem_capi.cu)
int pw_restrict(const MeshDev& g, TabDev t, const PwDir& d, int dir, int transposed, const double* v, double* c, int ldc,
                cudaStream_t st);
// out = v - W c (transposed: v - M W c) on the interior nodes, out = v elsewhere (out may alias v)
int pw_prolong(const MeshDev& g, TabDev t, const PwDir& d, int dir, int transposed, const double* c, int ldc, const double* v,
               double* out, cudaStream_t st);

// ---- member selection: z -= l_c (m_c . z - l_c . y) / (m_c . l_c), l_c = lfx (x) lfy, m_c = M_p l_c ---------------------------
// materialise l_c and m_c (lfx indexed by the GLOBAL line, lfy by iy)
int member_vectors(const MeshDev& g, TabDev t, Regions rg, const double* lfx, const double* lfy, double* lc, double* mc,
                   cudaStream_t st);
// z -= lc * (sums[0] - sums[1]) * inv_den        (sums[0] = m_c . z, sums[1] = l_c . y, on the device)
int member_update(const MeshDev& g, const double* lc, const double* sums, double inv_den, double* z, cudaStream_t st);

}  // namespace semb
