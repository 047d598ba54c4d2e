// Host-callable wrappers of the auxiliary kernels (sem_aux.cu): pointwise operators, boundary rows, the standalone
// colour-ordered gather-scatter, and the deterministic reductions / vector updates used by the Krylov solver.
#pragma once
#include "sem_common.cuh"

namespace semb {

// Plain (runtime-P) device copies of the GLL tables for the O(sqrt(N)) and set-up kernels.
struct TabDev {
    const double* D;    // (P+1)^2
    const double* Ks;   // (P+1)^2
    const double* w;    // P+1
};

// y = M x (diagonal mass matrix, SEM.py:170-183) or m = diag(M)
int aux_mass_apply(const MeshDev& g, TabDev t, const double* x, double* y, cudaStream_t st);
// d = diag(K)  (assembled, no boundary modification)
int aux_stiffness_diag(const MeshDev& g, TabDev t, double* d, cudaStream_t st);

// Pressure-Neumann rows (NS:119,157): y[boundary nodes] = (K c)[boundary nodes]; skip_pin=1 keeps y at the pin node.
int aux_neumann_rows(const MeshDev& g, TabDev t, const double* c, double* y, int pin_gx, int pin_iy, int skip_pin,
                     cudaStream_t st);

// dense [NX][NY] <-> padded [NX][LD] repacking on the device (host copies stay 1-D and run at full PCIe rate)
int aux_pad(const MeshDev& g, const double* dense, double* vec, cudaStream_t st);
int aux_unpad(const MeshDev& g, const double* vec, double* dense, cudaStream_t st);
// the same for node lines line0 .. line0 + nlines - 1 only (`dense` and `vec` still point to line 0)
int aux_pad_lines(const MeshDev& g, const double* dense, double* vec, int line0, int nlines, cudaStream_t st);
int aux_unpad_lines(const MeshDev& g, const double* vec, double* dense, int line0, int nlines, cudaStream_t st);

// SEM.assemble (4-index) / SEM.scatter, element array [m][n][i][j]
int aux_gather_scatter(const MeshDev& g, const double* elem, double* y, cudaStream_t st);
int aux_scatter(const MeshDev& g, const double* x, double* elem, cudaStream_t st);

// ---- Krylov building blocks.  A "multi-vector" is nf fields of vlen doubles stored back to back (n = nf*vlen). ----
struct RedScratch {
    double* partials;   // device [max_blocks * max_k]
    unsigned* counter;  // device, zero-initialised ticket
    int max_blocks;
    int max_k;
    // multi_axpy: partial sums of the basis slices [slice][n] (used when n is small), one ticket per element block
    double* axpy_partials;
    unsigned* axpy_counter;
    long long axpy_len;
    int axpy_blocks;
    int sm_count;
};

// h[j] = <V_j, w>, j = 0..k-1, V_j = V + j*n.  Two-stage deterministic sum (fixed block order, no float atomics).
// skip: number of leading doubles of every field excluded from the sums (an interface line owned by the left rank).
int aux_multi_dot(const double* V, long long n, int k, const double* w, double* h, int nf, long long vlen,
                  long long skip, RedScratch rs, cudaStream_t st);
// w += sign * sum_j h[j] V_j   (h on the device)
int aux_multi_axpy(const double* V, long long n, int k, const double* h, double sign, double* w, RedScratch rs,
                   cudaStream_t st);
// out = sum_j h[j] V_j
int aux_multi_comb(const double* V, long long n, int k, const double* h, double* out, RedScratch rs, cudaStream_t st);
// v = w / sqrt(*nrm2); v2 (optional) receives a second copy
int aux_scale_inv_norm(const double* w, const double* nrm2, double* v, double* v2, long long n, cudaStream_t st);
// y = a*x + b*y   (b == 0: y = a*x without reading y)
int aux_axpby(double a, const double* x, double b, double* y, long long n, cudaStream_t st);

// ---- preconditioner pieces ----------------------------------------------------------------------------------------
// CD Jacobi: z = r / diag(K) away from Dirichlet rows, z = r on them.
int aux_cd_jacobi(const MeshDev& g, const BCSpec& bc, const double* dK, const double* r, double* z, cudaStream_t st);
// NS velocity Jacobi: z_u = r_u / (dK + gxu), z_v = r_v / (dK + gyv) inside, identity on the boundary.
int aux_ns_jacobi(const MeshDev& g, const double* dK, const double* gxu, const double* gyv, const double* ru,
                  const double* rv, double* zu, double* zv, cudaStream_t st);
// NS pressure block of the block lower-triangular preconditioner:
//   z_p = (r_c - inner * div) / M_p,  inner = 0 on boundary and pin rows, M_p = diag(M) with 1 at the pin (NS:208-212)
int aux_ns_schur_mass(const MeshDev& g, TabDev t, const double* rc, const double* div, double* zp, int pin_gx,
                      int pin_iy, cudaStream_t st);

// SEM.eval_interpolation on an ij-meshgrid (SEM.py:248-273): see k_interpolate
int aux_interpolate(const MeshDev& g, const double* f, int nxp, const int* mx, const double* Sx, int nyp, const int* ny,
                    const double* Sy, double* out, int ldo, cudaStream_t st);
// y += scale * (*coef) * x, coefficient on the device
int aux_axpy_dev(long long n, const double* x, const double* coef, double scale, double* y, cudaStream_t st);
// out = a - b
int aux_sub(const double* a, const double* b, double* out, long long n, cudaStream_t st);

}  // namespace semb
