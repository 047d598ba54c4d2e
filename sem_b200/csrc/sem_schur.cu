// Kernels of the Navier-Stokes Schur-complement preconditioner (see sem_schur.cuh; CPU mirror: oracle/ns_precond.py).
#include "sem_schur.cuh"

namespace semb {

__device__ __forceinline__ double asm_w1(const double* __restrict__ w, int P, int q, int nel) {
    const int j = q % P;
    if (j != 0) return w[j];
    double s = 0.0;
    if (q > 0) s += w[P];
    if (q < nel * P) s += w[0];
    return s;
}

// 0: inner (continuity row), 1: boundary ring, 2: pin (the pin wins on the boundary: JVP form, NS:157-158), -1: pad
__device__ __forceinline__ int region(const MeshDev& g, const Regions& rg, int ix, int iy) {
    if (iy >= g.NY) return -1;
    const int gix = g.gx0 + ix;
    if (gix == rg.pin_gx && iy == rg.pin_iy) return 2;
    if (gix == 0 || gix == g.NXg - 1 || iy == 0 || iy == g.NY - 1) return 1;
    return 0;
}

__global__ void k_schur_rhs(const MeshDev g, const Regions rg, const double* __restrict__ rc, const double* __restrict__ div,
                            double* __restrict__ y) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)g.NX * g.LD) return;
    const int reg = region(g, rg, (int)(idx / g.LD), (int)(idx % g.LD));
    double v = 0.0;
    if (reg >= 0) v = rc[idx] - (reg == 0 ? div[idx] : 0.0);
    y[idx] = v;
}
int schur_rhs(const MeshDev& g, Regions rg, const double* rc, const double* div, double* y, cudaStream_t st) {
    const long long tot = (long long)g.NX * g.LD;
    k_schur_rhs<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g, rg, rc, div, y);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

__global__ void k_schur_inner(const MeshDev g, const TabDev t, const Regions rg, const double* __restrict__ src,
                              const double* __restrict__ pin_src, int scale_mass, double* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)g.NX * g.LD) return;
    const int ix = (int)(idx / g.LD), iy = (int)(idx % g.LD);
    const int reg = region(g, rg, ix, iy);
    double v = 0.0;
    if (reg == 0) {
        v = src[idx];
        if (scale_mass) {
            const double m = (0.5 * g.dx * asm_w1(t.w, g.P, g.gx0 + ix, (g.NXg - 1) / g.P)) * (0.5 * g.dy * asm_w1(t.w, g.P, iy, g.ney));
            v /= m;
        }
    } else if (reg == 2 && pin_src) {
        v = pin_src[idx];
    }
    out[idx] = v;
}
int schur_inner(const MeshDev& g, TabDev t, Regions rg, const double* src, const double* pin_src, int scale_mass, double* out,
                cudaStream_t st) {
    const long long tot = (long long)g.NX * g.LD;
    k_schur_inner<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g, t, rg, src, pin_src, scale_mass, out);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

// boundary nodes of the slab: threads 0..NY-1 -> line 0, NY..2NY-1 -> last line, then first / last column of every line.
// Returns false for threads that have no node (global W / E lines live on the first / last rank only).
__device__ __forceinline__ bool ring_node(const MeshDev& g, int t, int& ix, int& iy) {
    if (t < g.NY) {
        if (g.has_left) return false;
        ix = 0; iy = t;
    } else if (t < 2 * g.NY) {
        if (g.has_right) return false;
        ix = g.NX - 1; iy = t - g.NY;
    } else if (t < 2 * g.NY + 2 * g.NX) {
        const int r = t - 2 * g.NY;
        ix = r % g.NX;
        iy = (r < g.NX) ? 0 : g.NY - 1;
        if ((ix == 0 && !g.has_left) || (ix == g.NX - 1 && !g.has_right)) return false;   // corners belong to W / E
    } else {
        return false;
    }
    return true;
}

__global__ void k_zero_boundary2(const MeshDev g, double* __restrict__ a, double* __restrict__ b) {
    int ix, iy;
    if (!ring_node(g, blockIdx.x * blockDim.x + threadIdx.x, ix, iy)) return;
    const long long o = (long long)ix * g.LD + iy;
    a[o] = 0.0;
    b[o] = 0.0;
}
int schur_zero_boundary2(const MeshDev& g, double* gx, double* gy, cudaStream_t st) {
    const int tot = 2 * g.NY + 2 * g.NX;
    k_zero_boundary2<<<(tot + 127) / 128, 128, 0, st>>>(g, gx, gy);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

__global__ void k_stokes_residual(const MeshDev g, const Regions rg, const double* __restrict__ y, const double* __restrict__ nr,
                                  const double* __restrict__ div, const double* __restrict__ z1, double* __restrict__ r1) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)g.NX * g.LD) return;
    const int reg = region(g, rg, (int)(idx / g.LD), (int)(idx % g.LD));
    double v = 0.0;
    if (reg == 0) v = y[idx] + div[idx];
    else if (reg == 1) v = y[idx] - nr[idx];
    else if (reg == 2) v = y[idx] - z1[idx];
    r1[idx] = v;
}
int schur_stokes_residual(const MeshDev& g, Regions rg, const double* y, const double* nr, const double* div, const double* z1,
                          double* r1, cudaStream_t st) {
    const long long tot = (long long)g.NX * g.LD;
    k_stokes_residual<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g, rg, y, nr, div, z1, r1);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

// ---- ring Chebyshev -------------------------------------------------------------------------------------------------------
__global__ void k_ring_init(const MeshDev g, const Regions rg, const double* __restrict__ y, const double* __restrict__ q,
                            const double* __restrict__ kdiag, double inv_theta, double* __restrict__ rho, double* __restrict__ e) {
    int ix, iy;
    if (!ring_node(g, blockIdx.x * blockDim.x + threadIdx.x, ix, iy)) return;
    if (g.gx0 + ix == rg.pin_gx && iy == rg.pin_iy) return;   // a pin on the boundary is not part of the ring
    const long long o = (long long)ix * g.LD + iy;
    const double r = y[o] - q[o];
    rho[o] = r;
    e[o] = r * inv_theta / kdiag[o];
}
int ring_init(const MeshDev& g, Regions rg, const double* y, const double* q, const double* kdiag, double inv_theta, double* rho,
              double* e, cudaStream_t st) {
    const int tot = 2 * g.NY + 2 * g.NX;
    k_ring_init<<<(tot + 127) / 128, 128, 0, st>>>(g, rg, y, q, kdiag, inv_theta, rho, e);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

__global__ void k_ring_step(const MeshDev g, const Regions rg, const double* __restrict__ q, const double* __restrict__ kdiag,
                            double a, double b, double* __restrict__ z, double* __restrict__ rho, double* __restrict__ e) {
    int ix, iy;
    if (!ring_node(g, blockIdx.x * blockDim.x + threadIdx.x, ix, iy)) return;
    if (g.gx0 + ix == rg.pin_gx && iy == rg.pin_iy) return;
    const long long o = (long long)ix * g.LD + iy;
    const double ev = e[o];
    z[o] += ev;
    const double r = rho[o] - q[o];
    rho[o] = r;
    e[o] = a * ev + b * r / kdiag[o];
}
int ring_step(const MeshDev& g, Regions rg, const double* q, const double* kdiag, double a, double b, double* z, double* rho,
              double* e, cudaStream_t st) {
    const int tot = 2 * g.NY + 2 * g.NX;
    k_ring_step<<<(tot + 127) / 128, 128, 0, st>>>(g, rg, q, kdiag, a, b, z, rho, e);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

// ---- coarse-space projector ----------------------------------------------------------------------------------------------------
// 1-D ingredients of direction `dir` at local node index q (x: line ix of the slab, global line gx0 + ix; y: iy)
struct Pw1 {
    int P, n, ne;       // global nodes / elements of the direction
    int off;            // global index of local index 0
};
__device__ __forceinline__ double elem_sign(int P, int m) { return (P & 1) && (m & 1) ? -1.0 : 1.0; }

// dir 0: one thread per (k, iy); c[k][iy] with k the GLOBAL vertex index.  Only the nodes of this slab contribute (an interface
// line is counted by the left rank), so the partial sums of all ranks add up to the global restriction.
__global__ void k_pw_restrict_x(const MeshDev g, const TabDev t, const PwDir d, int transposed, const double* __restrict__ v,
                                double* __restrict__ c, int ldc) {
    const int iy = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;                         // global vertex 0 .. ne
    if (iy >= g.NY) return;
    const int P = g.P, ne = (d.n - 1) / P;
    double s = 0.0;
    if (iy >= 1 && iy <= g.NY - 2) {
        const int first = g.has_left ? 1 : 0;         // first local line this rank counts
        auto term = [&](int gnode, double wv) {
            const int ix = gnode - g.gx0;
            if (gnode < 1 || gnode > d.n - 2 || ix < first || ix >= g.NX) return;
            double x = v[(long long)ix * g.LD + iy];
            if (!transposed) x *= 0.5 * g.dx * asm_w1(t.w, P, gnode, ne);
            s = fma(wv, x, s);
        };
        if (k >= 1) {
            const double sg = elem_sign(P, k - 1);
            for (int j = 1; j < P; ++j) term((k - 1) * P + j, sg * d.wr[j]);
        }
        if (k < ne) {
            const double sg = elem_sign(P, k);
            term(k * P, sg * d.wl[0]);
            for (int j = 1; j < P; ++j) term(k * P + j, sg * d.wl[j]);
        } else {
            term(k * P, elem_sign(P, k - 1) * d.wr[P]);
        }
    }
    c[(long long)k * ldc + iy] = s;
}

// dir 1: one thread per (ix, k); c[ix][k]
__global__ void k_pw_restrict_y(const MeshDev g, const TabDev t, const PwDir d, int transposed, const double* __restrict__ v,
                                double* __restrict__ c, int ldc) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int ix = blockIdx.y;
    const int P = g.P, ne = g.ney;
    if (k > ne) return;
    const int gix = g.gx0 + ix;
    double s = 0.0;
    if (gix >= 1 && gix <= g.NXg - 2) {
        const double* row = v + (long long)ix * g.LD;
        auto term = [&](int node, double wv) {
            if (node < 1 || node > g.NY - 2) return;
            double x = row[node];
            if (!transposed) x *= 0.5 * g.dy * asm_w1(t.w, P, node, ne);
            s = fma(wv, x, s);
        };
        if (k >= 1) {
            const double sg = elem_sign(P, k - 1);
            for (int j = 1; j < P; ++j) term((k - 1) * P + j, sg * d.wr[j]);
        }
        if (k < ne) {
            const double sg = elem_sign(P, k);
            term(k * P, sg * d.wl[0]);
            for (int j = 1; j < P; ++j) term(k * P + j, sg * d.wl[j]);
        } else {
            term(k * P, elem_sign(P, k - 1) * d.wr[P]);
        }
    }
    c[(long long)ix * ldc + k] = s;
}

int pw_restrict(const MeshDev& g, TabDev t, const PwDir& d, int dir, int transposed, const double* v, double* c, int ldc,
                cudaStream_t st) {
    const int ne = (d.n - 1) / g.P;
    if (dir == 0) k_pw_restrict_x<<<dim3((unsigned)((g.NY + 127) / 128), (unsigned)(ne + 1)), 128, 0, st>>>(g, t, d, transposed, v, c, ldc);
    else k_pw_restrict_y<<<dim3((unsigned)((ne + 1 + 63) / 64), (unsigned)g.NX), 64, 0, st>>>(g, t, d, transposed, v, c, ldc);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

// out = v - W c' (transposed: v - M W c') on the interior nodes of the direction, out = v elsewhere.  One thread per node.
__global__ void k_pw_prolong(const MeshDev g, const TabDev t, const PwDir d, int dir, int transposed, const double* __restrict__ c,
                             int ldc, const double* v, double* out) {
    const int iy = blockIdx.x * blockDim.x + threadIdx.x;
    const int ix = blockIdx.y;
    if (iy >= g.NY) return;
    const long long o = (long long)ix * g.LD + iy;
    const int P = g.P;
    const int gix = g.gx0 + ix;
    double val = v[o];
    const bool interior2d = gix >= 1 && gix <= g.NXg - 2 && iy >= 1 && iy <= g.NY - 2;
    if (interior2d) {
        const int node = dir == 0 ? gix : iy;
        const int ne = dir == 0 ? (g.NXg - 1) / P : g.ney;
        const int j = node % P;
        double pv;
        auto cval = [&](int k) { return dir == 0 ? c[(long long)k * ldc + iy] : c[(long long)ix * ldc + k]; };
        if (j == 0) {
            const int k = node / P;   // interior vertex: 0 < k < ne
            pv = elem_sign(P, k) * d.wl[0] * cval(k);
        } else {
            const int m = node / P;
            pv = elem_sign(P, m) * (d.wl[j] * cval(m) + d.wr[j] * cval(m + 1));
        }
        if (transposed) pv *= 0.5 * (dir == 0 ? g.dx : g.dy) * asm_w1(t.w, P, node, ne);
        val -= pv;
    }
    out[o] = val;
}
int pw_prolong(const MeshDev& g, TabDev t, const PwDir& d, int dir, int transposed, const double* c, int ldc, const double* v,
               double* out, cudaStream_t st) {
    k_pw_prolong<<<dim3((unsigned)((g.NY + 127) / 128), (unsigned)g.NX), 128, 0, st>>>(g, t, d, dir, transposed, c, ldc, v, out);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

// ---- member selection --------------------------------------------------------------------------------------------------------------
__global__ void k_member_vectors(const MeshDev g, const TabDev t, const Regions rg, const double* __restrict__ lfx,
                                 const double* __restrict__ lfy, double* __restrict__ lc, double* __restrict__ mc) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)g.NX * g.LD) return;
    const int ix = (int)(idx / g.LD), iy = (int)(idx % g.LD);
    double l = 0.0, m = 0.0;
    if (iy < g.NY) {
        const int gix = g.gx0 + ix;
        l = lfx[gix] * lfy[iy];
        m = (0.5 * g.dx * asm_w1(t.w, g.P, gix, (g.NXg - 1) / g.P)) * (0.5 * g.dy * asm_w1(t.w, g.P, iy, g.ney));
        if (gix == rg.pin_gx && iy == rg.pin_iy) m = 1.0;   // M_p: 1 at the pin (NS:208-212)
        m *= l;
    }
    lc[idx] = l;
    mc[idx] = m;
}
int member_vectors(const MeshDev& g, TabDev t, Regions rg, const double* lfx, const double* lfy, double* lc, double* mc,
                   cudaStream_t st) {
    const long long tot = (long long)g.NX * g.LD;
    k_member_vectors<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g, t, rg, lfx, lfy, lc, mc);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

__global__ void k_member_update(long long n, const double* __restrict__ lc, const double* __restrict__ sums, double inv_den,
                                double* __restrict__ z) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const double coef = (sums[0] - sums[1]) * inv_den;
    z[idx] = fma(-coef, lc[idx], z[idx]);
}
int member_update(const MeshDev& g, const double* lc, const double* sums, double inv_den, double* z, cudaStream_t st) {
    const long long tot = (long long)g.NX * g.LD;
    k_member_update<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(tot, lc, sums, inv_den, z);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace semb
