// Auxiliary kernels of sem_b200 (sm_100a): pointwise operators, boundary rows, standalone gather-scatter and the
// deterministic reductions / vector updates of the Krylov solver.  All of them are plain HBM-streaming kernels.
#include "sem_aux.cuh"

namespace semb {

// ---------------------------------------------------------------------------------------------------------------
// helpers (runtime polynomial order)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double asm_w(const double* __restrict__ w, int P, int q, int nel) {
    const int j = q % P;
    if (j != 0) return w[j];
    double s = 0.0;
    if (q > 0) s += w[P];
    if (q < nel * P) s += w[0];
    return s;
}

__device__ __forceinline__ double asm_kdiag(const double* __restrict__ Ks, int P, int q, int nel) {
    const int j = q % P;
    const int n = P + 1;
    if (j != 0) return Ks[j * n + j];
    double s = 0.0;
    if (q > 0) s += Ks[P * n + P];
    if (q < nel * P) s += Ks[0];
    return s;
}

// ---------------------------------------------------------------------------------------------------------------
// mass matrix (diagonal): M = assemble((dx/2 w_i)(dy/2 w_j))       SEM.py:170-183
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_mass(const MeshDev g, const TabDev t, const double* __restrict__ x, double* __restrict__ y) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long tot = (long long)g.NX * g.LD;
    if (idx >= tot) return;
    const int ix = (int)(idx / g.LD), iy = (int)(idx % g.LD);
    double v = 0.0;
    if (iy < g.NY) {
        v = (0.5 * g.dx * asm_w(t.w, g.P, ix, g.nex)) * (0.5 * g.dy * asm_w(t.w, g.P, iy, g.ney));
        if (x) v *= x[idx];
    }
    y[idx] = v;
}

int aux_mass_apply(const MeshDev& g, TabDev t, const double* x, double* y, cudaStream_t st) {
    const long long tot = (long long)g.NX * g.LD;
    k_mass<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g, t, x, y);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

__global__ void k_kdiag(const MeshDev g, const TabDev t, double* __restrict__ d) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long tot = (long long)g.NX * g.LD;
    if (idx >= tot) return;
    const int ix = (int)(idx / g.LD), iy = (int)(idx % g.LD);
    double v = 0.0;
    if (iy < g.NY) {
        // diagonal of the GLOBAL assembled K: global line index, so interface lines carry both ranks' elements
        const int gix = g.gx0 + ix, nexg = (g.NXg - 1) / g.P;
        const double wxA = 0.5 * g.dx * asm_w(t.w, g.P, gix, nexg);
        const double wyA = 0.5 * g.dy * asm_w(t.w, g.P, iy, g.ney);
        v = wyA * (2.0 / g.dx) * asm_kdiag(t.Ks, g.P, gix, nexg) + wxA * (2.0 / g.dy) * asm_kdiag(t.Ks, g.P, iy, g.ney);
    }
    d[idx] = v;
}

int aux_stiffness_diag(const MeshDev& g, TabDev t, double* d, cudaStream_t st) {
    const long long tot = (long long)g.NX * g.LD;
    k_kdiag<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g, t, d);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// pressure-Neumann rows: y[b] = (K c)[b] on the boundary nodes b of the (global) domain        NS:119, NS:157
// one thread per boundary node; O(sqrt(N)) work
// ---------------------------------------------------------------------------------------------------------------
__device__ double line_stiff(const double* __restrict__ Ks, int P, const double* __restrict__ c, long long stride,
                             int q, int nel) {
    // sum over the element(s) of a 1-D line that contain node q of row(q) . c   (c addressed as c[k*stride])
    const int n = P + 1;
    const int j = q % P;
    double s = 0.0;
    if (j != 0) {
        const int e = q / P;
        for (int k = 0; k <= P; ++k) s = fma(Ks[j * n + k], c[(long long)(e * P + k) * stride], s);
    } else {
        if (q > 0) {
            const int e = q / P - 1;
            for (int k = 0; k <= P; ++k) s = fma(Ks[P * n + k], c[(long long)(e * P + k) * stride], s);
        }
        if (q < nel * P) {
            const int e = q / P;
            double s2 = 0.0;
            for (int k = 0; k <= P; ++k) s2 = fma(Ks[k], c[(long long)(e * P + k) * stride], s2);
            s += s2;
        }
    }
    return s;
}

__global__ void k_neumann(const MeshDev g, const TabDev t, const double* __restrict__ c, double* __restrict__ y,
                          int pin_gx, int pin_iy, int skip_pin) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int ix, iy;
    if (idx < g.NY) {
        if (g.has_left) return;
        ix = 0; iy = idx;
    } else if (idx < 2 * g.NY) {
        if (g.has_right) return;
        ix = g.NX - 1; iy = idx - g.NY;
    } else if (idx < 2 * g.NY + 2 * g.NX) {
        const int r = idx - 2 * g.NY;
        ix = r % g.NX;
        iy = (r < g.NX) ? 0 : g.NY - 1;
        if ((ix == 0 && !g.has_left) || (ix == g.NX - 1 && !g.has_right)) return;   // corners done by W / E
    } else {
        return;
    }
    if (skip_pin && g.gx0 + ix == pin_gx && iy == pin_iy) return;
    const double wxA = 0.5 * g.dx * asm_w(t.w, g.P, ix, g.nex);
    const double wyA = 0.5 * g.dy * asm_w(t.w, g.P, iy, g.ney);
    const double sx = line_stiff(t.Ks, g.P, c + iy, g.LD, ix, g.nex);
    const double sy = line_stiff(t.Ks, g.P, c + (long long)ix * g.LD, 1, iy, g.ney);
    y[(long long)ix * g.LD + iy] = wyA * (2.0 / g.dx) * sx + wxA * (2.0 / g.dy) * sy;
}

int aux_neumann_rows(const MeshDev& g, TabDev t, const double* c, double* y, int pin_gx, int pin_iy, int skip_pin,
                     cudaStream_t st) {
    const int tot = 2 * g.NY + 2 * g.NX;
    k_neumann<<<(tot + 127) / 128, 128, 0, st>>>(g, t, c, y, pin_gx, pin_iy, skip_pin);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// dense <-> padded repacking
// --------------------------------------------------------------------------------------------
template <bool PAD>
__global__ void k_repack(const MeshDev g, const double* __restrict__ src, double* __restrict__ dst, long long first,
                         long long count) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const long long idx = first + k;
    const long long ix = idx / g.NY, iy = idx % g.NY;
    if (PAD) dst[ix * g.LD + iy] = src[idx];
    else dst[idx] = src[ix * g.LD + iy];
}

int aux_pad_lines(const MeshDev& g, const double* dense, double* vec, int line0, int nlines, cudaStream_t st) {
    const long long count = (long long)nlines * g.NY;
    if (count <= 0) return 0;
    k_repack<true><<<(unsigned)((count + 255) / 256), 256, 0, st>>>(g, dense, vec, (long long)line0 * g.NY, count);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

int aux_unpad_lines(const MeshDev& g, const double* vec, double* dense, int line0, int nlines, cudaStream_t st) {
    const long long count = (long long)nlines * g.NY;
    if (count <= 0) return 0;
    k_repack<false><<<(unsigned)((count + 255) / 256), 256, 0, st>>>(g, vec, dense, (long long)line0 * g.NY, count);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

int aux_pad(const MeshDev& g, const double* dense, double* vec, cudaStream_t st) {
    return aux_pad_lines(g, dense, vec, 0, g.NX, st);
}

int aux_unpad(const MeshDev& g, const double* vec, double* dense, cudaStream_t st) {
    return aux_unpad_lines(g, vec, dense, 0, g.NX, st);
}

// ---------------------------------------------------------------------------------------------------------------
// standalone gather-scatter (SEM.assemble for 4-index arrays, SEM.py:126-131) -- colour ordered.
// Elements are 4-coloured by the parity of (m, n); elements of one colour share no node, so each colour pass is a
// plain read-modify-write without atomics and the four passes run in a fixed order: the sum at every shared node
// is formed in the same order on every run (bitwise reproducible).
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_gs_colour(const MeshDev g, const double* __restrict__ elem, double* __restrict__ y, int cx, int cy,
                            int ncx, int ncy) {
    const int n = g.P + 1;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long tot = (long long)ncx * ncy * n * n;
    if (idx >= tot) return;
    const int j = (int)(idx % n);
    const int i = (int)((idx / n) % n);
    const long long e = idx / (n * n);
    const int en = (int)(e % ncy), em = (int)(e / ncy);
    const int m = 2 * em + cx, nn = 2 * en + cy;
    const double v = elem[(((long long)m * g.ney + nn) * n + i) * n + j];
    y[(long long)(m * g.P + i) * g.LD + nn * g.P + j] += v;
}

int aux_gather_scatter(const MeshDev& g, const double* elem, double* y, cudaStream_t st) {
    SEM_CUDA(cudaMemsetAsync(y, 0, sizeof(double) * (size_t)g.NX * g.LD, st));
    const int n = g.P + 1;
    for (int cx = 0; cx < 2; ++cx)
        for (int cy = 0; cy < 2; ++cy) {
            const int ncx = (g.nex - cx + 1) / 2, ncy = (g.ney - cy + 1) / 2;
            const long long tot = (long long)ncx * ncy * n * n;
            if (tot <= 0) continue;
            k_gs_colour<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g, elem, y, cx, cy, ncx, ncy);
            SEM_CUDA(cudaGetLastError());
        }
    return 0;
}

__global__ void k_scatter(const MeshDev g, const double* __restrict__ x, double* __restrict__ elem) {
    const int n = g.P + 1;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long tot = (long long)g.nex * g.ney * n * n;
    if (idx >= tot) return;
    const int j = (int)(idx % n);
    const int i = (int)((idx / n) % n);
    const long long e = idx / (n * n);
    const int nn = (int)(e % g.ney), m = (int)(e / g.ney);
    elem[idx] = x[(long long)(m * g.P + i) * g.LD + nn * g.P + j];
}

int aux_scatter(const MeshDev& g, const double* x, double* elem, cudaStream_t st) {
    const int n = g.P + 1;
    const long long tot = (long long)g.nex * g.ney * n * n;
    k_scatter<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g, x, elem);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Krylov building blocks
// ---------------------------------------------------------------------------------------------------------------
constexpr int DOT_JT = 8;        // basis vectors per block column
constexpr int DOT_THREADS = 256;

// partial sums of <V_j, w> over this block's share of the vector; last block of each column reduces in fixed order
__global__ void __launch_bounds__(DOT_THREADS) k_multi_dot(const double* __restrict__ V, long long n, int k,
                                                          const double* __restrict__ w, double* __restrict__ h,
                                                          int nf, long long vlen, long long skip,
                                                          double* __restrict__ partials, unsigned* __restrict__ counter) {
    const int jt = blockIdx.y;
    const int j0 = jt * DOT_JT;
    const int nj = min(DOT_JT, k - j0);
    const int nb = gridDim.x;
    double acc[DOT_JT];
#pragma unroll
    for (int q = 0; q < DOT_JT; ++q) acc[q] = 0.0;
    for (int f = 0; f < nf; ++f) {
        const long long base = (long long)f * vlen;
        for (long long e = skip + (long long)blockIdx.x * DOT_THREADS + threadIdx.x; e < vlen;
             e += (long long)nb * DOT_THREADS) {
            const double wv = w[base + e];
#pragma unroll
            for (int q = 0; q < DOT_JT; ++q)
                if (q < nj) acc[q] = fma(V[(long long)(j0 + q) * n + base + e], wv, acc[q]);
        }
    }
    __shared__ double red[DOT_JT][DOT_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < DOT_JT; ++q) {
        double v = acc[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[q][wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < DOT_JT) {
        double v = 0.0;
        for (int q = 0; q < DOT_THREADS / 32; ++q) v += red[threadIdx.x][q];
        partials[((long long)jt * DOT_JT + threadIdx.x) * nb + blockIdx.x] = v;
    }
    __shared__ unsigned ticket;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) ticket = atomicAdd(&counter[jt], 1u);
    __syncthreads();
    if (ticket == (unsigned)(nb - 1)) {   // last block of this column: fixed-order final sum
        __threadfence();
        if (threadIdx.x < nj) {
            const double* p = partials + ((long long)jt * DOT_JT + threadIdx.x) * nb;
            double v = 0.0;
            for (int b = 0; b < nb; ++b) v += p[b];
            h[j0 + threadIdx.x] = v;
        }
        if (threadIdx.x == 0) counter[jt] = 0u;
    }
}

int aux_multi_dot(const double* V, long long n, int k, const double* w, double* h, int nf, long long vlen,
                  long long skip, RedScratch rs, cudaStream_t st) {
    if (k <= 0) return 0;
    if (k > rs.max_k) { set_error("aux_multi_dot: k exceeds the reduction scratch"); return -2; }
    long long want = (vlen - skip + (long long)DOT_THREADS * 4 - 1) / ((long long)DOT_THREADS * 4);
    int nb = (int)(want < 1 ? 1 : (want > rs.max_blocks ? rs.max_blocks : want));
    dim3 grid((unsigned)nb, (unsigned)((k + DOT_JT - 1) / DOT_JT));
    k_multi_dot<<<grid, DOT_THREADS, 0, st>>>(V, n, k, w, h, nf, vlen, skip, rs.partials, rs.counter);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

constexpr int AXPY_THREADS = 256;
constexpr int AXPY_HT = 128;   // coefficients staged per tile

// acc = sum over the basis vectors j in [ja, jb) of h[j] * V_j[e]  (fixed order; two accumulators for load parallelism)
__device__ __forceinline__ double axpy_partial(const double* __restrict__ V, long long n, const double* __restrict__ h,
                                               int ja, int jb, long long e, bool active, double* sh) {
    double a0 = 0.0, a1 = 0.0;
    for (int j0 = ja; j0 < jb; j0 += AXPY_HT) {
        const int nj = min(AXPY_HT, jb - j0);
        __syncthreads();
        if (threadIdx.x < nj) sh[threadIdx.x] = h[j0 + threadIdx.x];
        __syncthreads();
        if (active) {
            int q = 0;
#pragma unroll 4
            for (; q + 1 < nj; q += 2) {
                a0 = fma(sh[q], V[(long long)(j0 + q) * n + e], a0);
                a1 = fma(sh[q + 1], V[(long long)(j0 + q + 1) * n + e], a1);
            }
            if (q < nj) a0 = fma(sh[q], V[(long long)(j0 + q) * n + e], a0);
        }
    }
    return a0 + a1;
}

// w -= / += V h (COMB: w = V h).  grid.y = J slices of the basis: on the reference's small meshes one thread per element
// leaves the GPU almost empty (15 k elements, 600 vectors: 37 us), so the sum over j is split, the slices' partial sums go
// to scratch and the last slice to finish an element block (integer ticket) adds them in slice order -- deterministic.
template <bool COMB>
__global__ void __launch_bounds__(AXPY_THREADS) k_multi_axpy(const double* __restrict__ V, long long n, int k,
                                                             const double* __restrict__ h, double sign,
                                                             double* __restrict__ w, double* __restrict__ partials,
                                                             unsigned* __restrict__ counter) {
    __shared__ double sh[AXPY_HT];
    const long long e = (long long)blockIdx.x * AXPY_THREADS + threadIdx.x;
    const int J = gridDim.y, jc = blockIdx.y;
    const int per = (k + J - 1) / J;
    const int ja = min(k, jc * per), jb = min(k, ja + per);
    const double acc = axpy_partial(V, n, h, ja, jb, e, e < n, sh);
    if (J == 1) {
        if (e < n) w[e] = COMB ? acc : fma(sign, acc, w[e]);
        return;
    }
    if (e < n) partials[(long long)jc * n + e] = acc;
    __shared__ unsigned ticket;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) ticket = atomicAdd(&counter[blockIdx.x], 1u);
    __syncthreads();
    if (ticket == (unsigned)(J - 1)) {
        __threadfence();
        if (e < n) {
            double v = 0.0;
            for (int q = 0; q < J; ++q) v += partials[(long long)q * n + e];
            w[e] = COMB ? v : fma(sign, v, w[e]);
        }
        if (threadIdx.x == 0) counter[blockIdx.x] = 0u;
    }
}

template <bool COMB>
static int launch_multi_axpy(const double* V, long long n, int k, const double* h, double sign, double* w, RedScratch rs,
                             cudaStream_t st) {
    const long long nb = (n + AXPY_THREADS - 1) / AXPY_THREADS;
    // enough slices for ~4 blocks per SM, at least 32 vectors per slice, bounded by the scratch
    long long J = (4ll * rs.sm_count + nb - 1) / nb;
    if (J > (k + 31) / 32) J = (k + 31) / 32;
    if (J > 16) J = 16;
    if (J < 1 || nb > rs.axpy_blocks || J * n > rs.axpy_len) J = 1;
    dim3 grid((unsigned)nb, (unsigned)J);
    k_multi_axpy<COMB><<<grid, AXPY_THREADS, 0, st>>>(V, n, k, h, sign, w, rs.axpy_partials, rs.axpy_counter);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

int aux_multi_axpy(const double* V, long long n, int k, const double* h, double sign, double* w, RedScratch rs,
                   cudaStream_t st) {
    if (k <= 0) return 0;
    return launch_multi_axpy<false>(V, n, k, h, sign, w, rs, st);
}

int aux_multi_comb(const double* V, long long n, int k, const double* h, double* out, RedScratch rs, cudaStream_t st) {
    return launch_multi_axpy<true>(V, n, k, h, 1.0, out, rs, st);
}

__global__ void k_scale_inv_norm(const double* __restrict__ w, const double* __restrict__ nrm2, double* __restrict__ v,
                                 double* __restrict__ v2, long long n) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const double s = 1.0 / sqrt(*nrm2);
    const double r = w[e] * s;
    v[e] = r;
    if (v2) v2[e] = r;
}

int aux_scale_inv_norm(const double* w, const double* nrm2, double* v, double* v2, long long n, cudaStream_t st) {
    k_scale_inv_norm<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w, nrm2, v, v2, n);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

__global__ void k_axpby(double a, const double* __restrict__ x, double b, double* __restrict__ y, long long n) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    y[e] = (b == 0.0) ? a * x[e] : fma(a, x[e], b * y[e]);
}

int aux_axpby(double a, const double* x, double b, double* y, long long n, cudaStream_t st) {
    k_axpby<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, x, b, y, n);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// preconditioner pieces
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_cd_jacobi(const MeshDev g, const BCSpec bc, const double* __restrict__ dK,
                            const double* __restrict__ r, double* __restrict__ z) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long tot = (long long)g.NX * g.LD;
    if (idx >= tot) return;
    const int ix = (int)(idx / g.LD), iy = (int)(idx % g.LD);
    double v = 0.0;
    if (iy < g.NY) {
        const int side = bc_side(bc, g.gx0 + ix, iy, g.NXg, g.NY);
        v = (side >= 0) ? r[idx] : r[idx] / dK[idx];
    }
    z[idx] = v;
}

int aux_cd_jacobi(const MeshDev& g, const BCSpec& bc, const double* dK, const double* r, double* z, cudaStream_t st) {
    const long long tot = (long long)g.NX * g.LD;
    k_cd_jacobi<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g, bc, dK, r, z);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

__device__ __forceinline__ double safe_diag(double dk, double react) {
    const double d = dk + react;
    // the reaction term Re*G u may cancel the stiffness diagonal; never divide by something smaller than 10% of it
    return (fabs(d) >= 0.1 * dk) ? d : dk;
}

__global__ void k_ns_jacobi(const MeshDev g, const double* __restrict__ dK, const double* __restrict__ gxu,
                            const double* __restrict__ gyv, const double* __restrict__ ru,
                            const double* __restrict__ rv, double* __restrict__ zu, double* __restrict__ zv) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long tot = (long long)g.NX * g.LD;
    if (idx >= tot) return;
    const int ix = (int)(idx / g.LD), iy = (int)(idx % g.LD);
    double a = 0.0, b = 0.0;
    if (iy < g.NY) {
        const int gix = g.gx0 + ix;
        const bool bnd = (gix == 0) || (gix == g.NXg - 1) || (iy == 0) || (iy == g.NY - 1);
        if (bnd) {
            a = ru[idx];
            b = rv[idx];
        } else {
            a = ru[idx] / safe_diag(dK[idx], gxu[idx]);
            b = rv[idx] / safe_diag(dK[idx], gyv[idx]);
        }
    }
    zu[idx] = a;
    zv[idx] = b;
}

int aux_ns_jacobi(const MeshDev& g, const double* dK, const double* gxu, const double* gyv, const double* ru,
                  const double* rv, double* zu, double* zv, cudaStream_t st) {
    const long long tot = (long long)g.NX * g.LD;
    k_ns_jacobi<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g, dK, gxu, gyv, ru, rv, zu, zv);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

__global__ void k_ns_schur_mass(const MeshDev g, const TabDev t, const double* __restrict__ rc,
                                const double* __restrict__ div, double* __restrict__ zp, int pin_gx, int pin_iy) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long tot = (long long)g.NX * g.LD;
    if (idx >= tot) return;
    const int ix = (int)(idx / g.LD), iy = (int)(idx % g.LD);
    double v = 0.0;
    if (iy < g.NY) {
        const int gix = g.gx0 + ix;
        const bool bnd = (gix == 0) || (gix == g.NXg - 1) || (iy == 0) || (iy == g.NY - 1);
        const bool pin = (gix == pin_gx) && (iy == pin_iy);
        const double m = (0.5 * g.dx * asm_w(t.w, g.P, gix, (g.NXg - 1) / g.P)) * (0.5 * g.dy * asm_w(t.w, g.P, iy, g.ney));
        v = rc[idx];   // m: GLOBAL assembled mass
        if (!bnd && !pin) v -= div[idx];
        if (!pin) v /= m;
    }
    zp[idx] = v;
}

int aux_ns_schur_mass(const MeshDev& g, TabDev t, const double* rc, const double* div, double* zp, int pin_gx,
                      int pin_iy, cudaStream_t st) {
    const long long tot = (long long)g.NX * g.LD;
    k_ns_schur_mass<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g, t, rc, div, zp, pin_gx, pin_iy);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

__global__ void k_sub(const double* a, const double* b, double* out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] - b[i];
}
int aux_sub(const double* a, const double* b, double* out, long long n, cudaStream_t st) {
    k_sub<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, b, out, n);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Tensor-product interpolation (SEM.eval_interpolation, SEM.py:248-273; the mesh-to-mesh transfer `change_inputs` of the
// OpenMDAO components, CD_Component.py:23-36): out[a][b] = sum_ij Sx[a][i] Sy[b][j] f[(mx[a] - m0) P + i][ny[b] P + j] for the
// plot columns a whose element column mx[a] (GLOBAL index, from x2xi SEM.py:23-36) lies in this slab, 0 for the others (a
// partitioned mesh sums the ranks' arrays).  One thread per output value, fixed summation order.
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_interpolate(const MeshDev g, const double* __restrict__ f, int nxp, const int* __restrict__ mx,
                              const double* __restrict__ Sx, int nyp, const int* __restrict__ ny, const double* __restrict__ Sy,
                              double* __restrict__ out, int ldo) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int a = blockIdx.y;
    if (b >= nyp) return;
    const int n1 = g.P + 1;
    const int m = mx[a] - g.gx0 / g.P;
    double val = 0.0;
    if (m >= 0 && m < g.nex) {
        const double* base = f + (size_t)(m * g.P) * g.LD + (size_t)ny[b] * g.P;
        for (int i = 0; i < n1; ++i) {
            double row = 0.0;
            for (int j = 0; j < n1; ++j) row = fma(Sy[b * n1 + j], base[(size_t)i * g.LD + j], row);
            val = fma(Sx[a * n1 + i], row, val);
        }
    }
    out[(size_t)a * ldo + b] = val;
}
int aux_interpolate(const MeshDev& g, const double* f, int nxp, const int* mx, const double* Sx, int nyp, const int* ny,
                    const double* Sy, double* out, int ldo, cudaStream_t st) {
    if (nxp <= 0 || nyp <= 0) return 0;
    k_interpolate<<<dim3((unsigned)((nyp + 127) / 128), (unsigned)nxp), 128, 0, st>>>(g, f, nxp, mx, Sx, nyp, ny, Sy, out, ldo);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

// y += scale * (*coef) * x with the coefficient on the device (projection b - l (l.b) / (l.l) without a host round trip)
__global__ void k_axpy_dev(long long n, const double* __restrict__ x, const double* __restrict__ coef, double scale, double* __restrict__ y) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = fma(scale * *coef, x[i], y[i]);
}
int aux_axpy_dev(long long n, const double* x, const double* coef, double scale, double* y, cudaStream_t st) {
    k_axpy_dev<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, x, coef, scale, y);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace semb
