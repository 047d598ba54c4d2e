// Fused matrix-free SEM operator apply for sm_100a: sum-factorised element contractions + on-chip gather-scatter.
//
// Replaces the reference's "assemble a CSR matrix on the host, then csr_matvec" path
// (SEM.py:170-245 + CD:73-121 / NS:93-160) by one kernel that never forms a matrix.
//
// Decomposition.  The uniform Cartesian element grid makes every element matrix a Kronecker product of the 1-D GLL
// tables (SEM.py:182,201-202,221-222), so an element apply is two 1-D contractions per table:
//     (K u)_e   = (2/dx) (Ks U) o (dy/2 1 w^T)  +  (dx/2 w 1^T) o (2/dy) (U Ks^T)
//     (G_x u)_e = (w o D U) (dy/2 w^T),   (G_y u)_e = (dx/2 w) (w o U D^T)
// and the gather-scatter of shared GLL nodes (SEM.assemble, SEM.py:113-146) is a sum over the (at most) two elements
// that share a node in each direction.
//
// Mapping.  A CTA owns a strip of Ty element rows (Ty*P owned nodes in y, contiguous in memory) and MARCHES over
// Mx element columns in x.  One thread per node column iy:
//   * x-contractions run in registers: the thread keeps the P+1 values of its element line, reuses the shared line
//     between consecutive elements, and carries the partial sum of the shared node (deterministic 2-term add);
//   * y-contractions run through shared memory: the P new node lines of a step are staged once ([P][pitch] tile,
//     pitch == 1 mod 16 doubles => conflict-free for both the row-wise and the column-wise accesses), one thread
//     per (line, element) contracts the P+1 values of its element and the results are assembled in shared memory
//     in two colour phases (all "lower" contributions are written, barrier, all "upper" contributions are added) --
//     atomic-free and bitwise reproducible.
//   * across strips the owner of a node computes it completely: a strip also stages the P nodes below it and
//     evaluates the single row of the element underneath that reaches its lowest owned node (y-halo, 1/Ty extra
//     reads); across marching chunks the first chunk column is primed from the P lines to its left (x-halo, 1/Mx).
//   Every global value is therefore read ~once and written exactly once, with no atomics and no second pass.
//
// Table operands (D, Ks, w) come from constant memory with compile-time offsets (P is a template parameter), so the
// DFMA stream needs no shared-memory bandwidth for them.
#pragma once
#include "sem_common.cuh"

#ifndef SEM_MARCH_MAXT
#define SEM_MARCH_MAXT 320   // upper bound of the CTA size chosen by march_geometry()
#endif

namespace semb {

template <int MODE> struct ModeTraits;
// NF contracted fields, NV staged advecting component, NACC y-accumulators, NOUT outputs, MINB resident CTAs/SM
// the register allocation is held to (ptxas -v: no spills at these bounds for P <= 8).
template <> struct ModeTraits<MODE_K>  { static constexpr int NF = 1, NV = 0, NACC = 1, NOUT = 1, MINB = 3; };
template <> struct ModeTraits<MODE_G>  { static constexpr int NF = 1, NV = 0, NACC = 1, NOUT = 2, MINB = 2; };
template <> struct ModeTraits<MODE_CD> { static constexpr int NF = 1, NV = 1, NACC = 1, NOUT = 1, MINB = 2; };
template <> struct ModeTraits<MODE_NS> { static constexpr int NF = 3, NV = 1, NACC = 3, NOUT = 3, MINB = 1; };
template <> struct ModeTraits<MODE_DIV> { static constexpr int NF = 2, NV = 0, NACC = 1, NOUT = 1, MINB = 2; };

// shared memory doubles per CTA
template <int P, int MODE>
__host__ __device__ constexpr size_t march_smem_doubles(int pitch) {
    return (size_t)(ModeTraits<MODE>::NF + ModeTraits<MODE>::NV + ModeTraits<MODE>::NACC) * P * pitch;
}

// 1-D contraction of an element line with row I of a table (compile-time offsets into constant memory)
// Opaque zero.  ptxas otherwise hoists all (P+1)^2 table loads out of the marching loop into registers (LICM), which
// costs ~100 registers at P=8; an offset it cannot prove loop-invariant keeps them as just-in-time LDCU operands.
__device__ __forceinline__ int opaque_zero(int zero_param) {
    int z;
    asm volatile("mov.u32 %0, %1;" : "=r"(z) : "r"(zero_param));
    return z;
}

template <int P, int I>
__device__ __forceinline__ double row_dot(const double* __restrict__ tab, const double (&r)[P + 1]) {
    constexpr int NP = Tab<P>::NP;
    const double2* __restrict__ t2 = reinterpret_cast<const double2*>(tab + I * NP);   // 16-byte aligned row
    double s = 0.0;
#pragma unroll
    for (int kk = 0; kk < NP / 2; ++kk) {
        const double2 c = t2[kk];
        s = (kk == 0) ? c.x * r[0] : fma(c.x, r[2 * kk], s);
        if (2 * kk + 1 <= P) s = fma(c.y, r[2 * kk + 1], s);
    }
    return s;
}

// assembled 1-D quadrature weight of node q (0 <= q <= nel*P) of a line of nel elements (without the h/2 factor)
template <int P>
__device__ __forceinline__ double asm_weight(int q, int nel) {
    const int j = q % P;
    if (j != 0) return c_tab<P>.w[j];
    double s = 0.0;
    if (q > 0) s += c_tab<P>.w[P];
    if (q < nel * P) s += c_tab<P>.w[0];
    return s;
}

template <int P, int MODE>
struct March {
    using TR = ModeTraits<MODE>;
    static constexpr int n = P + 1;
    static constexpr int NF = TR::NF, NV = TR::NV, NACC = TR::NACC, NOUT = TR::NOUT;

    // ---- x-contraction of row I: contributions of ONE element to the NOUT outputs at its node (I, iy) ----------
    template <int I>
    static __device__ __forceinline__ void xrow(const double (&r)[NF][n], double Uc, double cKx, double wyA,
                                                double cc, double (&x)[NOUT], int z) {
        const double* tD = c_tab<P>.D + 2 * z;
        const double* tK = c_tab<P>.Ks + 2 * z;
        const double gw = c_tab<P>.w[I] * wyA;
        if constexpr (MODE == MODE_K) {
            x[0] = cKx * row_dot<P, I>(tK, r[0]);
        } else if constexpr (MODE == MODE_G) {
            x[0] = cc * gw * row_dot<P, I>(tD, r[0]);
            x[1] = 0.0;
        } else if constexpr (MODE == MODE_CD) {
            x[0] = fma(cc * Uc * gw, row_dot<P, I>(tD, r[0]), cKx * row_dot<P, I>(tK, r[0]));
        } else if constexpr (MODE == MODE_DIV) {
            x[0] = gw * row_dot<P, I>(tD, r[0]);
        } else {
            const double sDa = row_dot<P, I>(tD, r[0]);
            const double sDb = row_dot<P, I>(tD, r[1]);
            const double sDc = row_dot<P, I>(tD, r[2]);
            const double cu = cc * Uc * gw;
            x[0] = fma(gw, sDc, fma(cu, sDa, cKx * row_dot<P, I>(tK, r[0])));
            x[1] = fma(cu, sDb, cKx * row_dot<P, I>(tK, r[1]));
            x[2] = gw * sDa;
        }
    }

    // ---- y-contraction of row J of one element line (values l[f][0..P] along y): contributions to the NACC
    //      shared-memory accumulators at node (line, J) -------------------------------------------------------------
    template <int J>
    static __device__ __forceinline__ void yrow(const double (&l)[NF][n], double Vc, double wxK, double wxA,
                                                double cc, double (&y)[NACC], int z) {
        const double* tD = c_tab<P>.D + 2 * z;
        const double* tK = c_tab<P>.Ks + 2 * z;
        const double gw = wxA * c_tab<P>.w[J];
        if constexpr (MODE == MODE_K) {
            y[0] = wxK * row_dot<P, J>(tK, l[0]);
        } else if constexpr (MODE == MODE_G) {
            y[0] = cc * gw * row_dot<P, J>(tD, l[0]);
        } else if constexpr (MODE == MODE_CD) {
            y[0] = fma(cc * Vc * gw, row_dot<P, J>(tD, l[0]), wxK * row_dot<P, J>(tK, l[0]));
        } else if constexpr (MODE == MODE_DIV) {
            y[0] = gw * row_dot<P, J>(tD, l[1]);
        } else {
            const double sDa = row_dot<P, J>(tD, l[0]);
            const double sDb = row_dot<P, J>(tD, l[1]);
            const double sDc = row_dot<P, J>(tD, l[2]);
            const double cv = cc * Vc * gw;
            y[0] = fma(cv, sDa, wxK * row_dot<P, J>(tK, l[0]));
            y[1] = fma(gw, sDc, fma(cv, sDb, wxK * row_dot<P, J>(tK, l[1])));
            y[2] = gw * sDb;
        }
    }

    // ---- the same row formulas with the contraction results passed in (used by the v2 kernel, which computes the
    //      contractions with folded even/odd tables): sK[f] = (Ks a_f)_row, sD[f] = (D a_f)_row -------------------------
    static __host__ __device__ constexpr bool x_needs_K(int f) { return (MODE == MODE_K || MODE == MODE_CD) ? f == 0 : (MODE == MODE_NS ? f < 2 : false); }
    static __host__ __device__ constexpr bool x_needs_D(int f) { return (MODE == MODE_NS) ? true : (MODE == MODE_K ? false : f == 0); }
    static __host__ __device__ constexpr bool y_needs_K(int f) { return x_needs_K(f); }
    static __host__ __device__ constexpr bool y_needs_D(int f) { return (MODE == MODE_NS) ? true : (MODE == MODE_K ? false : (MODE == MODE_DIV ? f == 1 : f == 0)); }

    template <int I>
    static __device__ __forceinline__ void xcombine(const double (&sK)[NF], const double (&sD)[NF], double Uc,
                                                    double cKx, double wyA, double cc, double (&x)[NOUT]) {
        const double gw = c_tab<P>.w[I] * wyA;
        if constexpr (MODE == MODE_K) {
            x[0] = cKx * sK[0];
        } else if constexpr (MODE == MODE_G) {
            x[0] = cc * gw * sD[0];
            x[1] = 0.0;
        } else if constexpr (MODE == MODE_CD) {
            x[0] = fma(cc * Uc * gw, sD[0], cKx * sK[0]);
        } else if constexpr (MODE == MODE_DIV) {
            x[0] = gw * sD[0];
        } else {
            const double cu = cc * Uc * gw;
            x[0] = fma(gw, sD[2], fma(cu, sD[0], cKx * sK[0]));
            x[1] = fma(cu, sD[1], cKx * sK[1]);
            x[2] = gw * sD[0];
        }
    }

    template <int J>
    static __device__ __forceinline__ void ycombine(const double (&sK)[NF], const double (&sD)[NF], double Vc,
                                                    double wxK, double wxA, double cc, double (&y)[NACC]) {
        const double gw = wxA * c_tab<P>.w[J];
        if constexpr (MODE == MODE_K) {
            y[0] = wxK * sK[0];
        } else if constexpr (MODE == MODE_G) {
            y[0] = cc * gw * sD[0];
        } else if constexpr (MODE == MODE_CD) {
            y[0] = fma(cc * Vc * gw, sD[0], wxK * sK[0]);
        } else if constexpr (MODE == MODE_DIV) {
            y[0] = gw * sD[1];
        } else {
            const double cv = cc * Vc * gw;
            y[0] = fma(cv, sD[0], wxK * sK[0]);
            y[1] = fma(gw, sD[2], fma(cv, sD[1], wxK * sK[1]));
            y[2] = gw * sD[1];
        }
    }

    // ---- finish one node: add x- and y-parts, pointwise terms, boundary rows -> out[] (the caller stores) ---------------
    // FAST: the caller knows that the node carries no boundary row, is not the pressure pin and does not lie on an
    // interface line owned by the neighbour rank -- only the element sums and the pointwise terms remain.
    // PW = false: the pointwise terms are known to be absent (their pointers are null).
    template <bool FAST = false, bool PW = true>
    static __device__ __forceinline__ void finish_vals(const MeshDev& g, const MarchArgs& A, int ix, int iy,
                                                       const double (&xp)[NOUT], const double (&yp)[NACC],
                                                       const double (&node)[NF], double wxA_line, double wyA,
                                                       double (&out)[NOUT]) {
        const int off = ix * g.LD + iy;
        const int gix = g.gx0 + ix;
        // interface lines are duplicated on two ranks: the element sums of both ranks are added by the halo exchange,
        // pointwise terms and boundary rows are contributed by the owner only (the rank for which it is NOT the last line)
        const bool owner = FAST || !(g.has_right && ix == g.NX - 1);
        if constexpr (MODE == MODE_K || MODE == MODE_DIV) {
            out[0] = xp[0] + yp[0];
        } else if constexpr (MODE == MODE_G) {
            out[0] = xp[0];
            out[1] = yp[0];
        } else if constexpr (MODE == MODE_CD) {
            double r0 = xp[0] + yp[0];
            if (PW && owner) {
                if (A.e0) r0 = fma(A.d0[off], A.e0[off], r0);
                if (A.e1) r0 = fma(A.d1[off], A.e1[off], r0);
            }
            if constexpr (!FAST) {
                const int side = bc_side(A.bc, gix, iy, g.NXg, g.NY);
                if (side >= 0) r0 = owner ? (node[0] - (A.bc.residual ? A.bc.val0[side] : 0.0)) : 0.0;
            }
            out[0] = r0;
        } else {
            double r0 = xp[0] + yp[0];
            double r1 = xp[1] + yp[1];
            double r2 = xp[2] + yp[2];
            if (PW && A.d0 && owner) {
                r0 = fma(A.d0[off], node[0], fma(A.d1[off], node[1], r0));
                r1 = fma(A.d2[off], node[0], fma(A.d3[off], node[1], r1));
            }
            if (PW && A.e0) r1 = fma(A.cbuoy * (wxA_line * wyA), A.e0[off], r1);  // -(Gr/Re) M T, M is an element sum
            if constexpr (!FAST) {
                const int side = bc_side(A.bc, gix, iy, g.NXg, g.NY);
                if (side >= 0) {
                    r0 = owner ? (node[0] - (A.bc.residual ? A.bc.val0[side] : 0.0)) : 0.0;
                    r1 = owner ? (node[1] - (A.bc.residual ? A.bc.val1[side] : 0.0)) : 0.0;
                }
                if (gix == A.bc.pin_gx && iy == A.bc.pin_iy) r2 = owner ? node[2] : 0.0;
            }
            out[0] = r0;
            out[1] = r1;
            out[2] = r2;
        }
    }

    static __device__ __forceinline__ void finish(const MeshDev& g, const MarchArgs& A, int ix, int iy,
                                                  const double (&xp)[NOUT], const double (&yp)[NACC],
                                                  const double (&node)[NF], double wxA_line, double wyA) {
        double out[NOUT];
        finish_vals(g, A, ix, iy, xp, yp, node, wxA_line, wyA, out);
        const int off = ix * g.LD + iy;
        double* const y[3] = {A.y0, A.y1, A.y2};
#pragma unroll
        for (int o = 0; o < NOUT; ++o)
            if (MODE != MODE_G || y[o]) y[o][off] = out[o];
    }

    template <int I>
    struct RowLoop {
        // rows I .. P-1 of the current element: x-part from registers, y-part from shared memory (or the carry)
        static __device__ __forceinline__ void run(const MeshDev& g, const MarchArgs& A, int m, int iy, int t,
                                                   const double (&r)[NF][n], double cKx, double wyA, double cc,
                                                   const double (&xcarry)[NOUT], const double (&ycarry)[NACC],
                                                   const double* __restrict__ sA, int pitch, int z) {
            if constexpr (I < P) {
                const int ix = m * P + I;
                const int off = ix * g.LD + iy;
                double Uc = 0.0;
                if constexpr (NV) Uc = A.U[off];
                double xp[NOUT], yp[NACC], node[NF];
                xrow<I>(r, Uc, cKx, wyA, cc, xp, z);
                if constexpr (I == 0) {
#pragma unroll
                    for (int o = 0; o < NOUT; ++o) xp[o] += xcarry[o];
#pragma unroll
                    for (int o = 0; o < NACC; ++o) yp[o] = ycarry[o];
                } else {
#pragma unroll
                    for (int o = 0; o < NACC; ++o) yp[o] = sA[(o * P + (I - 1)) * pitch + t];
                }
#pragma unroll
                for (int f = 0; f < NF; ++f) node[f] = r[f][I];
                const double wxl = 0.5 * g.dx * asm_weight<P>(ix, g.nex);
                finish(g, A, ix, iy, xp, yp, node, wxl, wyA);
                RowLoop<I + 1>::run(g, A, m, iy, t, r, cKx, wyA, cc, xcarry, ycarry, sA, pitch, z);
            }
        }
    };

    template <int J>
    struct YLoop {
        // rows J .. P-1 of one element line: write to the accumulators (colour phase 1); row P is returned in top[]
        static __device__ __forceinline__ void run(const double (&l)[NF][n], const double* __restrict__ sV, int col0,
                                                   double wxK, double wxA, double cc, double* __restrict__ sA,
                                                   int accStride, double (&top)[NACC], int z) {
            double Vc = 0.0;
            if constexpr (NV) Vc = sV[col0 + J];
            double y[NACC];
            yrow<J>(l, Vc, wxK, wxA, cc, y, z);
            if constexpr (J < P) {
#pragma unroll
                for (int o = 0; o < NACC; ++o) sA[o * accStride + col0 + J] = y[o];
                YLoop<J + 1>::run(l, sV, col0, wxK, wxA, cc, sA, accStride, top, z);
            } else {
#pragma unroll
                for (int o = 0; o < NACC; ++o) top[o] = y[o];
            }
        }
    };

    // ---- y phase over the staged lines.  FULL: slots 0..P-1 (lines m*P+1 .. m*P+P); else only slot P-1 ---------------
    template <bool FULL>
    static __device__ __forceinline__ void yphase(const MeshDev& g, const MarchArgs& A, int lineP, int nty, int halo,
                                                  bool last_strip, double* __restrict__ sU, double* __restrict__ sA,
                                                  int pitch, double cc, int z) {
        const int q = threadIdx.x;
        const int nlines = FULL ? P : 1;
        const int nfull = nlines * nty;
        int slot, nn;
        if (FULL) { slot = q % P; nn = q / P; } else { slot = P - 1; nn = q; }
        const bool full_item = q < nfull;
        const bool halo_item = (halo > 0) && (q >= nfull) && (q < nfull + nlines);
        if (halo_item) { slot = FULL ? (q - nfull) : (P - 1); nn = -1; }
        double top[NACC];
#pragma unroll
        for (int o = 0; o < NACC; ++o) top[o] = 0.0;
        const int col0 = halo + nn * P;   // column of node j = 0 of element nn (halo element: column 0)
        const int accStride = P * pitch;
        double* sAl = sA + slot * pitch;
        if (full_item || halo_item) {
            const int ix = lineP - (P - 1) + slot;   // local line of this slot (slot P-1 <-> lineP)
            const double wxA = 0.5 * g.dx * asm_weight<P>(ix, g.nex);
            const double wxK = wxA * (2.0 / g.dy);
            double l[NF][n];
#pragma unroll
            for (int f = 0; f < NF; ++f)
#pragma unroll
                for (int k = 0; k < n; ++k) l[f][k] = sU[(f * P + slot) * pitch + col0 + k];
            const double* sV = sU + (NF * P + slot) * pitch;
            if (full_item) {
                YLoop<0>::run(l, sV, col0, wxK, wxA, cc, sAl, accStride, top, z);
            } else {
                double Vc = 0.0;
                if constexpr (NV) Vc = sV[col0 + P];
                yrow<P>(l, Vc, wxK, wxA, cc, top, z);
            }
        }
        __syncthreads();   // colour phase 2: every element adds its top row to the node it shares with the element above
        if (full_item) {
            if (nn == nty - 1) {
                if (last_strip) {
#pragma unroll
                    for (int o = 0; o < NACC; ++o) sAl[o * accStride + col0 + P] = top[o];
                }
            } else {
#pragma unroll
                for (int o = 0; o < NACC; ++o) sAl[o * accStride + col0 + P] += top[o];
            }
        } else if (halo_item) {
#pragma unroll
            for (int o = 0; o < NACC; ++o) sAl[o * accStride + halo] += top[o];
        }
        __syncthreads();
    }
};

template <int P, int MODE>
__global__ void __launch_bounds__(SEM_MARCH_MAXT, ModeTraits<MODE>::MINB) sem_march_kernel(const MeshDev g, const MarchArgs A, const int Ty,
                                                         const int Mx, const int pitch, const int m_lo, const int m_hi) {
    using MM = March<P, MODE>;
    constexpr int n = P + 1;
    constexpr int NF = MM::NF, NV = MM::NV, NACC = MM::NACC, NOUT = MM::NOUT;
    extern __shared__ double smem[];
    double* sU = smem;                                   // [NF+NV][P][pitch] staged node lines
    double* sA = smem + (NF + NV) * P * pitch;   // [NACC][P][pitch]  y-contraction accumulators

    const int n0 = blockIdx.x * Ty;
    const int nty = min(Ty, g.ney - n0);
    const int m0 = m_lo + blockIdx.y * Mx;   // the launch covers the element columns m_lo .. m_hi - 1
    const int m1 = min(m0 + Mx, m_hi);
    const int halo = (n0 > 0) ? P : 0;
    const int ybase = n0 * P - halo;
    const int ncol = halo + nty * P + 1;
    const bool last_strip = (n0 + nty == g.ney);
    const int nown = nty * P + (last_strip ? 1 : 0);
    const int t = threadIdx.x;
    const int iy = ybase + t;
    const bool stage = t < ncol;
    const bool own = (t >= halo) && (t < halo + nown);
    const double cc = A.cconv;

    const double* fld[3] = {A.a, A.b, A.c};
    const double wyA = own ? 0.5 * g.dy * asm_weight<P>(iy, g.ney) : 0.0;
    const double cKx = wyA * (2.0 / g.dx);

    double r[NF][n];
    double xcarry[NOUT], ycarry[NACC];
#pragma unroll
    for (int o = 0; o < NOUT; ++o) xcarry[o] = 0.0;

    const int z0 = opaque_zero(A.zero);
    // ---- prologue: line m0*P.  Its x-part from the element on the left (x-halo), its y-part from a 1-line y phase ----
    {
        const int ix = m0 * P;
        if (stage) {
            const int off = ix * g.LD + iy;
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                r[f][P] = fld[f][off];
                sU[(f * P + (P - 1)) * pitch + t] = r[f][P];
            }
            if constexpr (NV) sU[(NF * P + (P - 1)) * pitch + t] = A.V[off];
        }
        if (own && m0 > 0) {
#pragma unroll
            for (int f = 0; f < NF; ++f)
#pragma unroll
                for (int k = 0; k < P; ++k) r[f][k] = fld[f][(ix - P + k) * g.LD + iy];
            double Uc = 0.0;
            if constexpr (NV) Uc = A.U[ix * g.LD + iy];
            MM::template xrow<P>(r, Uc, cKx, wyA, cc, xcarry, z0);
        }
        __syncthreads();
        MM::template yphase<false>(g, A, ix, nty, halo, last_strip, sU, sA, pitch, cc, z0);
#pragma unroll
        for (int o = 0; o < NACC; ++o) ycarry[o] = own ? sA[(o * P + (P - 1)) * pitch + t] : 0.0;
    }

    // ---- march over the element columns of this chunk ------------------------------------------------------------------
    for (int m = m0; m < m1; ++m) {
        __syncthreads();   // accumulators / staged lines of the previous step are no longer read
        const int z = opaque_zero(A.zero);
        if (stage) {
#pragma unroll
            for (int f = 0; f < NF; ++f) r[f][0] = r[f][P];
#pragma unroll
            for (int k = 1; k <= P; ++k) {
                const int off = (m * P + k) * g.LD + iy;
#pragma unroll
                for (int f = 0; f < NF; ++f) r[f][k] = fld[f][off];
            }
#pragma unroll
            for (int k = 1; k <= P; ++k) {
#pragma unroll
                for (int f = 0; f < NF; ++f) sU[(f * P + (k - 1)) * pitch + t] = r[f][k];
                if constexpr (NV) sU[(NF * P + (k - 1)) * pitch + t] = A.V[(m * P + k) * g.LD + iy];
            }
        }
        __syncthreads();
        MM::template yphase<true>(g, A, m * P + P, nty, halo, last_strip, sU, sA, pitch, cc, z);
        if (own) {
            MM::template RowLoop<0>::run(g, A, m, iy, t, r, cKx, wyA, cc, xcarry, ycarry, sA, pitch, z);
            double Uc = 0.0;
            if constexpr (NV) Uc = A.U[(m * P + P) * g.LD + iy];
            MM::template xrow<P>(r, Uc, cKx, wyA, cc, xcarry, z);
#pragma unroll
            for (int o = 0; o < NACC; ++o) ycarry[o] = sA[(o * P + (P - 1)) * pitch + t];
        }
    }

    // ---- epilogue: the last line of the slab has no element to its right ------------------------------------------------
    if (own && m1 == g.nex) {
        const int ix = g.nex * P;
        double node[NF];
#pragma unroll
        for (int f = 0; f < NF; ++f) node[f] = r[f][P];
        const double wxl = 0.5 * g.dx * asm_weight<P>(ix, g.nex);
        MM::finish(g, A, ix, iy, xcarry, ycarry, node, wxl, wyA);
    }
}

// Host-side launch geometry shared by all instantiations.
struct MarchGeom {
    int Ty, Mx, pitch, threads;
    int m_lo = 0, m_hi = 0;   // element columns covered by the launch
    dim3 grid;
    const XchArgs* xch = nullptr;   // v3 only: in-kernel interface exchange (fused partitioned apply)
};

inline MarchGeom march_geometry(const MeshDev& g, int Ty_req, int Mx_req, int sm_count, int m_lo, int m_hi) {
    MarchGeom q;
    const int P = g.P;
    int Ty = Ty_req > 0 ? Ty_req : (256 / P > 0 ? 256 / P : 1);
    if (Ty > g.ney) Ty = g.ney;
    while (Ty * P + P + 1 > 320 && Ty > 1) --Ty;
    const int strips = (g.ney + Ty - 1) / Ty;
    int Mx = Mx_req;
    if (Mx <= 0) {
        // enough chunks for ~8 CTAs per SM, but chunks of at least 8 columns (x-halo cost 1/Mx) when the mesh allows
        const int want = (8 * sm_count + strips - 1) / strips;
        Mx = (g.nex + want - 1) / want;
        if (Mx < 8) Mx = 8;
        if (Mx > 64) Mx = 64;
    }
    if (Mx > g.nex) Mx = g.nex;
    q.Ty = Ty;
    q.Mx = Mx;
    const int ncol = P + Ty * P + 1;
    q.pitch = ncol + ((17 - ncol % 16) % 16);   // smallest pitch >= ncol with pitch % 16 == 1
    q.threads = round_up(ncol > Ty * P + P ? ncol : Ty * P + P, 32);
    q.m_lo = m_lo;
    q.m_hi = m_hi;
    q.grid = dim3((unsigned)strips, (unsigned)((m_hi - m_lo + Mx - 1) / Mx), 1);
    return q;
}

}  // namespace semb
