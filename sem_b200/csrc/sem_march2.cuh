// v2 of the fused marching operator (even polynomial orders, one- and two-field modes K / G / CD / DIV).
//
// Same decomposition and arithmetic as sem_march.cuh (the per-row formulas xrow / yrow / finish_vals are shared), but the
// data movement is rebuilt around what the round-1 ncu capture of v1 showed (profiles/README.md): v1 was latency bound
// (top stall long_scoreboard, 27 % warps active) and spent as many issue slots on LDCU constant fetches, address
// arithmetic and 8-byte shared/global accesses as on DFMAs (ADU pipe 48 % at 35 % of the HBM roofline).
//
//   * TMA staging: the P node lines of the next marching step are fetched with cp.async.bulk (one 2 KB bulk copy per line
//     and field, issued by one thread, completion on an mbarrier) into the other half of a double buffer while the current
//     step computes -- global-load latency is off the critical path and costs no registers or LSU issue slots.
//   * two node columns per thread (x phase) and two node lines per thread (y phase): every table operand fetched from the
//     constant bank feeds two DFMAs, every shared/global access moves 16 bytes (LDS.128 / STS.128 / STG.128).
//   * the advecting velocity at the thread's own nodes is loaded straight into registers at the top of the step and
//     consumed after the y phase.
//   * shared memory rows are 16-byte aligned (TMA) with pitch == 2 (mod 16) doubles and the y-phase threads are mapped
//     (line pair, element) so that 8 consecutive lanes hit 8 distinct 16-byte bank groups: conflict-free LDS.128.
#pragma once
#include "sem_march.cuh"
#include "sem_tma.cuh"

#ifndef SEM_MARCH2_MAXT
#define SEM_MARCH2_MAXT 160
#endif

namespace semb {

template <int MODE> struct March2Traits { static constexpr int MINB = 3; };
template <> struct March2Traits<MODE_K> { static constexpr int MINB = 4; };

template <int P, int MODE>
struct March2 {
    using MM = March<P, MODE>;
    static constexpr int n = P + 1, H = P / 2;
    static constexpr int NF = MM::NF, NV = MM::NV, NACC = MM::NACC, NOUT = MM::NOUT;
    static constexpr int NS_ = NF + NV;   // TMA-staged fields per buffer: contracted fields, then V

    // rows I .. P-1 of the current element for the thread's two columns
    template <int I>
    struct RowLoop2 {
        static __device__ __forceinline__ void run(const MeshDev& g, const MarchArgs& A, int m, int iy0, int c0,
                                                   bool own0, bool own1, const double (&r)[2][NF][n],
                                                   const double2 (&Ur)[n], const double (&cKx)[2],
                                                   const double (&wyA)[2], double cc, const double (&xcarry)[2][NOUT],
                                                   const double (&ycarry)[2][NACC], const double* __restrict__ sA,
                                                   int pitch, int z) {
            if constexpr (I < P) {
                const int ix = m * P + I;
                double xp[2][NOUT], yp[2][NACC], node[2][NF], out[2][NOUT];
                MM::template xrow<I>(r[0], Ur[I].x, cKx[0], wyA[0], cc, xp[0], z);
                MM::template xrow<I>(r[1], Ur[I].y, cKx[1], wyA[1], cc, xp[1], z);
                if constexpr (I == 0) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
#pragma unroll
                        for (int o = 0; o < NOUT; ++o) xp[c][o] += xcarry[c][o];
#pragma unroll
                        for (int o = 0; o < NACC; ++o) yp[c][o] = ycarry[c][o];
                    }
                } else {
#pragma unroll
                    for (int o = 0; o < NACC; ++o) {
                        const double2 a = *reinterpret_cast<const double2*>(sA + (o * P + (I - 1)) * pitch + c0);
                        yp[0][o] = a.x;
                        yp[1][o] = a.y;
                    }
                }
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int f = 0; f < NF; ++f) node[c][f] = r[c][f][I];
                const double wxl = 0.5 * g.dx * asm_weight<P>(ix, g.nex);
                MM::finish_vals(g, A, ix, iy0, xp[0], yp[0], node[0], wxl, wyA[0], out[0]);
                MM::finish_vals(g, A, ix, iy0 + 1, xp[1], yp[1], node[1], wxl, wyA[1], out[1]);
                store_rows(g, A, ix, iy0, own0, own1, out);
                RowLoop2<I + 1>::run(g, A, m, iy0, c0, own0, own1, r, Ur, cKx, wyA, cc, xcarry, ycarry, sA, pitch, z);
            }
        }
    };

    static __device__ __forceinline__ void store_rows(const MeshDev& g, const MarchArgs& A, int ix, int iy0, bool own0,
                                                      bool own1, const double (&out)[2][NOUT]) {
        const int off = ix * g.LD + iy0;
        double* const y[3] = {A.y0, A.y1, A.y2};
#pragma unroll
        for (int o = 0; o < NOUT; ++o) {
            if (MODE == MODE_G && !y[o]) continue;
            if (own0 && own1) {
                *reinterpret_cast<double2*>(y[o] + off) = make_double2(out[0][o], out[1][o]);   // 16-byte aligned
            } else if (own0) {
                y[o][off] = out[0][o];
            } else if (own1) {
                y[o][off + 1] = out[1][o];
            }
        }
    }

    // y phase over the P staged lines of one buffer, two lines (slots sp and sp + P/2) per thread
    static __device__ __forceinline__ void yphase2(const MeshDev& g, int line1, int nty, int halo, bool last_strip,
                                                   const double* __restrict__ sU, double* __restrict__ sA, int pitch,
                                                   double cc, int z) {
        const int q = threadIdx.x;
        const int nfull = H * nty;
        int sp, nn;
        sp = q % H;
        nn = q / H;
        const bool full_item = q < nfull;
        const bool halo_item = (halo > 0) && (q >= nfull) && (q < nfull + H);
        if (halo_item) { sp = q - nfull; nn = -1; }
        const int col0 = halo + nn * P;   // even
        const int accStride = P * pitch;
        double top[2][NACC];
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int o = 0; o < NACC; ++o) top[c][o] = 0.0;
        if (full_item || halo_item) {
            double wxA[2], wxK[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                wxA[c] = 0.5 * g.dx * asm_weight<P>(line1 + sp + c * H, g.nex);
                wxK[c] = wxA[c] * (2.0 / g.dy);
            }
            double l[2][NF][n];
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    const double* row = sU + (f * P + sp + c * H) * pitch + col0;
#pragma unroll
                    for (int k = 0; k < P; k += 2) {
                        const double2 v = *reinterpret_cast<const double2*>(row + k);
                        l[c][f][k] = v.x;
                        l[c][f][k + 1] = v.y;
                    }
                    l[c][f][P] = row[P];
                }
            const double* sV0 = sU + (NF * P + sp) * pitch + col0;
            const double* sV1 = sV0 + H * pitch;
            if (full_item) {
                YPairs<0>::run(l, sV0, sV1, wxK, wxA, cc, sA + sp * pitch + col0, H * pitch, accStride, z);
            }
            double Vc0 = 0.0, Vc1 = 0.0;
            if constexpr (NV) { Vc0 = sV0[P]; Vc1 = sV1[P]; }
            MM::template yrow<P>(l[0], Vc0, wxK[0], wxA[0], cc, top[0], z);
            MM::template yrow<P>(l[1], Vc1, wxK[1], wxA[1], cc, top[1], z);
        }
        __syncthreads();   // colour phase 2: top rows are added to the node shared with the element above
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            double* dst = sA + (sp + c * H) * pitch;
            if (full_item) {
                if (nn == nty - 1) {
                    if (last_strip) {
#pragma unroll
                        for (int o = 0; o < NACC; ++o) dst[o * accStride + col0 + P] = top[c][o];
                    }
                } else {
#pragma unroll
                    for (int o = 0; o < NACC; ++o) dst[o * accStride + col0 + P] += top[c][o];
                }
            } else if (halo_item) {
#pragma unroll
                for (int o = 0; o < NACC; ++o) dst[o * accStride + halo] += top[c][o];
            }
        }
        __syncthreads();
    }

    // rows J, J+1 of both lines of a pair: compute and store as 16-byte pairs (colour phase 1)
    template <int J>
    struct YPairs {
        static __device__ __forceinline__ void run(const double (&l)[2][NF][n], const double* __restrict__ sV0,
                                                   const double* __restrict__ sV1, const double (&wxK)[2],
                                                   const double (&wxA)[2], double cc, double* __restrict__ acc,
                                                   int lineStride, int accStride, int z) {
            if constexpr (J < P) {
                double2 v0 = make_double2(0.0, 0.0), v1 = v0;
                if constexpr (NV) {
                    v0 = *reinterpret_cast<const double2*>(sV0 + J);
                    v1 = *reinterpret_cast<const double2*>(sV1 + J);
                }
                double ya[2][NACC], yb[2][NACC];
                MM::template yrow<J>(l[0], v0.x, wxK[0], wxA[0], cc, ya[0], z);
                MM::template yrow<J>(l[1], v1.x, wxK[1], wxA[1], cc, ya[1], z);
                MM::template yrow<J + 1>(l[0], v0.y, wxK[0], wxA[0], cc, yb[0], z);
                MM::template yrow<J + 1>(l[1], v1.y, wxK[1], wxA[1], cc, yb[1], z);
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int o = 0; o < NACC; ++o)
                        *reinterpret_cast<double2*>(acc + c * lineStride + o * accStride + J) =
                            make_double2(ya[c][o], yb[c][o]);
                YPairs<J + 2>::run(l, sV0, sV1, wxK, wxA, cc, acc, lineStride, accStride, z);
            }
        }
    };
};

template <int P, int MODE>
__global__ void __launch_bounds__(SEM_MARCH2_MAXT, March2Traits<MODE>::MINB)
    sem_march2_kernel(const MeshDev g, const MarchArgs A, const int Ty, const int Mx, const int pitch) {
    static_assert(P % 2 == 0, "v2 needs an even polynomial order (16-byte aligned strips)");
    using M2 = March2<P, MODE>;
    using MM = March<P, MODE>;
    constexpr int n = P + 1;
    constexpr int NF = MM::NF, NV = MM::NV, NACC = MM::NACC, NOUT = MM::NOUT, NS_ = M2::NS_;
    extern __shared__ __align__(16) double smem2[];
    double* sU = smem2;                                  // [2][NS_][P][pitch]
    double* sA = smem2 + 2 * NS_ * P * pitch;            // [NACC][P][pitch]
    uint64_t* bar = reinterpret_cast<uint64_t*>(sA + NACC * P * pitch);   // [2]

    const int n0 = blockIdx.x * Ty;
    const int nty = min(Ty, g.ney - n0);
    const int m0 = blockIdx.y * Mx;
    const int m1 = min(m0 + Mx, g.nex);
    const int halo = (n0 > 0) ? P : 0;
    const int ybase = n0 * P - halo;                     // even
    const int ncol = halo + nty * P + 1;
    const int ncolp = (ncol + 1) & ~1;
    const bool last_strip = (n0 + nty == g.ney);
    const int nown = nty * P + (last_strip ? 1 : 0);
    const int t = threadIdx.x;
    const int c0 = 2 * t;
    const int iy0 = ybase + c0;
    const bool own0 = (c0 >= halo) && (c0 < halo + nown);
    const bool own1 = (c0 + 1 >= halo) && (c0 + 1 < halo + nown);
    const bool xthr = own0 || own1;
    const double cc = A.cconv;
    const uint32_t line_bytes = (uint32_t)ncolp * 8u;

    const double* fld[4] = {A.a, A.b, A.c, nullptr};
    if constexpr (NV) fld[NF] = A.V;

    double wyA[2], cKx[2];
    wyA[0] = own0 ? 0.5 * g.dy * asm_weight<P>(iy0, g.ney) : 0.0;
    wyA[1] = own1 ? 0.5 * g.dy * asm_weight<P>(iy0 + 1, g.ney) : 0.0;
    cKx[0] = wyA[0] * (2.0 / g.dx);
    cKx[1] = wyA[1] * (2.0 / g.dx);

    if (t == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // one thread issues the bulk copies of `nlines` node lines of every staged field into buffer `buf`
    auto issue = [&](int buf, int line_first, int nlines, int slot_first) {
        mbar_expect_tx(&bar[buf], (uint32_t)(NS_ * nlines) * line_bytes);
#pragma unroll
        for (int f = 0; f < NS_; ++f)
            for (int k = 0; k < nlines; ++k)
                bulk_g2s(sU + ((buf * NS_ + f) * P + slot_first + k) * pitch,
                         fld[f] + (size_t)(line_first + k) * g.LD + ybase, line_bytes, &bar[buf]);
    };

    double r[2][NF][n];
    double xcarry[2][NOUT], ycarry[2][NACC];
    double2 U0 = make_double2(0.0, 0.0);
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int o = 0; o < NOUT; ++o) xcarry[c][o] = 0.0;

    // ---- prologue: line m0*P (x-part from the element on the left, y-part from a one-line y phase) ---------------------
    const int z0 = opaque_zero(A.zero);
    {
        const int ix = m0 * P;
        if (t == 0) {
            if (m0 > 0) issue(1, ix - P + 1, P, 0);   // lines (m0-1)P+1 .. m0*P -> slots 0 .. P-1
            else issue(1, 0, 1, P - 1);               // line 0 -> slot P-1
            issue(0, ix + 1, P, 0);                   // first marching step
        }
        if (xthr) {
            if (m0 > 0) {
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    const double2 v = *reinterpret_cast<const double2*>(fld[f] + (size_t)(ix - P) * g.LD + iy0);
                    r[0][f][0] = v.x;
                    r[1][f][0] = v.y;
                }
            }
            if constexpr (NV) U0 = *reinterpret_cast<const double2*>(A.U + (size_t)ix * g.LD + iy0);
        }
        mbar_wait(&bar[1], 0);
        const double* sB = sU + 1 * NS_ * P * pitch;
        if (xthr && m0 > 0) {
#pragma unroll
            for (int f = 0; f < NF; ++f)
#pragma unroll
                for (int k = 1; k <= P; ++k) {
                    const double2 v = *reinterpret_cast<const double2*>(sB + (f * P + (k - 1)) * pitch + c0);
                    r[0][f][k] = v.x;
                    r[1][f][k] = v.y;
                }
            MM::template xrow<P>(r[0], U0.x, cKx[0], wyA[0], cc, xcarry[0], z0);
            MM::template xrow<P>(r[1], U0.y, cKx[1], wyA[1], cc, xcarry[1], z0);
        }
        MM::template yphase<false>(g, A, ix, nty, halo, last_strip, const_cast<double*>(sB), sA, pitch, cc, z0);
        if (xthr) {
#pragma unroll
            for (int o = 0; o < NACC; ++o) {
                const double2 a = *reinterpret_cast<const double2*>(sA + (o * P + (P - 1)) * pitch + c0);
                ycarry[0][o] = a.x;
                ycarry[1][o] = a.y;
            }
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const double2 v = *reinterpret_cast<const double2*>(sB + (f * P + (P - 1)) * pitch + c0);
                r[0][f][P] = v.x;
                r[1][f][P] = v.y;
            }
        }
        __syncthreads();   // buffer 1 and the accumulators are free again
    }

    // ---- march ----------------------------------------------------------------------------------------------------------------
    uint32_t ph0 = 0, ph1 = 1;
    for (int m = m0; m < m1; ++m) {
        const int b = (m - m0) & 1;
        const int z = opaque_zero(A.zero);
        if (t == 0 && m + 1 < m1) issue(1 - b, (m + 1) * P + 1, P, 0);   // prefetch the next step
        double2 Ur[n];
        Ur[0] = U0;
        if constexpr (NV) {
            if (xthr) {
#pragma unroll
                for (int i = 1; i <= P; ++i)
                    Ur[i] = *reinterpret_cast<const double2*>(A.U + (size_t)(m * P + i) * g.LD + iy0);
            }
        }
        if (b == 0) { mbar_wait(&bar[0], ph0); ph0 ^= 1; } else { mbar_wait(&bar[1], ph1); ph1 ^= 1; }
        const double* sB = sU + b * NS_ * P * pitch;
        M2::yphase2(g, m * P + 1, nty, halo, last_strip, sB, sA, pitch, cc, z);
        if (xthr) {
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                r[0][f][0] = r[0][f][P];
                r[1][f][0] = r[1][f][P];
#pragma unroll
                for (int k = 1; k <= P; ++k) {
                    const double2 v = *reinterpret_cast<const double2*>(sB + (f * P + (k - 1)) * pitch + c0);
                    r[0][f][k] = v.x;
                    r[1][f][k] = v.y;
                }
            }
            M2::template RowLoop2<0>::run(g, A, m, iy0, c0, own0, own1, r, Ur, cKx, wyA, cc, xcarry, ycarry, sA, pitch, z);
            MM::template xrow<P>(r[0], Ur[P].x, cKx[0], wyA[0], cc, xcarry[0], z);
            MM::template xrow<P>(r[1], Ur[P].y, cKx[1], wyA[1], cc, xcarry[1], z);
#pragma unroll
            for (int o = 0; o < NACC; ++o) {
                const double2 a = *reinterpret_cast<const double2*>(sA + (o * P + (P - 1)) * pitch + c0);
                ycarry[0][o] = a.x;
                ycarry[1][o] = a.y;
            }
            U0 = Ur[P];
        }
        __syncthreads();   // buffer b and the accumulators are free for the next step
    }

    // ---- epilogue: the last line of the slab has no element to its right ------------------------------------------------------
    if (xthr && m1 == g.nex) {
        const int ix = g.nex * P;
        double node[2][NF], out[2][NOUT];
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int f = 0; f < NF; ++f) node[c][f] = r[c][f][P];
        const double wxl = 0.5 * g.dx * asm_weight<P>(ix, g.nex);
        MM::finish_vals(g, A, ix, iy0, xcarry[0], ycarry[0], node[0], wxl, wyA[0], out[0]);
        MM::finish_vals(g, A, ix, iy0 + 1, xcarry[1], ycarry[1], node[1], wxl, wyA[1], out[1]);
        M2::store_rows(g, A, ix, iy0, own0, own1, out);
    }
}

template <int P, int MODE>
__host__ __device__ constexpr size_t march2_smem_bytes(int pitch) {
    return (size_t)(2 * (ModeTraits<MODE>::NF + ModeTraits<MODE>::NV) + ModeTraits<MODE>::NACC) * P * pitch * 8 + 16;
}

inline MarchGeom march2_geometry(const MeshDev& g, int Ty_req, int Mx_req, int sm_count) {
    MarchGeom q;
    const int P = g.P, H = P / 2;
    int Ty = Ty_req > 0 ? Ty_req : (256 / P > 0 ? 256 / P : 1);
    if (Ty > g.ney) Ty = g.ney;
    auto threads_for = [&](int ty) {
        const int ncolp = (P + ty * P + 1 + 1) & ~1;
        const int a = ncolp / 2, b = H * (ty + 1);
        return round_up(a > b ? a : b, 32);
    };
    while (threads_for(Ty) > SEM_MARCH2_MAXT && Ty > 1) --Ty;
    const int strips = (g.ney + Ty - 1) / Ty;
    int Mx = Mx_req;
    if (Mx <= 0) {
        const int want = (8 * sm_count + strips - 1) / strips;
        Mx = (g.nex + want - 1) / want;
        if (Mx < 8) Mx = 8;
        if (Mx > 64) Mx = 64;
    }
    if (Mx > g.nex) Mx = g.nex;
    q.Ty = Ty;
    q.Mx = Mx;
    const int ncolp = (P + Ty * P + 1 + 1) & ~1;
    q.pitch = ncolp + ((18 - ncolp % 16) % 16);   // smallest even pitch >= ncolp with pitch % 16 == 2
    q.threads = threads_for(Ty);
    q.grid = dim3((unsigned)strips, (unsigned)((g.nex + Mx - 1) / Mx), 1);
    return q;
}

}  // namespace semb
