// v3 of the fused marching operator (even polynomial orders, all modes): one WARP per strip, TMA-staged inputs,
// even/odd-folded contractions.
//
// Same decomposition as sem_march.cuh (a strip of element rows marches over Mx element columns; x contractions in
// registers along the march, y contractions through shared memory; owner computes; every value read about once and
// written exactly once; no atomics).  What changed, and the ncu finding behind each change (profiles/README.md):
//
//   * a strip is owned by ONE warp (CTA = 32 threads, EW = 64/P element rows for two node columns per lane).  The lane
//     that finishes a node column in the x phase and the lanes that contract the element lines above it in the y phase
//     sit in the same warp, so a step needs two __syncwarp() and no CTA barrier.  v1/v2 and the first v3 lost most of
//     their issue slots to `stall_barrier` with 2-3 CTAs of 3-5 warps per SM.
//   * ALL inputs of a marching step -- the contracted fields and both advecting velocity components -- arrive by TMA bulk
//     copies (cp.async.bulk, one per node line and field, issued by the lanes of the warp, completion on an mbarrier)
//     into a two-stage ring that is refilled as soon as a stage has been consumed (two steps of loads in flight per
//     warp).  v1/v2 fetched part of the inputs with LDG inside the step: top stall long_scoreboard, spilled registers.
//   * even/odd folding: Ks is centro-symmetric and diag(w) D centro-antisymmetric, so with e_k = a_k + a_{P-k},
//     o_k = a_k - a_{P-k} the rows i and P-i of a contraction share their partial sums E_i, O_i:
//         (Ks a)_i = E + O, (Ks a)_{P-i} = E - O;   (wD a)_i = E' + O', (wD a)_{P-i} = O' - E'.
//     81 instead of 162 DFMAs per element line for K and D together, and half as many table operands.
//   * the folded tables are interleaved (Ke, De | Ko, Do) in constant memory with compile-time offsets, so one LDCU.128
//     feeds a K and a D partial sum for both columns (lines) of the lane: one uniform table fetch per four DFMAs.
//   * shared-memory pitches are compile-time constants: every LDS/STS of the step is [register + immediate].
//   * the top row of every element line goes to a small side array instead of a second colour pass over the accumulator
//     tile; the consumer adds it when it finishes the shared node (fixed two-term order, bitwise reproducible).
//   * interior nodes take a branch-free path; boundary rows (O(sqrt N) nodes) are overwritten by a fix-up pass that only
//     the lanes / steps that touch the boundary execute.  Pointwise Jacobian terms are a template flag (PW).
//   * the topmost node column of the mesh (NY = ney*P + 1) belongs to the last strip, which holds the remaining
//     ney mod EW element rows (possibly none): no lane ever handles a 65th column.
#pragma once
#include "sem_march.cuh"
#include "sem_tma.cuh"

#include <cmath>
#include <type_traits>

namespace semb {

// Folded 1-D tables of order P.  With n = P + 1 nodes: HE = ceil(n/2) even-part entries, HO = floor(n/2) odd-part entries;
// rows i = 0 .. HE-1, row i paired with row P-i (an even order has the unpaired middle row H = P/2).  Row layout (doubles):
//   pairs k = 0 .. HE-1 : ( Ke[i][k], De[i][k] )     Ke = (Ks[i][k] + Ks[i][P-k]) / 2   (middle column: Ks[i][H]),  same for De from wD
//   pairs k = 0 .. HO-1 : ( Ko[i][k], Do[i][k] )     Ko = (Ks[i][k] - Ks[i][P-k]) / 2
// with wD[i][k] = w_i D[i][k] (GLL.py:30,45-59,73-81).
template <int P>
struct __align__(16) Tab3 {
    static constexpr int HE = (P + 2) / 2, HO = (P + 1) / 2;
    static constexpr int RS = 2 * (P + 1);
    double T[HE * RS];
};
template <int P>
__constant__ Tab3<P> c_tab3;

// NC node columns per lane in the x phase, NL node lines per lane in the y phase (NL >= NC).  Two per lane (16-byte shared
// and global accesses, one table fetch per four DFMAs) for the one- and two-field modes at even orders up to 10; one per
// lane for NS (three contracted fields), for odd orders (an odd element width breaks the 16-byte alignment of the line
// pairs) and for orders >= 12 (the unrolled two-column code no longer fits the register file).
template <int P, int MODE>
struct March3Traits {
    static constexpr bool WIDE = (MODE != MODE_NS) && (P % 2 == 0) && (P <= 10);
    static constexpr int NC = WIDE ? 2 : 1, NL = WIDE ? 2 : 1;
};

// compile-time geometry of a warp strip
template <int P, int MODE>
struct March3Geom {
    using TR = ModeTraits<MODE>;
    static constexpr int NC = March3Traits<P, MODE>::NC, NL = March3Traits<P, MODE>::NL;
    static constexpr int EW = (32 * NC / P) > 0 ? (32 * NC / P) : 1;     // element rows per strip
    // staged columns: halo (P, or P + 1 when the strip starts at an odd node so that the TMA source stays 16-byte
    // aligned) + own + top node, rounded up to an even count
    static constexpr int NCOLP = (P + (P % 2) + EW * P + 1 + 1) & ~1;
    // Row pitch (doubles).  The x phase reads consecutive columns (any pitch is conflict free).  In the y phase 8
    // consecutive lanes -- G = P/NL line slots of 8/G elements -- read 16-byte units at slot*(PITCH/2) + element*(P/2):
    // PITCH/2 == 1 (mod 8) makes that the lane number for every P; when P is a power of two any ODD PITCH/2 gives 8
    // distinct units, which saves up to 16 % of the tile (P = 8: 74 instead of 82 doubles, NS 42 instead of 50).
    static constexpr bool POW2 = (P & (P - 1)) == 0;
    static constexpr int PITCH = POW2 ? NCOLP + ((NCOLP % 4 == 2) ? 0 : 2) : NCOLP + ((18 - NCOLP % 16) % 16);
    static constexpr int TPW = (EW + 1 + 1) & ~1;                        // top-row entries per line
    static constexpr int NSTG = TR::NF + 2 * TR::NV;                     // staged fields: contracted, then U, then V
    // One 2-D tensor-map copy per field and step (box = P lines x PITCH columns, which lands with exactly the row pitch of the
    // tile) needs every field's stage to start 128-byte aligned: the stage stride of a field FS is P * PITCH rounded up to
    // 16 doubles -- unless that padding costs a resident warp per SM (B200: 228 KB of shared memory per SM, 1 KB reserved per
    // CTA), in which case the order keeps P one-line bulk copies per field (16-byte alignment).
    // SPLIT (NS): the contracted fields are double-buffered, the two advecting components have ONE stage each with their own
    // mbarriers -- V is consumed by the y phase and refilled right after it, U by the x phase and refilled after that, so each
    // refill still has the other phase of the step to arrive.  Two field stages less per warp: 30.7 instead of 36.1 KB at
    // P = 8, 7 instead of 6 resident warps per SM.  Measured at 67 M nodes (NS Jacobian apply, ms before -> after): P = 4 1.138 ->
    // 1.036 (95 % of the roofline), P = 8 1.243 -> 1.221 (81 %), P = 12 2.99 -> 2.76, P = 16 3.98 -> 2.95; P = 10 loses its 2-D
    // tensor maps to the padding rule below and part of its uniform table operands (1.90 -> 2.09) and stays double-buffered.
    // SEM_NO_SPLIT_UV builds the fully double-buffered NS kernel of every order for A/B runs.
#ifdef SEM_NO_SPLIT_UV
    static constexpr bool SPLIT = false;
#else
    static constexpr bool SPLIT = (MODE == MODE_NS) && P != 10;
#endif
    static constexpr int NSTAGES = SPLIT ? 2 * TR::NF + 2 * TR::NV : 2 * NSTG;     // field stages per warp
    static constexpr int FS_NAT = P * PITCH, FS_PAD = (FS_NAT + 15) / 16 * 16;
    static constexpr int smem_doubles(int fs) { return NSTAGES * fs + TR::NACC * P * PITCH + TR::NACC * P * TPW; }
    static constexpr int resident(int fs) {
        const int r = 233472 / (smem_doubles(fs) * 8 + 32 + 1024);
        return r > 32 ? 32 : r;
    }
    static constexpr bool TMA2D = resident(FS_PAD) == resident(FS_NAT);
    static constexpr int FS = TMA2D ? FS_PAD : FS_NAT;
    static constexpr int SMEM_DOUBLES = smem_doubles(FS);
    static constexpr size_t SMEM_BYTES = (size_t)SMEM_DOUBLES * 8 + 32;   // + 4 mbarriers
    static_assert(EW * P / NC <= 32 && (P / NL) * EW <= 32, "a strip must fit one warp in both phases");
};

template <int NC>
struct VecN { double v[NC]; };

template <int NC>
__device__ __forceinline__ VecN<NC> ld_vec(const double* p) {
    VecN<NC> r;
    if constexpr (NC == 2) {
        const double2 t = *reinterpret_cast<const double2*>(p);
        r.v[0] = t.x;
        r.v[1] = t.y;
    } else {
        r.v[0] = p[0];
    }
    return r;
}

template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

template <int P, int MODE, bool PW>
struct March3 {
    using MM = March<P, MODE>;
    using GE = March3Geom<P, MODE>;
    static constexpr int n = P + 1, HE = Tab3<P>::HE, HO = Tab3<P>::HO;
    static constexpr bool MID = (P % 2 == 0);                 // even order: row / column P/2 has no partner
    static __host__ __device__ constexpr bool has_partner(int I) { return !(MID && I == HE - 1); }
    static constexpr int NF = MM::NF, NV = MM::NV, NACC = MM::NACC, NOUT = MM::NOUT;
    static constexpr int NC = GE::NC, NL = GE::NL, EW = GE::EW, PITCH = GE::PITCH, TPW = GE::TPW, NSTG = GE::NSTG, FS = GE::FS;
    static constexpr bool SPLIT = GE::SPLIT;
    static constexpr int G = P / NL;   // lanes per element in the y phase
    static constexpr bool HAS_BC = (MODE == MODE_CD || MODE == MODE_NS);
    static constexpr bool NODE_FAST = (MODE == MODE_NS) && PW;   // the NS JVP multiplies the node values by pointwise diagonals
    static constexpr int NPD = NODE_FAST ? 4 : 1;                // pointwise diagonals prefetched into registers per node

    // ---- folded contraction: rows I and P-I of (Ks a) and (wD a) for NV_ vectors at once --------------------------------
    // sK[0] / sD[0]: row I, sK[1] / sD[1]: row P-I (when row I has a partner).  DIR 0: x phase, 1: y phase (selects the fields needed).
    template <int I, int DIR, int NV_>
    static __device__ __forceinline__ void contract(const double (&e)[NV_][NF][HE], const double (&o)[NV_][NF][HO],
                                                    double (&sK)[2][NV_][NF], double (&sD)[2][NV_][NF], int z) {
        // z: always 0, but uniform and loop-variant (see the kernel) -- keeps the table fetches just-in-time LDCUs
        const double2* __restrict__ t2 = reinterpret_cast<const double2*>(
            __builtin_assume_aligned(c_tab3<P>.T + I * Tab3<P>::RS + 2 * z, 16));
        double EK[NV_][NF], ED[NV_][NF], OK[NV_][NF], OD[NV_][NF];
#pragma unroll
        for (int v = 0; v < NV_; ++v)
#pragma unroll
            for (int f = 0; f < NF; ++f) EK[v][f] = ED[v][f] = OK[v][f] = OD[v][f] = 0.0;
        constexpr bool PAIRED = has_partner(I);   // the middle row of an even order: De and Ko rows are zero
#pragma unroll
        for (int k = 0; k < HE; ++k) {
            const double2 c = t2[k];
#pragma unroll
            for (int v = 0; v < NV_; ++v)
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    const bool nk = DIR == 0 ? MM::x_needs_K(f) : MM::y_needs_K(f);
                    const bool nd = (DIR == 0 ? MM::x_needs_D(f) : MM::y_needs_D(f)) && PAIRED;
                    if (nk) EK[v][f] = (k == 0) ? c.x * e[v][f][0] : fma(c.x, e[v][f][k], EK[v][f]);
                    if (nd) ED[v][f] = (k == 0) ? c.y * e[v][f][0] : fma(c.y, e[v][f][k], ED[v][f]);
                }
        }
#pragma unroll
        for (int k = 0; k < HO; ++k) {
            const double2 c = t2[HE + k];
#pragma unroll
            for (int v = 0; v < NV_; ++v)
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    const bool nk = (DIR == 0 ? MM::x_needs_K(f) : MM::y_needs_K(f)) && PAIRED;
                    const bool nd = DIR == 0 ? MM::x_needs_D(f) : MM::y_needs_D(f);
                    if (nk) OK[v][f] = (k == 0) ? c.x * o[v][f][0] : fma(c.x, o[v][f][k], OK[v][f]);
                    if (nd) OD[v][f] = (k == 0) ? c.y * o[v][f][0] : fma(c.y, o[v][f][k], OD[v][f]);
                }
        }
#pragma unroll
        for (int v = 0; v < NV_; ++v)
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                sK[0][v][f] = EK[v][f] + OK[v][f];
                sK[1][v][f] = EK[v][f] - OK[v][f];
                sD[0][v][f] = ED[v][f] + OD[v][f];
                sD[1][v][f] = OD[v][f] - ED[v][f];
            }
    }

    // fold an element line a[0..P] into its even / odd parts
    static __device__ __forceinline__ void fold(const double (&a)[n], double (&e)[HE], double (&o)[HO]) {
#pragma unroll
        for (int k = 0; k < HO; ++k) {
            e[k] = a[k] + a[P - k];
            o[k] = a[k] - a[P - k];
        }
        if constexpr (MID) e[HO] = a[HO];
    }

    // ---- per-node combination of the contractions (sD already carries the quadrature weight of its own direction) -------
    // x phase: wy = assembled y-weight of the column (dy/2 included), ckx = wy * 2/dx, ccw = cconv * wy
    static __device__ __forceinline__ void xcomb(const double (&sK)[NF], const double (&sD)[NF], double Uc, double ckx,
                                                 double wy, double ccw, double (&x)[NOUT]) {
        if constexpr (MODE == MODE_K) {
            x[0] = ckx * sK[0];
        } else if constexpr (MODE == MODE_G) {
            x[0] = ccw * sD[0];
            x[1] = 0.0;
        } else if constexpr (MODE == MODE_CD) {
            x[0] = fma(ccw * Uc, sD[0], ckx * sK[0]);
        } else if constexpr (MODE == MODE_DIV) {
            x[0] = wy * sD[0];
        } else {
            const double cu = ccw * Uc;
            x[0] = fma(wy, sD[2], fma(cu, sD[0], ckx * sK[0]));
            x[1] = fma(cu, sD[1], ckx * sK[1]);
            x[2] = wy * sD[0];
        }
    }
    // y phase: wx = assembled x-weight of the line (dx/2 included), cky = wx * 2/dy, ccw = cconv * wx
    static __device__ __forceinline__ void ycomb(const double (&sK)[NF], const double (&sD)[NF], double Vc, double cky,
                                                 double wx, double ccw, double (&y)[NACC]) {
        if constexpr (MODE == MODE_K) {
            y[0] = cky * sK[0];
        } else if constexpr (MODE == MODE_G) {
            y[0] = ccw * sD[0];
        } else if constexpr (MODE == MODE_CD) {
            y[0] = fma(ccw * Vc, sD[0], cky * sK[0]);
        } else if constexpr (MODE == MODE_DIV) {
            y[0] = wx * sD[1];
        } else {
            const double cv = ccw * Vc;
            y[0] = fma(cv, sD[0], cky * sK[0]);
            y[1] = fma(wx, sD[2], fma(cv, sD[1], cky * sK[1]));
            y[2] = wx * sD[1];
        }
    }

    // ---- one y-phase item: NL lines (slots sp, sp+G, ..) of the element whose node j = 0 sits in tile column col0 -------
    // FULLROWS: rows 0 .. P-1 -> accumulator tile sA, row P -> sT[line][topslot];  else only row P -> sT[line][topslot].
    // Every lane computes (convergent code keeps the table fetches on the uniform datapath); `active` guards the stores.
    template <bool FULLROWS>
    static __device__ __forceinline__ void yitem(bool active, int sp, int col0, int topslot, const double (&wx)[NL], double cc,
                                                 double ky, const double* __restrict__ sB, const double* __restrict__ sVb,
                                                 double* __restrict__ sA, double* __restrict__ sT, int z) {
        double cky[NL], ccw[NL];
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            cky[l] = wx[l] * ky;
            ccw[l] = cc * wx[l];
        }
        constexpr bool V2 = (P % 2 == 0);   // even order: col0 is even, 16-byte shared accesses
        double e[NL][NF][HE], o[NL][NF][HO];
#pragma unroll
        for (int l = 0; l < NL; ++l)
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const double* row = sB + f * FS + (sp + l * G) * PITCH + col0;
                double a[n];
                if constexpr (V2) {
#pragma unroll
                    for (int k = 0; k < P; k += 2) {
                        const double2 v = *reinterpret_cast<const double2*>(row + k);
                        a[k] = v.x;
                        a[k + 1] = v.y;
                    }
                    a[P] = row[P];
                } else {
#pragma unroll
                    for (int k = 0; k <= P; ++k) a[k] = row[k];
                }
                fold(a, e[l][f], o[l][f]);
            }
        // V line of slot sp (NV modes): in the stage of the buffer, or (SPLIT) in the single V stage sVb
        const double* sV = (SPLIT ? sVb : sB + (NF + 1) * FS) + sp * PITCH + col0;
        if constexpr (FULLROWS) {
            double Y[NL][NACC][n];
            static_for<0, HE>([&](auto Jc) {
                constexpr int J = decltype(Jc)::value;
                constexpr bool PAIRED = has_partner(J);
                double sK[2][NL][NF], sD[2][NL][NF];
                contract<J, 1, NL>(e, o, sK, sD, z);
#pragma unroll
                for (int l = 0; l < NL; ++l) {
                    double Vlo = 0.0, Vhi = 0.0;
                    if constexpr (NV) {
                        Vlo = sV[l * G * PITCH + J];
                        if constexpr (PAIRED) Vhi = sV[l * G * PITCH + P - J];
                    }
                    double y[NACC];
                    ycomb(sK[0][l], sD[0][l], Vlo, cky[l], wx[l], ccw[l], y);
#pragma unroll
                    for (int a = 0; a < NACC; ++a) Y[l][a][J] = y[a];
                    if constexpr (PAIRED) {
                        ycomb(sK[1][l], sD[1][l], Vhi, cky[l], wx[l], ccw[l], y);
#pragma unroll
                        for (int a = 0; a < NACC; ++a) Y[l][a][P - J] = y[a];
                    }
                }
            });
            if (active) {
#pragma unroll
                for (int l = 0; l < NL; ++l)
#pragma unroll
                    for (int a = 0; a < NACC; ++a) {
                        double* dst = sA + (a * P + sp + l * G) * PITCH + col0;
                        if constexpr (V2) {
#pragma unroll
                            for (int j = 0; j < P; j += 2)
                                *reinterpret_cast<double2*>(dst + j) = make_double2(Y[l][a][j], Y[l][a][j + 1]);
                        } else {
#pragma unroll
                            for (int j = 0; j < P; ++j) dst[j] = Y[l][a][j];
                        }
                        sT[(a * P + sp + l * G) * TPW + topslot] = Y[l][a][P];
                    }
            }
        } else {
            double sK[2][NL][NF], sD[2][NL][NF];
            contract<0, 1, NL>(e, o, sK, sD, z);
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                double Vhi = 0.0;
                if constexpr (NV) Vhi = sV[l * G * PITCH + P];
                double y[NACC];
                ycomb(sK[1][l], sD[1][l], Vhi, cky[l], wx[l], ccw[l], y);
                if (active) {
#pragma unroll
                    for (int a = 0; a < NACC; ++a) sT[(a * P + sp + l * G) * TPW + topslot] = y[a];
                }
            }
        }
    }

    // ---- y phase over the P staged lines of one buffer ----------------------------------------------------------------------
    // main pass: lane q < G*nty contracts NL lines of element q / G (slots q % G, ..); halo pass: lanes q < G evaluate the
    // top row of the element below the strip (its node j = 0 is tile column halo - P) -> sT[line][0].  wx: x-weights of the
    // slots (lane % G) + l*G -- the same for both passes because q < G implies q % G == q.
    static __device__ __forceinline__ void yphase(int nty, int halo, const double (&wx)[NL], double cc, double ky,
                                                  const double* __restrict__ sB, const double* __restrict__ sVb,
                                                  double* __restrict__ sA, double* __restrict__ sT, int z) {
        const int q = threadIdx.x;
        const int sp = q % G;
        const int nn = q / G;
        // Divergent on purpose: running both passes with all lanes and guarded stores measured slower (CD 0.48 ms against
        // 0.39 ms at config 5: the convergent code spills uniform registers, R2UR.FILL).
        if (q < G * nty) yitem<true>(true, sp, halo + nn * P, nn + 1, wx, cc, ky, sB, sVb, sA, sT, z);
        if (halo > 0 && q < G) yitem<false>(true, sp, halo - P, 0, wx, cc, ky, sB, sVb, sA, sT, z);
    }

    // store the NOUT outputs of the lane's NC nodes of line ix
    static __device__ __forceinline__ void store(const MeshDev& g, const MarchArgs& A, int ix, int iy0,
                                                 const bool (&own)[NC], const double (&out)[NC][NOUT]) {
        const int off = ix * g.LD + iy0;
        double* const y[3] = {A.y0, A.y1, A.y2};
#pragma unroll
        for (int o = 0; o < NOUT; ++o) {
            if (MODE == MODE_G && !y[o]) continue;
            if constexpr (NC == 2) {
                if (own[1]) {
                    *reinterpret_cast<double2*>(y[o] + off) = make_double2(out[0][o], out[1][o]);   // 16-byte aligned
                } else if (own[0]) {
                    y[o][off] = out[0][o];
                }
            } else {
                if (own[0]) y[o][off] = out[0][o];
            }
        }
    }

    // interior path: element sums + pointwise terms, no boundary logic.  pd: the NS Jacobian diagonals of this node
    // (Re G_x u, Re G_y u, Re G_x v, Re G_y v), fetched into registers at the top of the step (zeros when absent).
    static __device__ __forceinline__ void finalize_fast(const MeshDev& g, const MarchArgs& A, int ix, int iy0,
                                                         const bool (&own)[NC], const double (&xp)[NC][NOUT],
                                                         const double (&yp)[NC][NACC], const double (&node)[NC][NF],
                                                         const double (&wyA)[NC], const double (&pd)[NC][NPD]) {
        double out[NC][NOUT];
        if constexpr (NODE_FAST) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                MM::template finish_vals<true, false>(g, A, ix, iy0 + c, xp[c], yp[c], node[c], 0.0, wyA[c], out[c]);
                out[c][0] = fma(pd[c][0], node[c][0], fma(pd[c][1], node[c][1], out[c][0]));
                out[c][1] = fma(pd[c][2], node[c][0], fma(pd[c][3], node[c][1], out[c][1]));
                if (A.e0) {   // buoyancy -(Gr/Re) M T (residual / coupled JVP only, not in the Krylov loop)
                    const double wxl = 0.5 * g.dx * asm_weight<P>(ix, g.nex);
                    out[c][1] = fma(A.cbuoy * (wxl * wyA[c]), A.e0[ix * g.LD + iy0 + c], out[c][1]);
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < NC; ++c)
                MM::template finish_vals<true, PW>(g, A, ix, iy0 + c, xp[c], yp[c], node[c], 0.0, wyA[c], out[c]);
        }
        store(g, A, ix, iy0, own, out);
    }

    // general path (last line of the slab: Dirichlet E line, interface line of a partition)
    static __device__ __forceinline__ void finalize_slow(const MeshDev& g, const MarchArgs& A, int ix, int iy0,
                                                         const bool (&own)[NC], const double (&xp)[NC][NOUT],
                                                         const double (&yp)[NC][NACC], const double (&node)[NC][NF],
                                                         const double (&wyA)[NC]) {
        double out[NC][NOUT];
        const double wxl = 0.5 * g.dx * asm_weight<P>(ix, g.nex);
#pragma unroll
        for (int c = 0; c < NC; ++c)
            MM::template finish_vals<false, true>(g, A, ix, iy0 + c, xp[c], yp[c], node[c], wxl, wyA[c], out[c]);
        store(g, A, ix, iy0, own, out);
    }

    // Boundary rows of the lines m*P .. m*P+P-1 at the lane's columns: Dirichlet rows (x - value, or x in JVP form) and the
    // pressure pin overwrite what the interior path stored (same thread, same address: program order).  Only lanes on
    // the S/N boundary columns, the step that holds the W line and the step that holds the pin line get here.
    static __device__ __noinline__ void bc_fixup(const MeshDev& g, const MarchArgs& A, int m, int iy0, int c0, bool own0,
                                                 bool own1, const double* __restrict__ sB, const double* aprev) {
        for (int R = 0; R < P; ++R) {
            const int ix = m * P + R, gix = g.gx0 + ix;
            for (int c = 0; c < NC; ++c) {
                if (!(c == 0 ? own0 : own1)) continue;
                const int iy = iy0 + c;
                const int side = bc_side(A.bc, gix, iy, g.NXg, g.NY);
                const bool pin = (MODE == MODE_NS) && gix == A.bc.pin_gx && iy == A.bc.pin_iy;
                if (side < 0 && !pin) continue;
                const int off = ix * g.LD + iy;
                double nv[NF];
                for (int f = 0; f < NF; ++f) nv[f] = (R == 0) ? aprev[c * NF + f] : sB[f * FS + (R - 1) * PITCH + c0 + c];
                if (side >= 0) {
                    A.y0[off] = nv[0] - (A.bc.residual ? A.bc.val0[side] : 0.0);
                    if constexpr (MODE == MODE_NS) A.y1[off] = nv[1] - (A.bc.residual ? A.bc.val1[side] : 0.0);
                }
                if constexpr (MODE == MODE_NS) {
                    if (pin) A.y2[off] = nv[2];
                }
            }
        }
    }

    // ---- x phase of one element column for the NC node columns of the lane -------------------------------------------------------
    // FULL: element m (lines m*P .. m*P+P): finishes and stores lines m*P .. m*P+P-1, leaves the x-part of line (m+1)*P
    //       in xc, its y-part in yc, its node values in a0 and its U in U0.
    // !FULL (prologue of a chunk that does not start at the left edge): only the carries are produced.
    // On entry a0 = node values of the first line of the element; the other P lines are the rows of the stage sB.
    template <bool FULL>
    static __device__ __forceinline__ void xphase(const MeshDev& g, const MarchArgs& A, int m, int iy0, int c0, int topi,
                                                  const bool (&own)[NC], bool colflag, const double* __restrict__ sB,
                                                  const double* __restrict__ sUb, const double* __restrict__ sA,
                                                  const double* __restrict__ sT,
                                                  const double (&wyA)[NC], double cc, double (&a0)[NC][NF],
                                                  double (&U0)[NC], double (&xc)[NC][NOUT], double (&yc)[NC][NACC],
                                                  const double (&pd)[P][NC][NPD], int z) {
        double ckx[NC], ccw[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            ckx[c] = wyA[c] * (2.0 / g.dx);
            ccw[c] = cc * wyA[c];
        }
        double e[NC][NF][HE], o[NC][NF][HO];
        double aprev[NC][NF];
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            double a[NC][n];
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                a[c][0] = a0[c][f];
                aprev[c][f] = a0[c][f];
            }
#pragma unroll
            for (int k = 1; k <= P; ++k) {
                const VecN<NC> v = ld_vec<NC>(sB + f * FS + (k - 1) * PITCH + c0);
#pragma unroll
                for (int c = 0; c < NC; ++c) a[c][k] = v.v[c];
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                a0[c][f] = a[c][P];
                fold(a[c], e[c][f], o[c][f]);
            }
        }
        // y-part / U / node values of line m*P + R (R >= 1) at the lane's columns
        auto load_y = [&](int R, double (&yp)[NC][NACC]) {
#pragma unroll
            for (int a = 0; a < NACC; ++a) {
                const VecN<NC> v = ld_vec<NC>(sA + (a * P + (R - 1)) * PITCH + c0);
#pragma unroll
                for (int c = 0; c < NC; ++c) yp[c][a] = v.v[c];
                if (topi >= 0) yp[0][a] += sT[(a * P + (R - 1)) * TPW + topi];   // only the first column can be a shared node
            }
        };
        auto load_U = [&](int R, double (&Uc)[NC]) {
            if constexpr (NV) {
                const VecN<NC> v = ld_vec<NC>((SPLIT ? sUb : sB + NF * FS) + (R - 1) * PITCH + c0);
#pragma unroll
                for (int c = 0; c < NC; ++c) Uc[c] = v.v[c];
            } else {
#pragma unroll
                for (int c = 0; c < NC; ++c) Uc[c] = 0.0;
            }
        };
        auto load_node = [&](int R, double (&node)[NC][NF]) {
#pragma unroll
            for (int f = 0; f < NF; ++f) {
#pragma unroll
                for (int c = 0; c < NC; ++c) node[c][f] = 0.0;
                if constexpr (NODE_FAST) {
                    const VecN<NC> v = ld_vec<NC>(sB + f * FS + (R - 1) * PITCH + c0);
#pragma unroll
                    for (int c = 0; c < NC; ++c) node[c][f] = v.v[c];
                }
            }
        };

        static_for<0, (FULL ? HE : 1)>([&](auto Ic) {
            constexpr int I = decltype(Ic)::value;
            double sK[2][NC][NF], sD[2][NC][NF];
            contract<I, 0, NC>(e, o, sK, sD, z);
            if constexpr (FULL) {   // row I
                const int ix = m * P + I;
                double Uc[NC], xp[NC][NOUT], yp[NC][NACC], node[NC][NF];
                if constexpr (I == 0) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) Uc[c] = U0[c];
                } else {
                    load_U(I, Uc);
                }
#pragma unroll
                for (int c = 0; c < NC; ++c) xcomb(sK[0][c], sD[0][c], Uc[c], ckx[c], wyA[c], ccw[c], xp[c]);
                if constexpr (I == 0) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
#pragma unroll
                        for (int q = 0; q < NOUT; ++q) xp[c][q] += xc[c][q];
#pragma unroll
                        for (int q = 0; q < NACC; ++q) yp[c][q] = yc[c][q];
#pragma unroll
                        for (int f = 0; f < NF; ++f) node[c][f] = aprev[c][f];
                    }
                } else {
                    load_y(I, yp);
                    load_node(I, node);
                }
                finalize_fast(g, A, ix, iy0, own, xp, yp, node, wyA, pd[I]);
            }
            if constexpr (has_partner(I)) {   // row P - I
                constexpr int R = P - I;
                double Uc[NC], xp[NC][NOUT];
                load_U(R, Uc);
#pragma unroll
                for (int c = 0; c < NC; ++c) xcomb(sK[1][c], sD[1][c], Uc[c], ckx[c], wyA[c], ccw[c], xp[c]);
                if constexpr (I == 0) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
#pragma unroll
                        for (int q = 0; q < NOUT; ++q) xc[c][q] = xp[c][q];
                        U0[c] = Uc[c];
                    }
                } else {
                    const int ix = m * P + R;
                    double yp[NC][NACC], node[NC][NF];
                    load_y(R, yp);
                    load_node(R, node);
                    finalize_fast(g, A, ix, iy0, own, xp, yp, node, wyA, pd[R]);
                }
            }
        });
        load_y(P, yc);
        if constexpr (FULL && HAS_BC) {
            const int gl = g.gx0 + m * P;
            const bool fix = colflag || gl == 0 || (MODE == MODE_NS && (unsigned)(A.bc.pin_gx - gl) < (unsigned)P);
            if (fix) bc_fixup(g, A, m, iy0, c0, own[0], own[NC - 1], sB, &aprev[0][0]);
        }
    }

    // ---- in-kernel interface exchange (partitioned apply, XchArgs) -----------------------------------------------------------
    // Called once, at the end of a chunk that holds an interface line (side 0: local line 0, side 1: the last line).  The strip
    // stores its segment of the line into the neighbour's mailbox (every lane re-reads what it stored itself: finalize_* and
    // bc_fixup write a node from the same lane, program order) and releases the strip's epoch flag there; then it acquires the
    // neighbour's flag for the same segment and adds what arrived (lower rank's partial sum first).  Every push precedes every
    // wait.  Not inlined on purpose: compiled into the kernel body it raised the uniform-register pressure of the marching
    // loop and ptxas demoted the table offset to a vector register -- every table fetch an LDC instead of an LDCU, the
    // slab apply 18 % slower.
    static __device__ __noinline__ void xch_exchange(const MeshDev& g, const MarchArgs& A, const XchArgs& X, bool first, bool last,
                                                     int iy0, bool own0, bool own1) {
        double* const y[3] = {A.y0, A.y1, A.y2};
        const int strip = blockIdx.x;
        const bool own[2] = {own0, own1};
        for (int pass = 0; pass < 2; ++pass) {            // 0: push, 1: wait + add
            for (int side = 0; side < 2; ++side) {
                if (!(side == 0 ? first : last)) continue;
                const size_t ix = side == 0 ? 0 : (size_t)g.nex * P;
                int fo = 0;
                for (int o = 0; o < NOUT; ++o) {
                    if (!y[o]) continue;
                    unsigned long long* ep = X.epoch[side] + fo * X.flag_stride + strip;
                    const unsigned long long e = *ep + 1ull;
                    double* line = y[o] + ix * g.LD + iy0;
                    if (pass == 0) {
                        double* dst = X.peer_slot[side] + (e & 1ull) * X.parity_stride + fo * X.slot_len + iy0;
                        for (int c = 0; c < NC; ++c)
                            if (own[c]) dst[c] = line[c];
                        __threadfence_system();
                        __syncwarp();
                        if (threadIdx.x == 0) st_release_sys(X.peer_flag[side] + fo * X.flag_stride + strip, e);
                    } else {
                        if (threadIdx.x == 0) wait_epoch(X.my_flag[side] + fo * X.flag_stride + strip, e);
                        __syncwarp();
                        const double* recv = X.my_slot[side] + (e & 1ull) * X.parity_stride + fo * X.slot_len + iy0;
                        for (int c = 0; c < NC; ++c)
                            if (own[c]) {
                                const double mine = line[c], other = __ldcg(recv + c);   // written by a peer: read it from L2
                                line[c] = side == 0 ? (other + mine) : (mine + other);
                            }
                        __syncwarp();
                        if (threadIdx.x == 0) *ep = e;
                    }
                    ++fo;
                }
            }
        }
    }
};

// XCH: the fused partitioned apply (in-kernel interface exchange after the march, edge chunks first).  A template flag and
// not a run-time test of X.mask, and nothing of it inside the marching loop: with a test and the push of line 0 compiled into
// the loop of the one-GPU kernel ptxas scheduled the CD apply differently (164 instead of 184 registers) and config 5 ran
// 4 % slower (0.856 against 0.895 of the HBM roofline).
template <int P, int MODE, bool PW, bool XCH = false>
__global__ void __launch_bounds__(32) sem_march3_kernel(const __grid_constant__ MeshDev g, const __grid_constant__ MarchArgs A,
                                                        const __grid_constant__ XchArgs X, const __grid_constant__ TmaMaps TM,
                                                        const int Mx, const int m_lo, const int m_hi) {
    using M3 = March3<P, MODE, PW>;
    using GE = March3Geom<P, MODE>;
    constexpr int NF = M3::NF, NACC = M3::NACC, NOUT = M3::NOUT, NSTG = M3::NSTG, NC = M3::NC, NL = M3::NL, G = M3::G;
    constexpr int EW = GE::EW, PITCH = GE::PITCH, TPW = GE::TPW, FS = GE::FS;
    constexpr bool SPLIT = GE::SPLIT;                     // single U and V stages with their own barriers (NS)
    constexpr int NBUF = SPLIT ? NF : NSTG;               // fields in a double-buffered stage
    constexpr int STAGE = NBUF * FS;                      // doubles per stage (FS: stage stride of a field, >= P * PITCH)
    extern __shared__ __align__(128) double smem3[];
    double* sS = smem3;                                   // [2][NBUF] field stages of FS doubles: [P][PITCH] staged node lines (TMA destination)
    double* sU1 = sS + 2 * STAGE;                         // SPLIT: the U stage, then the V stage
    double* sV1 = sU1 + FS;
    double* sA = sS + GE::NSTAGES * FS;                   // [NACC][P][PITCH]    y-part rows 0..P-1 of every element line
    double* sT = sA + NACC * P * PITCH;                   // [NACC][P][TPW]      y-part row P (node shared with the element above)
    uint64_t* bar = reinterpret_cast<uint64_t*>(sT + NACC * P * TPW);   // [4]: stage 0, stage 1, U, V

    // back-to-back applies: the next launch may place its CTAs as soon as all of ours have started, i.e. into the slots our
    // tail frees; its set-up (barriers, accumulator clear, weights) then overlaps our tail and its first loads start the
    // moment this grid completes instead of a launch latency later
    pdl_launch_dependents();
    const int lane = threadIdx.x;
    const int n0 = blockIdx.x * EW;
    const int nty = max(0, min(EW, g.ney - n0));          // the last strip holds the remaining ney % EW rows (maybe none)
    const bool last_strip = (blockIdx.x == gridDim.x - 1);
    // the launch covers the element columns m_lo .. m_hi - 1 in chunks of Mx; a fused partitioned apply puts the chunks
    // that finish the interface lines first (lowest blockIdx.y = scheduled first), the interior chunks after them
    int m0 = m_lo + blockIdx.y * Mx, m1 = min(m0 + Mx, m_hi);
    if constexpr (XCH) {
        // selects on uniform values with static indices: a branch or an indexed read of the parameter array yields per-thread
        // registers, and the table offset z = m >> 30 derived from them turns every table fetch from LDCU into LDC (ADU
        // pipe; measured: 18 % slower)
        const int by = blockIdx.y;
        const int e0 = by == 0 ? X.e_lo[0] : X.e_lo[1], e1 = by == 0 ? X.e_hi[0] : X.e_hi[1];
        const int i0 = m_lo + (by - X.nedge) * Mx, i1 = min(i0 + Mx, m_hi);
        m0 = by < X.nedge ? e0 : i0;
        m1 = by < X.nedge ? e1 : i1;
    }
    // y-halo: the P nodes below the strip, one more when the strip starts at an odd node (odd orders) so that the TMA
    // source address stays 16-byte aligned
    const int halo = (n0 > 0) ? P + ((n0 * P - P) & 1) : 0;
    const int ybase = n0 * P - halo;                      // even
    const int ncolp = (halo + nty * P + 1 + 1) & ~1;      // staged columns per line (even)
    const int nown = nty * P + (last_strip ? 1 : 0);      // the strip owns tile columns halo .. halo + nown - 1
    const int cr = NC * lane;                             // first own column of this lane (relative)
    const int c0 = halo + cr;                             // ... in the staged tile
    const int iy0 = ybase + c0;
    const bool xthr = cr < nown;
    bool own[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) own[c] = cr + c < nown;
    const int topi = (xthr && cr % P == 0) ? cr / P : -1;
    const double cc = A.cconv;
    const double ky = 2.0 / g.dy;
    const uint32_t line_bytes = (uint32_t)ncolp * 8u;

    // staged field f (a compare-select chain: a pointer array indexed at run time would live in local memory)
    auto field = [&](int f) -> const double* {
        const double* p = A.a;
        if (NF > 1 && f == 1) p = A.b;
        if (NF > 2 && f == 2) p = A.c;
        if (M3::NV && f == NF) p = A.U;
        if (M3::NV && f == NF + 1) p = A.V;
        return p;
    };

    double wyA[NC];
    bool colflag = false;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        wyA[c] = own[c] ? 0.5 * g.dy * asm_weight<P>(iy0 + c, g.ney) : 0.0;
        colflag = colflag || (own[c] && (iy0 + c == 0 || iy0 + c == g.NY - 1 || (MODE == MODE_NS && iy0 + c == A.bc.pin_iy)));
    }
    // x-weights of the y-phase slots of this lane: interior lines of an element column, the same in every step except
    // for the last slot of the last column of the mesh (no element to its right)
    const int sp = lane % G;
    double wx_in[NL], wx_end[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const int slot = sp + l * G;
        wx_in[l] = 0.5 * g.dx * (slot == P - 1 ? c_tab<P>.w[P] + c_tab<P>.w[0] : c_tab<P>.w[slot + 1]);
        wx_end[l] = 0.5 * g.dx * (slot == P - 1 ? c_tab<P>.w[P] : c_tab<P>.w[slot + 1]);
    }

    if (lane == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        if constexpr (SPLIT) {
            mbar_init(&bar[2], 1);
            mbar_init(&bar[3], 1);
        }
        mbar_fence_init();
    }
    // the accumulator column of the topmost mesh node and the halo entry of sT are never written by the y phase
    for (int i = lane; i < NACC * P * (PITCH + TPW); i += 32) sA[i] = 0.0;
    __syncwarp();

    // one elected lane issues the bulk copies of node lines line_first .. line_first + nlines - 1 of every staged field into slots
    // slot_first .. of buffer `buf`.  (One copy per lane does not parallelise: UBLKCP is a uniform-datapath instruction
    // and the compiler serialises the lanes with an ELECT loop, ~10 instructions per copy.)
    // With tensor maps (GE::TMA2D) the same call is ONE copy per field: the box of P lines x PITCH columns whose slot
    // `slot_first` is line `line_first`; lines before line 0 (the one-line call of the prologue) arrive as zeros.
    // copies of the fields [f_lo, f_hi) into the stage that starts at `dst` (field f at dst + (f - f_lo) * FS), on barrier bp
    auto issue_fields = [&](double* dst, uint64_t* bp, int f_lo, int f_hi, int line_first, int nlines, int slot_first) {
        if (elect_one()) {
            if constexpr (GE::TMA2D) {
                mbar_expect_tx(bp, (uint32_t)((f_hi - f_lo) * P * PITCH * 8));
#pragma unroll
                for (int f = 0; f < NSTG; ++f)
                    if (f >= f_lo && f < f_hi) tma_load_2d(dst + (f - f_lo) * FS, &TM.m[f], ybase, line_first - slot_first, bp);
            } else {
                mbar_expect_tx(bp, (uint32_t)((f_hi - f_lo) * nlines) * line_bytes);
#pragma unroll
                for (int f = 0; f < NSTG; ++f)
                    if (f >= f_lo && f < f_hi) {
                        const double* src = field(f) + (size_t)line_first * g.LD + ybase;
                        double* d = dst + (f - f_lo) * FS + slot_first * PITCH;
                        for (int k = 0; k < nlines; ++k) bulk_g2s(d + k * PITCH, src + (size_t)k * g.LD, line_bytes, bp);
                    }
            }
        }
        __syncwarp();
    };
    auto issue = [&](int buf, int line_first, int nlines, int slot_first) {
        issue_fields(sS + buf * STAGE, &bar[buf], 0, NBUF, line_first, nlines, slot_first);
    };
    auto issue_U = [&](int line_first, int nlines, int slot_first) {
        issue_fields(sU1, &bar[2], NF, NF + 1, line_first, nlines, slot_first);
    };
    auto issue_V = [&](int line_first, int nlines, int slot_first) {
        issue_fields(sV1, &bar[3], NF + 1, NF + 2, line_first, nlines, slot_first);
    };
    uint32_t phU = 0, phV = 0;

    double a0[NC][NF], U0[NC], xc[NC][NOUT], yc[NC][NACC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        U0[c] = 0.0;
#pragma unroll
        for (int q = 0; q < NOUT; ++q) xc[c][q] = 0.0;
#pragma unroll
        for (int q = 0; q < NACC; ++q) yc[c][q] = 0.0;
#pragma unroll
        for (int f = 0; f < NF; ++f) a0[c][f] = 0.0;
    }

    // Jacobian diagonals of the lines a step finishes (NS JVP): loaded at the top of the step, consumed after the y phase
    constexpr int NPD = M3::NPD;
    double pd[P][NC][NPD];
#pragma unroll
    for (int R = 0; R < P; ++R)
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int k = 0; k < NPD; ++k) pd[R][c][k] = 0.0;

    pdl_wait();   // everything above touched parameters, constant tables and shared memory only

    // ---- prologue: carries of line m0*P (x-part from the element on the left, y-part from a y phase over its lines) -------
    {
        const int ix = m0 * P;
        if (m0 > 0) issue(1, ix - P + 1, P, 0);   // lines (m0-1)P+1 .. m0*P -> slots 0 .. P-1
        else issue(1, 0, 1, P - 1);               // line 0 -> slot P-1 (the other slots hold stale values, results unused)
        if constexpr (SPLIT) {
            if (m0 > 0) { issue_V(ix - P + 1, P, 0); issue_U(ix - P + 1, P, 0); }
            else { issue_V(0, 1, P - 1); issue_U(0, 1, P - 1); }
        }
        issue(0, ix + 1, P, 0);                   // first marching step
        if (xthr && m0 > 0) {
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const VecN<NC> v = ld_vec<NC>(field(f) + (size_t)(ix - P) * g.LD + iy0);
#pragma unroll
                for (int c = 0; c < NC; ++c) a0[c][f] = v.v[c];
            }
        }
        double wx[NL];
#pragma unroll
        for (int l = 0; l < NL; ++l) wx[l] = 0.5 * g.dx * asm_weight<P>(ix - P + 1 + sp + l * G, g.nex);
        mbar_wait(&bar[1], 0);
        if constexpr (SPLIT) { mbar_wait(&bar[3], phV); phV ^= 1; }
        const double* sB = sS + STAGE;
        M3::yphase(nty, halo, wx, cc, ky, sB, sV1, sA, sT, 0);
        __syncwarp();
        if constexpr (SPLIT) {
            issue_V(ix + 1, P, 0);                // V of the first marching step: the single V stage is free again
            mbar_wait(&bar[2], phU);
            phU ^= 1;
        }
        if (xthr) {
            if (m0 > 0) {
                M3::template xphase<false>(g, A, m0 - 1, iy0, c0, topi, own, colflag, sB, sU1, sA, sT, wyA, cc, a0, U0, xc, yc, pd, 0);
            } else {
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    const VecN<NC> v = ld_vec<NC>(sB + f * FS + (P - 1) * PITCH + c0);
#pragma unroll
                    for (int c = 0; c < NC; ++c) a0[c][f] = v.v[c];
                }
                if constexpr (M3::NV) {
                    const VecN<NC> v = ld_vec<NC>((SPLIT ? sU1 : sB + NF * FS) + (P - 1) * PITCH + c0);
#pragma unroll
                    for (int c = 0; c < NC; ++c) U0[c] = v.v[c];
                }
#pragma unroll
                for (int a = 0; a < NACC; ++a) {
                    const VecN<NC> v = ld_vec<NC>(sA + (a * P + (P - 1)) * PITCH + c0);
#pragma unroll
                    for (int c = 0; c < NC; ++c) yc[c][a] = v.v[c];
                    if (topi >= 0) yc[0][a] += sT[(a * P + (P - 1)) * TPW + topi];
                }
            }
        }
        __syncwarp();   // buffer 1 and the accumulators are free again
        if constexpr (SPLIT) issue_U(ix + 1, P, 0);
        if (m0 + 1 < m1) issue(1, (m0 + 1) * P + 1, P, 0);
    }

    // ---- march --------------------------------------------------------------------------------------------------------------------
    uint32_t ph0 = 0, ph1 = 1;
    // The table offset z (below) must stay on the uniform datapath.  Derived from m it does not always: ptxas keeps m in a
    // vector register in many instantiations (every table fetch then is an LDC on the ADU pipe: P = 10 CD / NS, P = 12 and
    // 16 CD, P = 8 NS residual, and every kernel whose m0 comes out of the selects of the XCH chunk mapping), so z comes
    // from a second counter that starts at a constant.  Measured at 67 M nodes (CD, fraction of the HBM roofline): P = 10
    // 54 -> 66 %, P = 12 38 -> 56 %, P = 16 28 -> 54 %; the one-GPU CD apply of the orders 4 .. 8, which was on LDCUs either
    // way, keeps the derivation from m (P = 8: 89.3 % against 88.3 % with the second counter).
    constexpr bool ZIT = XCH || !(MODE == MODE_CD && !PW && P >= 4 && P <= 8);
    int it = 0;
    for (int m = m0; m < m1; ++m, ++it) {
        const int b = (ZIT ? it : m - m0) & 1;
        if constexpr (M3::NODE_FAST) {
            if (A.d0 && xthr) {
                const double* const dp[4] = {A.d0, A.d1, A.d2, A.d3};
#pragma unroll
                for (int R = 0; R < P; ++R)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const VecN<NC> v = ld_vec<NC>(dp[k] + (size_t)(m * P + R) * g.LD + iy0);
#pragma unroll
                        for (int c = 0; c < NC; ++c) pd[R][c][k] = v.v[c];
                    }
            }
        }
        if (b == 0) { mbar_wait(&bar[0], ph0); ph0 ^= 1; } else { mbar_wait(&bar[1], ph1); ph1 ^= 1; }
        const double* sB = sS + b * STAGE;
        double wx[NL];
#pragma unroll
        for (int l = 0; l < NL; ++l) wx[l] = (m + 1 == g.nex) ? wx_end[l] : wx_in[l];
        // z == 0 (m < 2^30), but the compiler cannot prove it: uniform and different in every trip, so the constant-table
        // fetches stay LDCUs next to their DFMAs instead of being hoisted into ~100 registers (and R2UR'd back)
        const int z = (ZIT ? it : m) >> 30;
        if constexpr (SPLIT) { mbar_wait(&bar[3], phV); phV ^= 1; }
        M3::yphase(nty, halo, wx, cc, ky, sB, sV1, sA, sT, z);
        __syncwarp();
        if constexpr (SPLIT) {
            if (m + 1 < m1) issue_V((m + 1) * P + 1, P, 0);   // the V stage is free: the x phase of this step hides the refill
            mbar_wait(&bar[2], phU);
            phU ^= 1;
        }
        if (xthr) M3::template xphase<true>(g, A, m, iy0, c0, topi, own, colflag, sB, sU1, sA, sT, wyA, cc, a0, U0, xc, yc, pd, z);
        __syncwarp();   // buffer b and the accumulators are free: refill the buffer with the lines of step m + 2
        if constexpr (SPLIT) {
            if (m + 1 < m1) issue_U((m + 1) * P + 1, P, 0);   // the U stage is free: the y phase of the next step hides the refill
        }
        if (m + 2 < m1) issue(b, (m + 2) * P + 1, P, 0);
    }

    // ---- epilogue: the last line of the slab has no element to its right ---------------------------------------------------------
    if (xthr && m1 == g.nex) M3::finalize_slow(g, A, g.nex * P, iy0, own, xc, yc, a0, wyA);

    // ---- fused partitioned apply: exchange of the interface lines --------------------------------------------------------------
    if constexpr (XCH) {
        const bool first = (X.mask & 1) && blockIdx.y == 0, last = (X.mask & 2) && m1 == g.nex;   // edge chunk 0 starts at column 0
        if (first || last) M3::xch_exchange(g, A, X, first, last, iy0, own[0], own[NC - 1]);
    }
}
// host mirror of March3Traits: node columns per lane of a (P, mode)
inline int march3_nc(int P, int mode) { return (mode != MODE_NS && P % 2 == 0 && P <= 10) ? 2 : 1; }

// grid: x = strips of EW element rows (+ the last strip with the remainder and the topmost node column), y = chunks of Mx
// element columns (the x-halo costs 1/Mx extra reads and one extra y phase per chunk).
inline MarchGeom march3_geometry(const MeshDev& g, int mode, int Mx_req, int sm_count, size_t smem_bytes, size_t smem_sm,
                                 int m_lo, int m_hi) {
    MarchGeom q;
    const int nc = march3_nc(g.P, mode);
    const int EW = (32 * nc / g.P) > 0 ? (32 * nc / g.P) : 1;
    const int strips = g.ney / EW + 1;
    int resident = (int)(smem_sm / (smem_bytes + 1024));
    if (resident > 32) resident = 32;
    if (resident < 1) resident = 1;
    int Mx = Mx_req;
    if (Mx <= 0) {
        // A one-warp CTA marches its chunk in sequence: Mx steps plus a priming phase worth c ~ 1.5 steps, and the launch
        // ends with a tail of about kappa * (Mx + c) steps in which the last CTAs finish alone.  With R = strips * ncol /
        // slots steps of work per resident slot, time ~ R (1 + c / Mx) + kappa (Mx + c) has its minimum at
        // Mx = sqrt(R c / kappa); c / kappa = 1.4 fits the measured chunk-length sweeps of all three modes on 120 ... 1024
        // columns (profiles/README.md; e.g. CD: 5 columns on a 128-column slab, 14 on the whole config-5 mesh).  A model
        // that counts whole resident rounds of CTAs predicted those sweeps badly (CTAs do not run in lock step), and a
        // balanced decomposition (one CTA per resident slot, each marching an equal share of the flattened (strip,
        // column) pairs) ran at HALF the speed: concurrently running warps then read line segments scattered over the
        // whole slab instead of neighbouring segments of the same node lines, and DRAM page locality is lost.
        // Tiny meshes (the reference's own) end up with 2-column chunks: the launch is pure latency.
        const double slots = (double)sm_count * resident;
        const double R = (double)strips * (double)(m_hi - m_lo) / slots;
        Mx = (int)(sqrt(1.4 * R) + 0.5);
        if (Mx < 2) Mx = 2;
        if (Mx > 32) Mx = 32;
    }
    if (Mx > m_hi - m_lo) Mx = m_hi - m_lo;
    q.Ty = EW;
    q.Mx = Mx;
    q.pitch = 0;
    q.threads = 32;
    q.m_lo = m_lo;
    q.m_hi = m_hi;
    q.grid = dim3((unsigned)strips, (unsigned)((m_hi - m_lo + Mx - 1) / Mx), 1);
    return q;
}

}  // namespace semb
