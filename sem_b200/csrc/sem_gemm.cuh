// Hand-written fp64 GEMM (DMMA) and the fast-diagonalisation plans built on it.  Replaces the cuBLAS DGEMMs of round 1.
//
// fp64 has no tcgen05 path: the only fp64 tensor instruction of sm_100a is DMMA.8x8x4 (mma.sync m8n8k4; the wider PTX
// shapes m16n8k4/k8/k16 compile to sequences of it -- cuobjdump of tools/fp64_peak.cu), measured 37.0 TFLOP/s from
// registers against 33.6 TFLOP/s for DFMA and 34.5 / 35.4 TFLOP/s for cublasDgemm at 4097^3 / 8193^3
// (profiles/r2_fp64_peak.json).  The win over the library is structural: on the reference's meshes every 1-D pencil is
// symmetric about the domain centre, its eigenvectors are even or odd, and each transform splits into an even and an odd
// half-size product -- half the flops of the four full GEMMs.
#pragma once
#include "sem_common.cuh"

namespace semb {

// C[M][N] = A[M][K] B[K][N], all row-major.  M, N multiples of the CTA tile (128 / 64 / 32), K a multiple of 16 (operands are zero padded internal
// buffers), leading dimensions even, pointers 16-byte aligned.  Up to GM_MAXP independent problems per launch (the even and
// the odd half of a folded transform), each repeated `batch` times with the given strides (fields).
constexpr int GM_MAXP = 2;
struct GemmProblem {
    const double* A; const double* B; double* C;
    int lda, ldb, ldc;
    long long sA, sB, sC;    // batch strides (0: shared operand)
    int M, N, K;
    // EPI_SCALE: C[m][n] = acc / (lx[m] + ly[n]) (0 where the denominator is <= den_floor: pseudo-inverse / padding)
    const double* lx; const double* ly;
};
struct GemmArgs {
    GemmProblem p[GM_MAXP];
    int nprob, batch;
    int tile;                // CTA tile of the kernel variant: 128, 64 or 32 (M and N of every problem are multiples of it)
    double den_floor;        // EPI_SCALE: denominators <= den_floor give 0
};
enum { EPI_NONE = 0, EPI_SCALE = 1 };
int gemm_launch(const GemmArgs& a, int epi, cudaStream_t st);

// ---------------------------------------------------------------------------------------------------------------
// Fast-diagonalisation plan:  z = (Qx (x) Qy) diag(1 / (lx_k + ly_l))^+ (Qx (x) Qy)^T r  on the node range
// [x.lo, x.lo + x.cnt) x [y.lo, y.lo + y.cnt) of a padded [NX][LD] vector; outside that range z = r (outside = 1, the identity
// rows of Dirichlet nodes) or z = 0 (outside = 0).  Q^T M1 Q = I, Q^T K1 Q = diag(l) for the 1-D pencils (set-up on the
// host side, sem_b200/device.py).  A direction whose pencil is symmetric about the centre is FOLDED: Qe / Qo hold the
// even / odd eigenvectors on the first ceil(cnt/2) / floor(cnt/2) nodes.
// ---------------------------------------------------------------------------------------------------------------
struct FdmDir {
    int lo, cnt;            // active node range
    int fold;               // parity-split direction
    int ne, no;             // even / odd part sizes (fold: ceil(cnt/2), floor(cnt/2); else cnt, 0) == number of even / odd modes
    int nep, nop;           // padded to multiples of the plan's GEMM tile (0 stays 0)
    double *Qe, *QeT;       // [nep][nep] row-major: Qe[node][mode], QeT[mode][node], zero padded
    double *Qo, *QoT;       // [nop][nop]
    double* lam;            // [nep + nop]: even modes, then odd modes; padding 0
};
struct FdmPlan {
    FdmDir x, y;
    int outside;            // 1: z = r outside the active range, 0: z = 0
    int tile;               // GEMM tile / padding unit: 128 for large meshes, 64 / 32 for small ones (fdm_tile_for)
    int rows, cols;         // padded buffer shape: rows = x.nep + x.nop, cols = max(round_up(y.cnt, 128), y.nep + y.nop)
    int nbuf;               // fields the work buffers hold
    double *bufA, *bufB;    // [nbuf][rows][cols] ping-pong work buffers (zero initialised; pads stay zero)
    double den_floor;       // lx + ly <= den_floor: null mode of an all-Neumann pencil (pseudo-inverse)
    int ready;
};
int fdm_plan_free(FdmPlan& p);
int fdm_tile_for(int n);
int fdm_dir_build(FdmDir& d, int lo, int cnt, int fold, const double* Qe, const double* Qo, const double* lam, int tile);
void fdm_dir_free(FdmDir& d);
// building blocks, also used by the distributed (partitioned) variant in sem_capi.cu: buffers are [rows][ld], `fs` doubles
// between the nf fields
int fdm_fold_x(const FdmDir& x, int ylo, int ycnt, const double* r, long long rstride, int LD, double* dst, int ld, long long fs,
               int nf, cudaStream_t st);
int fdm_unfold_x(const FdmDir& x, int ylo, int ycnt, const double* src, int ld, long long fs, const double* r, double* z,
                 long long stride, const MeshDev& g, int outside, int nf, cudaStream_t st);
int fdm_fold_y(const FdmDir& y, const double* src, double* dst, int rows, int ld, long long fs, int nf, cudaStream_t st);
int fdm_unfold_y(const FdmDir& y, const double* src, double* dst, int rows, int ld, long long fs, int nf, cudaStream_t st);
// dst[rows][part] = src[rows][part] Q(part) (transposed: Q^T); lx != null: spectral scaling 1 / (lx[row] + ly[col])
int fdm_step_y(const FdmDir& y, bool transposed, const double* src, double* dst, int rows, int ld, long long fs, const double* lx,
               double den_floor, int nf, int tile, cudaStream_t st);
// Qe/Qo/lam are DEVICE arrays in the compact (unpadded) shapes [ne][ne], [no][no], [ne + no] per direction.
int fdm_plan_build(FdmPlan& p, const MeshDev& g, int xlo, int xcnt, int xfold, const double* Qxe, const double* Qxo,
                   const double* lamx, int ylo, int ycnt, int yfold, const double* Qye, const double* Qyo, const double* lamy,
                   int outside, int nbuf, double den_floor);
// nf fields `stride` doubles apart.  r and z may alias.
int fdm_plan_apply(FdmPlan& p, const MeshDev& g, const double* r, double* z, int nf, long long stride, cudaStream_t st);

}  // namespace semb
