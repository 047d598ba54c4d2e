// fp64 GEMM on DMMA.8x8x4 + the fast-diagonalisation plans (see sem_gemm.cuh).
#include "sem_gemm.cuh"

#include <algorithm>
#include <cstring>

namespace semb {

// ---------------------------------------------------------------------------------------------------------------
// GEMM kernel.  CTA tile 128 x 128 x 16, 8 warps as 4 (m) x 2 (n), warp tile 32 x 64 = 4 x 8 DMMA fragments (64 fp64
// accumulators per thread), 4-stage cp.async ring (LDGSTS, 16-byte chunks; padded shared-memory pitches -- a dense TMA box
// would put the four k-rows of a B fragment into the same banks).  Per k-step of 4 a warp issues 12 LDS.64 for 32 DMMAs.
//   DMMA m8n8k4 fragments: A[lane/4][lane%4], B[k = lane%4][n = lane/4], C[lane/4][2*(lane%4) + {0,1}].
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d[0]), "+d"(d[1])
                 : "d"(a), "d"(b));
}

// CF: tile configuration -- BM x BN CTA tile, WM x WN warp tile (multiples of 8), BK = 16, STAGES-deep cp.async ring.
//   GemmBig  : 128 x 128, warps 32 x 64 (8 warps, 64 accumulators per thread)  -- large transforms, one CTA per SM
//   GemmMid  :  64 x  64, warps 32 x 32 (4 warps)  -- a few hundred nodes per direction: enough CTAs to fill the SMs
//   GemmSmall:  32 x  32, warps 16 x 16 (4 warps)  -- the reference's own meshes (65 x 65 nodes): a 128-wide tile would pad a
//               32 x 63 x 32 product 32-fold and leave it to ONE SM (24-40 us per launch in the round-2 launch list)
template <int BM_, int BN_, int WM_, int WN_, int STAGES_>
struct GemmCfg {
    static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, STAGES = STAGES_, BK = 16;
    static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN, THREADS = 32 * WARPS_M * WARPS_N;
    static constexpr int FM = WM / 8, FN = WN / 8;
    static constexpr int LDA = BK + 4;      // = 4 (mod 16): conflict-free LDS.64 fragment loads
    static constexpr int LDB = BN + 4;      // = 4 (mod 16) for BN a multiple of 16
    static constexpr int A_CHUNKS = BM * BK / 2, B_CHUNKS = BK * BN / 2;   // 16-byte chunks per stage
    static constexpr size_t SMEM = (size_t)STAGES * (BM * LDA + BK * LDB) * sizeof(double);
    static_assert(A_CHUNKS % THREADS == 0 && B_CHUNKS % THREADS == 0, "tile loads must divide evenly over the threads");
};
using GemmBig = GemmCfg<128, 128, 32, 64, 4>;
using GemmMid = GemmCfg<64, 64, 32, 32, 3>;
using GemmSmall = GemmCfg<32, 32, 16, 16, 3>;

template <class CF, int EPI>
__global__ void __launch_bounds__(CF::THREADS, 1) k_dgemm(const __grid_constant__ GemmArgs g) {
    extern __shared__ __align__(16) double gsm[];
    const int prob = blockIdx.z % g.nprob, bat = blockIdx.z / g.nprob;
    const GemmProblem& q = g.p[prob];
    const int bm = blockIdx.y * CF::BM, bn = blockIdx.x * CF::BN;
    if (bm >= q.M || bn >= q.N) return;
    double* sA = gsm;
    double* sB = gsm + CF::STAGES * CF::BM * CF::LDA;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = (warp / CF::WARPS_N) * CF::WM, wn = (warp % CF::WARPS_N) * CF::WN;
    const double* __restrict__ Ag = q.A + bat * q.sA + (size_t)bm * q.lda;
    const double* __restrict__ Bg = q.B + bat * q.sB + bn;
    const int KT = q.K / CF::BK;

    auto load = [&](int stage, int kt) {
        const double* a = Ag + kt * CF::BK;
        const double* b = Bg + (size_t)kt * CF::BK * q.ldb;
        double* da = sA + stage * CF::BM * CF::LDA;
        double* db = sB + stage * CF::BK * CF::LDB;
#pragma unroll
        for (int i = 0; i < CF::A_CHUNKS / CF::THREADS; ++i) {
            const int c = tid + i * CF::THREADS;
            const int r = c / (CF::BK / 2), h = c % (CF::BK / 2);
            cp_async16(da + r * CF::LDA + h * 2, a + (size_t)r * q.lda + h * 2);
        }
#pragma unroll
        for (int i = 0; i < CF::B_CHUNKS / CF::THREADS; ++i) {
            const int c = tid + i * CF::THREADS;
            const int r = c / (CF::BN / 2), h = c % (CF::BN / 2);
            cp_async16(db + r * CF::LDB + h * 2, b + (size_t)r * q.ldb + h * 2);
        }
    };

    double acc[CF::FM][CF::FN][2];
#pragma unroll
    for (int i = 0; i < CF::FM; ++i)
#pragma unroll
        for (int j = 0; j < CF::FN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < CF::STAGES - 1; ++s) {
        if (s < KT) load(s, s);
        cp_async_commit();
    }
    const int arow = wm + (lane >> 2), acol = lane & 3;
    const int brow = lane & 3, bcol = wn + (lane >> 2);
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<CF::STAGES - 2>();
        __syncthreads();
        {   // refill the stage consumed in the previous trip
            const int nk = kt + CF::STAGES - 1;
            if (nk < KT) load(nk % CF::STAGES, nk);
            cp_async_commit();
        }
        const int st = kt % CF::STAGES;
        const double* a = sA + st * CF::BM * CF::LDA + arow * CF::LDA + acol;
        const double* b = sB + st * CF::BK * CF::LDB + brow * CF::LDB + bcol;
#pragma unroll
        for (int k4 = 0; k4 < CF::BK / 4; ++k4) {
            double af[CF::FM], bf[CF::FN];
#pragma unroll
            for (int i = 0; i < CF::FM; ++i) af[i] = a[i * 8 * CF::LDA + k4 * 4];
#pragma unroll
            for (int j = 0; j < CF::FN; ++j) bf[j] = b[k4 * 4 * CF::LDB + j * 8];
#pragma unroll
            for (int i = 0; i < CF::FM; ++i)
#pragma unroll
                for (int j = 0; j < CF::FN; ++j) dmma(acc[i][j], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    double* __restrict__ Cg = q.C + bat * q.sC;
#pragma unroll
    for (int i = 0; i < CF::FM; ++i) {
        const int row = bm + wm + i * 8 + (lane >> 2);
        double lxr = 0.0;
        if constexpr (EPI == EPI_SCALE) lxr = q.lx[row];
#pragma unroll
        for (int j = 0; j < CF::FN; ++j) {
            const int col = bn + wn + j * 8 + 2 * (lane & 3);
            double v0 = acc[i][j][0], v1 = acc[i][j][1];
            if constexpr (EPI == EPI_SCALE) {
                const double d0 = lxr + q.ly[col], d1 = lxr + q.ly[col + 1];
                v0 = d0 > g.den_floor ? v0 / d0 : 0.0;
                v1 = d1 > g.den_floor ? v1 / d1 : 0.0;
            }
            *reinterpret_cast<double2*>(Cg + (size_t)row * q.ldc + col) = make_double2(v0, v1);
        }
    }
}

template <class CF>
static int gemm_launch_cfg(const GemmArgs& a, int epi, int maxM, int maxN, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        SEM_CUDA(cudaFuncSetAttribute(k_dgemm<CF, EPI_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CF::SMEM));
        SEM_CUDA(cudaFuncSetAttribute(k_dgemm<CF, EPI_SCALE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CF::SMEM));
        configured = true;
    }
    dim3 grid((unsigned)(maxN / CF::BN), (unsigned)(maxM / CF::BM), (unsigned)(a.nprob * a.batch));
    if (epi == EPI_SCALE) k_dgemm<CF, EPI_SCALE><<<grid, CF::THREADS, CF::SMEM, st>>>(a);
    else k_dgemm<CF, EPI_NONE><<<grid, CF::THREADS, CF::SMEM, st>>>(a);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

int gemm_launch(const GemmArgs& a, int epi, cudaStream_t st) {
    int maxM = 0, maxN = 0, tile = a.tile;
    if (tile != 32 && tile != 64 && tile != 128) { set_error("gemm_launch: tile must be 32, 64 or 128"); return -2; }
    for (int i = 0; i < a.nprob; ++i) {
        const GemmProblem& q = a.p[i];
        if (q.M % tile || q.N % tile || q.K % 16 || q.K <= 0 || (q.lda & 1) || (q.ldb & 1) || (q.ldc & 1)) {
            set_error("gemm_launch: dimensions must be padded (M, N to the tile, K to 16, even leading dimensions)");
            return -2;
        }
        maxM = std::max(maxM, q.M);
        maxN = std::max(maxN, q.N);
    }
    if (a.nprob < 1 || maxM == 0 || maxN == 0) return 0;
    if (tile == 128) return gemm_launch_cfg<GemmBig>(a, epi, maxM, maxN, st);
    if (tile == 64) return gemm_launch_cfg<GemmMid>(a, epi, maxM, maxN, st);
    return gemm_launch_cfg<GemmSmall>(a, epi, maxM, maxN, st);
}

// ---------------------------------------------------------------------------------------------------------------
// fold / unfold kernels (streaming).  A direction with active range [lo, lo + cnt) and h = cnt / 2:
//   even part e_i = v[lo+i] + v[lo+cnt-1-i] (i < h), e_h = v[lo+h] when cnt is odd;   odd part o_i = v[lo+i] - v[lo+cnt-1-i].
//   unfolded direction: e_i = v[lo+i], no odd part.
// Buffers are [rows][ld] with the even part at offset 0 and the odd part at offset nep in the folded direction.
// ---------------------------------------------------------------------------------------------------------------
// x fold: src = user vector [NX][LD] (+ field stride), dst rows [0, ne) even / [nep, nep + no) odd, columns j - ylo
__global__ void k_fold_x(const double* __restrict__ r, long long rstride, int LD, FdmDir x, int ylo, int ycnt,
                         double* __restrict__ dst, int ld, long long dstride) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= ycnt) return;
    const double* src = r + blockIdx.z * rstride;
    double* d = dst + blockIdx.z * dstride;
    const double a = src[(size_t)(x.lo + i) * LD + ylo + j];
    if (!x.fold) {
        d[(size_t)i * ld + j] = a;
        return;
    }
    const int h = x.cnt / 2;
    if (i < h) {
        const double b = src[(size_t)(x.lo + x.cnt - 1 - i) * LD + ylo + j];
        d[(size_t)i * ld + j] = a + b;
        d[(size_t)(x.nep + i) * ld + j] = a - b;
    } else {
        d[(size_t)i * ld + j] = a;   // middle line (cnt odd)
    }
}

// y fold: src buffer [rows][lds] columns 0 .. y.cnt-1 -> dst [rows][ldd] even cols [0, ne), odd cols [nep, nep + no)
__global__ void k_fold_y(const double* __restrict__ src, int lds, long long sstride, FdmDir y, double* __restrict__ dst, int ldd,
                         long long dstride, int rows) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (i >= rows || j >= y.ne) return;
    const double* s = src + blockIdx.z * sstride + (size_t)i * lds;
    double* d = dst + blockIdx.z * dstride + (size_t)i * ldd;
    if (!y.fold) {
        d[j] = s[j];
        return;
    }
    const int h = y.cnt / 2;
    const double a = s[j];
    if (j < h) {
        const double b = s[y.cnt - 1 - j];
        d[j] = a + b;
        d[y.nep + j] = a - b;
    } else {
        d[j] = a;
    }
}

__global__ void k_unfold_y(const double* __restrict__ src, int lds, long long sstride, FdmDir y, double* __restrict__ dst, int ldd,
                           long long dstride, int rows) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (i >= rows || j >= y.ne) return;
    const double* s = src + blockIdx.z * sstride + (size_t)i * lds;
    double* d = dst + blockIdx.z * dstride + (size_t)i * ldd;
    if (!y.fold) {
        d[j] = s[j];
        return;
    }
    const int h = y.cnt / 2;
    const double a = s[j];
    if (j < h) {
        const double b = s[y.nep + j];
        d[j] = a + b;
        d[y.cnt - 1 - j] = a - b;
    } else {
        d[j] = a;
    }
}

// x unfold + outside rule: z[x.lo + i][ylo + j] from the even / odd rows of src; every node outside the active range gets
// r (outside = 1) or 0.  One thread per node of the [NX][NY] field.
__global__ void k_unfold_x(const double* __restrict__ src, int ld, long long sstride, FdmDir x, int ylo, int ycnt,
                           const double* r, double* z, long long fstride, int NX, int NY, int LD, int outside) {
    const int iy = blockIdx.x * blockDim.x + threadIdx.x;
    const int ix = blockIdx.y;
    if (iy >= NY) return;
    const size_t off = blockIdx.z * fstride + (size_t)ix * LD + iy;
    const int i = ix - x.lo, j = iy - ylo;
    if (i < 0 || i >= x.cnt || j < 0 || j >= ycnt) {
        z[off] = outside ? r[off] : 0.0;
        return;
    }
    const double* s = src + blockIdx.z * sstride;
    if (!x.fold) {
        z[off] = s[(size_t)i * ld + j];
        return;
    }
    const int h = x.cnt / 2;
    const int m = (i < x.cnt - 1 - i) ? i : x.cnt - 1 - i;   // folded index
    const double a = s[(size_t)m * ld + j];
    if (m >= h) {   // middle line
        z[off] = a;
        return;
    }
    const double b = s[(size_t)(x.nep + m) * ld + j];
    z[off] = (i == m) ? a + b : a - b;
}

// ---------------------------------------------------------------------------------------------------------------
// Padding unit of the transform buffers: 128 (the big GEMM tile) for large directions, 32 (the small tile) below GM_SMALL_MAX.
static int padu(int v, int unit) { return v > 0 ? round_up(v, unit) : 0; }

void fdm_dir_free(FdmDir& d) {
    if (d.Qe) cudaFree(d.Qe);
    if (d.QeT) cudaFree(d.QeT);
    if (d.Qo) cudaFree(d.Qo);
    if (d.QoT) cudaFree(d.QoT);
    if (d.lam) cudaFree(d.lam);
    std::memset(&d, 0, sizeof(d));
}

int fdm_plan_free(FdmPlan& p) {
    fdm_dir_free(p.x);
    fdm_dir_free(p.y);
    if (p.bufA) cudaFree(p.bufA);
    if (p.bufB) cudaFree(p.bufB);
    std::memset(&p, 0, sizeof(p));
    return 0;
}

__global__ void k_pad_square(const double* __restrict__ src, int n, double* __restrict__ dst, double* __restrict__ dstT, int np) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (i >= n || j >= n) return;
    const double v = src[(size_t)i * n + j];
    dst[(size_t)i * np + j] = v;
    dstT[(size_t)j * np + i] = v;
}

int fdm_tile_for(int n) { return n >= 1024 ? 128 : (n >= 200 ? 64 : 32); }

int fdm_dir_build(FdmDir& d, int lo, int cnt, int fold, const double* Qe, const double* Qo, const double* lam, int tile) {
    std::memset(&d, 0, sizeof(d));
    d.lo = lo;
    d.cnt = cnt;
    d.fold = fold;
    d.ne = fold ? (cnt + 1) / 2 : cnt;
    d.no = fold ? cnt / 2 : 0;
    d.nep = padu(d.ne, tile);
    d.nop = padu(d.no, tile);
    auto square = [&](const double* src, int n, int np, double** Q, double** QT) -> int {
        if (n == 0) return 0;
        SEM_CUDA(cudaMalloc(Q, sizeof(double) * (size_t)np * np));
        SEM_CUDA(cudaMalloc(QT, sizeof(double) * (size_t)np * np));
        SEM_CUDA(cudaMemset(*Q, 0, sizeof(double) * (size_t)np * np));
        SEM_CUDA(cudaMemset(*QT, 0, sizeof(double) * (size_t)np * np));
        k_pad_square<<<dim3((unsigned)((n + 127) / 128), (unsigned)n), 128>>>(src, n, *Q, *QT, np);
        SEM_CUDA(cudaGetLastError());
        return 0;
    };
    if (square(Qe, d.ne, d.nep, &d.Qe, &d.QeT)) return -1;
    if (square(Qo, d.no, d.nop, &d.Qo, &d.QoT)) return -1;
    SEM_CUDA(cudaMalloc(&d.lam, sizeof(double) * (d.nep + d.nop)));
    SEM_CUDA(cudaMemset(d.lam, 0, sizeof(double) * (d.nep + d.nop)));
    SEM_CUDA(cudaMemcpy(d.lam, lam, sizeof(double) * d.ne, cudaMemcpyDeviceToDevice));
    if (d.no) SEM_CUDA(cudaMemcpy(d.lam + d.nep, lam + d.ne, sizeof(double) * d.no, cudaMemcpyDeviceToDevice));
    SEM_CUDA(cudaDeviceSynchronize());
    return 0;
}

int fdm_plan_build(FdmPlan& p, const MeshDev& g, int xlo, int xcnt, int xfold, const double* Qxe, const double* Qxo,
                   const double* lamx, int ylo, int ycnt, int yfold, const double* Qye, const double* Qyo, const double* lamy,
                   int outside, int nbuf, double den_floor) {
    fdm_plan_free(p);
    if (xlo < 0 || xcnt < 1 || xlo + xcnt > g.NX || ylo < 0 || ycnt < 1 || ylo + ycnt > g.NY) {
        set_error("fdm_plan_build: active range outside the mesh");
        return -2;
    }
    p.tile = fdm_tile_for(std::min(xcnt, ycnt));
    if (fdm_dir_build(p.x, xlo, xcnt, xfold, Qxe, Qxo, lamx, p.tile)) return -1;
    if (fdm_dir_build(p.y, ylo, ycnt, yfold, Qye, Qyo, lamy, p.tile)) return -1;
    p.outside = outside;
    p.den_floor = den_floor;
    p.rows = p.x.nep + p.x.nop;
    p.cols = std::max(padu(ycnt, p.tile), p.y.nep + p.y.nop);
    p.nbuf = nbuf;
    const size_t bytes = sizeof(double) * (size_t)nbuf * p.rows * p.cols;
    SEM_CUDA(cudaMalloc(&p.bufA, bytes));
    SEM_CUDA(cudaMalloc(&p.bufB, bytes));
    SEM_CUDA(cudaMemset(p.bufA, 0, bytes));
    SEM_CUDA(cudaMemset(p.bufB, 0, bytes));
    SEM_CUDA(cudaDeviceSynchronize());
    p.ready = 1;
    return 0;
}

int fdm_fold_x(const FdmDir& x, int ylo, int ycnt, const double* r, long long rstride, int LD, double* dst, int ld, long long fs,
               int nf, cudaStream_t st) {
    k_fold_x<<<dim3((unsigned)((ycnt + 127) / 128), (unsigned)x.ne, (unsigned)nf), 128, 0, st>>>(r, rstride, LD, x, ylo, ycnt, dst, ld, fs);
    SEM_CUDA(cudaGetLastError());
    return 0;
}
int fdm_unfold_x(const FdmDir& x, int ylo, int ycnt, const double* src, int ld, long long fs, const double* r, double* z,
                 long long stride, const MeshDev& g, int outside, int nf, cudaStream_t st) {
    k_unfold_x<<<dim3((unsigned)((g.NY + 127) / 128), (unsigned)g.NX, (unsigned)nf), 128, 0, st>>>(src, ld, fs, x, ylo, ycnt, r, z, stride,
                                                                                               g.NX, g.NY, g.LD, outside);
    SEM_CUDA(cudaGetLastError());
    return 0;
}
int fdm_fold_y(const FdmDir& y, const double* src, double* dst, int rows, int ld, long long fs, int nf, cudaStream_t st) {
    k_fold_y<<<dim3((unsigned)((y.ne + 127) / 128), (unsigned)rows, (unsigned)nf), 128, 0, st>>>(src, ld, fs, y, dst, ld, fs, rows);
    SEM_CUDA(cudaGetLastError());
    return 0;
}
int fdm_unfold_y(const FdmDir& y, const double* src, double* dst, int rows, int ld, long long fs, int nf, cudaStream_t st) {
    k_unfold_y<<<dim3((unsigned)((y.ne + 127) / 128), (unsigned)rows, (unsigned)nf), 128, 0, st>>>(src, ld, fs, y, dst, ld, fs, rows);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

// one transform step along x (rows): C[part rows][cols] = Q[part] * B[part rows][cols] for the even and the odd part
static int step_x(const FdmPlan& p, bool transposed, const double* src, double* dst, int ncols, int nf, cudaStream_t st) {
    GemmArgs a;
    std::memset(&a, 0, sizeof(a));
    const long long fs = (long long)p.rows * p.cols;
    int np = 0;
    if (p.x.ne) a.p[np++] = GemmProblem{transposed ? p.x.QeT : p.x.Qe, src, dst, p.x.nep, p.cols, p.cols, 0, fs, fs, p.x.nep, ncols, p.x.nep, nullptr, nullptr};
    if (p.x.no)
        a.p[np++] = GemmProblem{transposed ? p.x.QoT : p.x.Qo, src + (size_t)p.x.nep * p.cols, dst + (size_t)p.x.nep * p.cols,
                                p.x.nop, p.cols, p.cols, 0, fs, fs, p.x.nop, ncols, p.x.nop, nullptr, nullptr};
    a.nprob = np;
    a.batch = nf;
    a.tile = p.tile;
    return gemm_launch(a, EPI_NONE, st);
}

int fdm_step_y(const FdmDir& y, bool transposed, const double* src, double* dst, int rows, int ld, long long fs, const double* lx,
               double den_floor, int nf, int tile, cudaStream_t st) {
    GemmArgs a;
    std::memset(&a, 0, sizeof(a));
    int np = 0;
    if (y.ne)
        a.p[np++] = GemmProblem{src, transposed ? y.QeT : y.Qe, dst, ld, y.nep, ld, fs, 0, fs, rows, y.nep, y.nep, lx, lx ? y.lam : nullptr};
    if (y.no)
        a.p[np++] = GemmProblem{src + y.nep, transposed ? y.QoT : y.Qo, dst + y.nep, ld, y.nop, ld, fs, 0, fs, rows,
                                y.nop, y.nop, lx, lx ? y.lam + y.nep : nullptr};
    a.nprob = np;
    a.batch = nf;
    a.tile = tile;
    a.den_floor = den_floor;
    return gemm_launch(a, lx ? EPI_SCALE : EPI_NONE, st);
}

int fdm_plan_apply(FdmPlan& p, const MeshDev& g, const double* r, double* z, int nf, long long stride, cudaStream_t st) {
    if (!p.ready) { set_error("fdm_plan_apply: plan not built"); return -2; }
    if (nf > p.nbuf) { set_error("fdm_plan_apply: more fields than work buffers"); return -2; }
    const long long fs = (long long)p.rows * p.cols;
    const int ycols = padu(p.y.cnt, p.tile);
    if (fdm_fold_x(p.x, p.y.lo, p.y.cnt, r, stride, g.LD, p.bufA, p.cols, fs, nf, st)) return -1;   // r -> A (x-folded node space)
    if (step_x(p, true, p.bufA, p.bufB, ycols, nf, st)) return -1;                                   // B = Qx^T A  (x modes, y nodes)
    if (fdm_fold_y(p.y, p.bufB, p.bufA, p.rows, p.cols, fs, nf, st)) return -1;
    if (fdm_step_y(p.y, false, p.bufA, p.bufB, p.rows, p.cols, fs, p.x.lam, p.den_floor, nf, p.tile, st)) return -1;   // B = (A Qy) / (lx + ly)
    if (fdm_step_y(p.y, true, p.bufB, p.bufA, p.rows, p.cols, fs, nullptr, 0.0, nf, p.tile, st)) return -1;    // A = B Qy^T
    if (fdm_unfold_y(p.y, p.bufA, p.bufB, p.rows, p.cols, fs, nf, st)) return -1;
    if (step_x(p, false, p.bufB, p.bufA, ycols, nf, st)) return -1;                                  // A = Qx B (x-folded node space)
    return fdm_unfold_x(p.x, p.y.lo, p.y.cnt, p.bufA, p.cols, fs, r, z, stride, g, p.outside, nf, st);
}

}  // namespace semb
