// One translation unit per polynomial order: compiled with -DSEM_P=<P> (see Makefile), so the 16 orders build in
// parallel.  Exposes a launcher and a table uploader for that order through the registry in sem_dispatch.h.
#include "sem_march.cuh"
#include "sem_march3.cuh"
#include <cmath>
#include <cstring>
#include <cstdlib>
#include "sem_dispatch.h"

#ifndef SEM_P
#error "compile with -DSEM_P=<polynomial order>"
#endif

#define SEM_CAT2(a, b) a##b
#define SEM_CAT(a, b) SEM_CAT2(a, b)

namespace semb {

// The round-1 v1 kernel (one column per thread, CTA barriers, LDG staging) is an independent second implementation kept for
// A/B runs only: it is compiled in with `make V1=1` (-DSEM_WITH_V1) and selected with SEM_B200_MARCH=1; the default library
// carries the v3 kernels only.
#ifdef SEM_WITH_V1
template <int P, int MODE>
static int launch_mode(const MeshDev& g, const MarchArgs& A, const MarchGeom& q, cudaStream_t st) {
    const size_t smem = march_smem_doubles<P, MODE>(q.pitch) * sizeof(double);
    static size_t configured = 0;   // per (P, MODE) instantiation
    if (smem > configured) {
        SEM_CUDA(cudaFuncSetAttribute(sem_march_kernel<P, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
        configured = smem;
    }
    sem_march_kernel<P, MODE><<<q.grid, q.threads, smem, st>>>(g, A, q.Ty, q.Mx, q.pitch, q.m_lo, q.m_hi);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

int SEM_CAT(march_launch_p, SEM_P)(int mode, const MeshDev& g, const MarchArgs& A, const MarchGeom& q,
                                   cudaStream_t st) {
    switch (mode) {
        case MODE_K: return launch_mode<SEM_P, MODE_K>(g, A, q, st);
        case MODE_G: return launch_mode<SEM_P, MODE_G>(g, A, q, st);
        case MODE_CD: return launch_mode<SEM_P, MODE_CD>(g, A, q, st);
        case MODE_NS: return launch_mode<SEM_P, MODE_NS>(g, A, q, st);
        case MODE_DIV: return launch_mode<SEM_P, MODE_DIV>(g, A, q, st);
    }
    set_error("march_launch: unknown mode");
    return -2;
}
#else
int SEM_CAT(march_launch_p, SEM_P)(int, const MeshDev&, const MarchArgs&, const MarchGeom&, cudaStream_t) {
    set_error("the v1 marching kernel is not part of this build (make V1=1)");
    return -2;
}
#endif

// ---- v3 kernel (one warp per strip, TMA-staged, folded tables): every order, all modes.
template <int P, int MODE, bool PW, bool XCH>
static int launch_mode3x(const MeshDev& g, const MarchArgs& A, const MarchGeom& q, cudaStream_t st) {
    constexpr size_t smem = March3Geom<P, MODE>::SMEM_BYTES;
    static bool configured = false;
    if (!configured) {
        if (smem > 48 * 1024)
            SEM_CUDA(cudaFuncSetAttribute(sem_march3_kernel<P, MODE, PW, XCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    XchArgs X;
    if (XCH) X = *q.xch;
    else std::memset(&X, 0, sizeof(X));
    TmaMaps TM;
    if constexpr (March3Geom<P, MODE>::TMA2D) {
        using TR = ModeTraits<MODE>;
        const double* f[5] = {A.a, TR::NF > 1 ? A.b : nullptr, TR::NF > 2 ? A.c : nullptr, nullptr, nullptr};
        if (TR::NV) { f[TR::NF] = A.U; f[TR::NF + 1] = A.V; }
        for (int k = 0; k < March3Geom<P, MODE>::NSTG; ++k)
            if (tmap_get(f[k], g, March3Geom<P, MODE>::PITCH, P, &TM.m[k])) return -1;
    }
    // Programmatic dependent launch (see pdl_wait in the kernel).  SEM_B200_PDL=0: plain launches; default: PDL except while the
    // stream is being captured into a CUDA graph; 2: also there (programmatic graph edges).
    static const int pdl = [] { const char* e = std::getenv("SEM_B200_PDL"); return e ? std::atoi(e) : 1; }();
    bool use_pdl = pdl != 0;
    if (pdl == 1) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); cs = cudaStreamCaptureStatusActive; }
        use_pdl = cs == cudaStreamCaptureStatusNone;
    }
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = q.grid;
    cfg.blockDim = dim3(32, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl ? 1 : 0;
    SEM_CUDA(cudaLaunchKernelEx(&cfg, sem_march3_kernel<P, MODE, PW, XCH>, g, A, X, TM, q.Mx, q.m_lo, q.m_hi));
    return 0;
}

// the fused partitioned apply (q.xch) exists for the modes without a post-operator: K, G, DIV, CD
template <int P, int MODE, bool PW>
static int launch_mode3(const MeshDev& g, const MarchArgs& A, const MarchGeom& q, cudaStream_t st) {
    if constexpr (MODE != MODE_NS) {
        if (q.xch) return launch_mode3x<P, MODE, PW, true>(g, A, q, st);
    } else {
        if (q.xch) { set_error("march3: no in-kernel exchange in the NS mode"); return -2; }
    }
    return launch_mode3x<P, MODE, PW, false>(g, A, q, st);
}

int SEM_CAT(march3_launch_p, SEM_P)(int mode, const MeshDev& g, const MarchArgs& A, const MarchGeom& q,
                                    cudaStream_t st) {
    switch (mode) {
        case MODE_K: return launch_mode3<SEM_P, MODE_K, false>(g, A, q, st);
        case MODE_G: return launch_mode3<SEM_P, MODE_G, false>(g, A, q, st);
        case MODE_DIV: return launch_mode3<SEM_P, MODE_DIV, false>(g, A, q, st);
        case MODE_CD:
            return (A.e0 || A.e1) ? launch_mode3<SEM_P, MODE_CD, true>(g, A, q, st)
                                  : launch_mode3<SEM_P, MODE_CD, false>(g, A, q, st);
        case MODE_NS:
            return (A.d0 || A.e0) ? launch_mode3<SEM_P, MODE_NS, true>(g, A, q, st)
                                  : launch_mode3<SEM_P, MODE_NS, false>(g, A, q, st);
    }
    return 1;
}

// shared memory per CTA (= warp) of the v3 kernel of a mode, 0 when there is none
size_t SEM_CAT(march3_smem_p, SEM_P)(int mode) {
    switch (mode) {
        case MODE_K: return March3Geom<SEM_P, MODE_K>::SMEM_BYTES;
        case MODE_G: return March3Geom<SEM_P, MODE_G>::SMEM_BYTES;
        case MODE_DIV: return March3Geom<SEM_P, MODE_DIV>::SMEM_BYTES;
        case MODE_CD: return March3Geom<SEM_P, MODE_CD>::SMEM_BYTES;
        case MODE_NS: return March3Geom<SEM_P, MODE_NS>::SMEM_BYTES;
    }
    return 0;
}

// Folded tables of the v3 kernel (layout: Tab3 in sem_march3.cuh).  The folding relies on Ks being centro-symmetric and
// diag(w) D centro-antisymmetric; the tables handed in are checked for it (they are, for GLL nodes: GLL.py:22-59).
int SEM_CAT(upload_tab3_p, SEM_P)(const double* D, const double* Ks, const double* w) {
    constexpr int P = SEM_P, n = P + 1, RS = Tab3<P>::RS, HE = Tab3<P>::HE, HO = Tab3<P>::HO;
    double scaleK = 0.0, scaleD = 0.0, asym = 0.0;
    for (int i = 0; i < n * n; ++i) {
        scaleK = std::fmax(scaleK, std::fabs(Ks[i]));
        scaleD = std::fmax(scaleD, std::fabs(D[i]));
    }
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < n; ++k) {
            asym = std::fmax(asym, std::fabs(Ks[i * n + k] - Ks[(P - i) * n + (P - k)]) / scaleK);
            asym = std::fmax(asym, std::fabs(w[i] * D[i * n + k] + w[P - i] * D[(P - i) * n + (P - k)]) / scaleD);
        }
    if (!(asym < 1e-12)) {
        set_error("upload_tab3: the GLL tables are not centro-(anti)symmetric to 1e-12; the folded kernel cannot use them");
        return -2;
    }
    Tab3<P> h;
    for (int i = 0; i < HE * RS; ++i) h.T[i] = 0.0;
    for (int i = 0; i < HE; ++i) {
        double* row = h.T + i * RS;
        for (int k = 0; k < HE; ++k) {
            const bool mid = (P % 2 == 0) && (k == HE - 1);   // the unpaired middle column of an even order
            const double ka = Ks[i * n + k], kb = Ks[i * n + (P - k)];
            const double da = w[i] * D[i * n + k], db = w[i] * D[i * n + (P - k)];
            row[2 * k] = mid ? ka : 0.5 * (ka + kb);
            row[2 * k + 1] = mid ? da : 0.5 * (da + db);
            if (k < HO) {
                row[2 * (HE + k)] = 0.5 * (ka - kb);
                row[2 * (HE + k) + 1] = 0.5 * (da - db);
            }
        }
    }
    SEM_CUDA(cudaMemcpyToSymbol(c_tab3<P>, &h, sizeof(h)));
    return 0;
}

size_t SEM_CAT(march_smem_p, SEM_P)(int mode, int pitch) {
    switch (mode) {
        case MODE_K: return march_smem_doubles<SEM_P, MODE_K>(pitch) * sizeof(double);
        case MODE_G: return march_smem_doubles<SEM_P, MODE_G>(pitch) * sizeof(double);
        case MODE_CD: return march_smem_doubles<SEM_P, MODE_CD>(pitch) * sizeof(double);
        case MODE_DIV: return march_smem_doubles<SEM_P, MODE_DIV>(pitch) * sizeof(double);
        default: return march_smem_doubles<SEM_P, MODE_NS>(pitch) * sizeof(double);
    }
}

// D, Ks: (P+1)^2 row-major host arrays, w: P+1.  Rows are re-packed to the padded stride of Tab<P>.
int SEM_CAT(upload_tab_p, SEM_P)(const double* D, const double* Ks, const double* w) {
    constexpr int P = SEM_P;
    Tab<P> h;
    for (int i = 0; i < (P + 1) * Tab<P>::NP; ++i) h.D[i] = h.Ks[i] = 0.0;
    for (int i = 0; i < Tab<P>::NP; ++i) h.w[i] = 0.0;
    for (int i = 0; i <= P; ++i) {
        for (int k = 0; k <= P; ++k) {
            h.D[i * Tab<P>::NP + k] = D[i * (P + 1) + k];
            h.Ks[i * Tab<P>::NP + k] = Ks[i * (P + 1) + k];
        }
        h.w[i] = w[i];
    }
    SEM_CUDA(cudaMemcpyToSymbol(c_tab<P>, &h, sizeof(h)));
    return 0;
}

}  // namespace semb
