// One translation unit per polynomial order: compiled with -DSEM_P=<P> (see Makefile), so the 16 orders build in
// parallel.  Exposes a launcher and a table uploader for that order through the registry in sem_dispatch.h.
#include "sem_march.cuh"
#include "sem_march2.cuh"
#include "sem_dispatch.h"

#ifndef SEM_P
#error "compile with -DSEM_P=<polynomial order>"
#endif

namespace semb {

template <int P, int MODE>
static int launch_mode(const MeshDev& g, const MarchArgs& A, const MarchGeom& q, cudaStream_t st) {
    const size_t smem = march_smem_doubles<P, MODE>(q.pitch) * sizeof(double);
    static size_t configured = 0;   // per (P, MODE) instantiation
    if (smem > configured) {
        SEM_CUDA(cudaFuncSetAttribute(sem_march_kernel<P, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
        configured = smem;
    }
    sem_march_kernel<P, MODE><<<q.grid, q.threads, smem, st>>>(g, A, q.Ty, q.Mx, q.pitch);
    SEM_CUDA(cudaGetLastError());
    return 0;
}

#define SEM_CAT2(a, b) a##b
#define SEM_CAT(a, b) SEM_CAT2(a, b)

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

#if SEM_P % 2 == 0
template <int P, int MODE>
static int launch_mode2(const MeshDev& g, const MarchArgs& A, const MarchGeom& q, cudaStream_t st) {
    const size_t smem = march2_smem_bytes<P, MODE>(q.pitch);
    static size_t configured = 0;
    if (smem > configured) {
        SEM_CUDA(cudaFuncSetAttribute(sem_march2_kernel<P, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
        configured = smem;
    }
    sem_march2_kernel<P, MODE><<<q.grid, q.threads, smem, st>>>(g, A, q.Ty, q.Mx, q.pitch);
    SEM_CUDA(cudaGetLastError());
    return 0;
}
#endif

// v2 kernel (TMA-staged, two columns per thread): even orders, modes K / G / CD / DIV.  Returns 1 when not available.
int SEM_CAT(march2_launch_p, SEM_P)(int mode, const MeshDev& g, const MarchArgs& A, const MarchGeom& q,
                                    cudaStream_t st) {
#if SEM_P % 2 == 0
    switch (mode) {
        case MODE_K: return launch_mode2<SEM_P, MODE_K>(g, A, q, st);
        case MODE_G: return launch_mode2<SEM_P, MODE_G>(g, A, q, st);
        case MODE_CD: return launch_mode2<SEM_P, MODE_CD>(g, A, q, st);
        case MODE_DIV: return launch_mode2<SEM_P, MODE_DIV>(g, A, q, st);
    }
#endif
    return 1;
}

size_t SEM_CAT(march2_smem_p, SEM_P)(int mode, int pitch) {
#if SEM_P % 2 == 0
    switch (mode) {
        case MODE_K: return march2_smem_bytes<SEM_P, MODE_K>(pitch);
        case MODE_G: return march2_smem_bytes<SEM_P, MODE_G>(pitch);
        case MODE_CD: return march2_smem_bytes<SEM_P, MODE_CD>(pitch);
        case MODE_DIV: return march2_smem_bytes<SEM_P, MODE_DIV>(pitch);
    }
#endif
    return 0;   // 0 = this (order, mode) has no v2 kernel
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
