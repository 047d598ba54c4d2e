// Registry of the per-order translation units (sem_march_inst.cu compiled once per P).
#pragma once
#include "sem_common.cuh"

namespace semb {
struct MarchGeom;

#define SEM_FOR_EACH_P(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)

#define SEM_DECL_P(P)                                                                                         \
    int march_launch_p##P(int mode, const MeshDev& g, const MarchArgs& A, const MarchGeom& q, cudaStream_t st); \
    size_t march_smem_p##P(int mode, int pitch);                                                               \
    int march3_launch_p##P(int mode, const MeshDev& g, const MarchArgs& A, const MarchGeom& q, cudaStream_t st); \
    size_t march3_smem_p##P(int mode);                                                                         \
    int upload_tab3_p##P(const double* D, const double* Ks, const double* w);                                  \
    int upload_tab_p##P(const double* D, const double* Ks, const double* w);
SEM_FOR_EACH_P(SEM_DECL_P)
#undef SEM_DECL_P

}  // namespace semb
