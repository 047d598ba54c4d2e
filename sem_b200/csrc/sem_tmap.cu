// Host side of the 2-D TMA staging of the marching kernels: tensor maps (cuTensorMapEncodeTiled, resolved through the
// runtime's driver entry point -- no link dependency on libcuda) of fields [NX][LD] fp64 for a box of P lines x PITCH columns.
#include "sem_common.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <cstring>
#include <mutex>

namespace semb {

namespace {
PFN_cuTensorMapEncodeTiled g_encode = nullptr;

int load_encode() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    SEM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !fn) { set_error("cuTensorMapEncodeTiled is not available in this driver"); return -1; }
    g_encode = (PFN_cuTensorMapEncodeTiled)fn;
    return 0;
}

// Encoding costs about a microsecond; a Krylov iteration reuses the same few vectors, so a small direct-mapped cache
// (keyed by everything that enters the descriptor) takes it off the launch path.
struct Entry {
    const double* base;
    int NX, LD, bc, bl;
    TMap map;
};
constexpr int CACHE = 64;
Entry g_cache[CACHE];
std::mutex g_mu;
}  // namespace

int tmap_get(const double* base, const MeshDev& g, int box_cols, int box_lines, TMap* out) {
    static_assert(sizeof(CUtensorMap) == sizeof(TMap), "CUtensorMap is expected to be 128 bytes");
    if (!base) { set_error("tmap_get: null field"); return -2; }
    if (((uintptr_t)base & 15u) != 0) { set_error("tmap_get: fields must be 16-byte aligned"); return -2; }
    std::lock_guard<std::mutex> lock(g_mu);
    Entry& e = g_cache[((uintptr_t)base >> 8) * 2654435761u % CACHE];
    if (e.base == base && e.NX == g.NX && e.LD == g.LD && e.bc == box_cols && e.bl == box_lines) {
        *out = e.map;
        return 0;
    }
    if (load_encode()) return -1;
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)g.LD, (cuuint64_t)g.NX};          // fastest first: columns (y), lines (x)
    const cuuint64_t strides[1] = {(cuuint64_t)g.LD * sizeof(double)};       // bytes between lines
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_lines};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
        return -1;
    }
    std::memcpy(&e.map, &m, sizeof(m));
    e.base = base; e.NX = g.NX; e.LD = g.LD; e.bc = box_cols; e.bl = box_lines;
    *out = e.map;
    return 0;
}

}  // namespace semb
