// Shared device-side definitions of the sem_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <string>

#define SEM_MAX_P 16

namespace semb {

// ---------------------------------------------------------------------------------------------------------------
// Mesh slab handled by one GPU.  Global vector layout (SEM.py:110): node (ix, iy) at iy + NY*ix, x slow / y fast.
// On the device every field is a padded 2-D array [NX][LD]; pads stay zero.
// ---------------------------------------------------------------------------------------------------------------
struct MeshDev {
    int P;
    int nex, ney;        // LOCAL element columns (x) and element rows (y)
    int NX, NY, LD;      // local node lines, nodes per line, padded pitch (doubles)
    int gx0;             // global line index of local line 0
    int NXg;             // global number of node lines
    int has_left;        // a neighbour rank owns elements left of local line 0   (line 0 is an interface)
    int has_right;       // a neighbour rank owns elements right of the last line  (last line is an interface)
    double dx, dy;
};

// 1-D GLL tables of order P in constant memory: every DFMA of the sum-factorised contractions takes its table
// operand straight from the constant bank (uniform across the warp), leaving shared memory to the nodal data.
// Rows are padded to an even length so that every row starts 16-byte aligned (LDCU.128 fetches two operands).
template <int P>
struct __align__(16) Tab {
    static constexpr int NP = (P + 2) / 2 * 2;   // padded row stride
    double D[(P + 1) * NP];    // D[i][k] = l_k'(xi_i)                      GLL.py:45-59
    double Ks[(P + 1) * NP];   // Ks[i][k] = sum_q w_q D[q][i] D[q][k]       GLL.py:73-81
    double w[NP];              // quadrature weights                         GLL.py:30
};
template <int P>
__constant__ Tab<P> c_tab;

// Boundary rows.  Sides W, E, S, N (later wins at corners, CD:62-71 / NS:78-88).
enum { SIDE_W = 0, SIDE_E = 1, SIDE_S = 2, SIDE_N = 3 };
struct BCSpec {
    int active[4];       // side carries a Dirichlet row
    double val0[4];      // value for output field 0 (T, or u)
    double val1[4];      // value for output field 1 (v)
    int residual;        // 1: row = x - value (residual form), 0: row = x (JVP form)
    int pin_gx, pin_iy;  // pressure pin node (global line, iy); pin_gx < 0: none
};

__device__ __forceinline__ int bc_side(const BCSpec& bc, int gix, int iy, int NXg, int NY) {
    int side = -1;
    if (gix == 0 && bc.active[SIDE_W]) side = SIDE_W;
    if (gix == NXg - 1 && bc.active[SIDE_E]) side = SIDE_E;
    if (iy == 0 && bc.active[SIDE_S]) side = SIDE_S;
    if (iy == NY - 1 && bc.active[SIDE_N]) side = SIDE_N;
    return side;
}

// Arguments of the fused marching operator kernel (see sem_march.cuh).
enum MarchMode {
    MODE_K = 0,    // y0 = K a
    MODE_G = 1,    // y0 = s G_x a ; y1 = s G_y a                      (s = cconv)
    MODE_CD = 2,   // y0 = K a + cconv (U o G_x a + V o G_y a) [+ d0 o e0 + d1 o e1], Dirichlet rows
    MODE_NS = 3,   // 3-field Navier-Stokes residual / JVP
    MODE_DIV = 4   // y0 = G_x a + G_y b                               (continuity rows, used by the preconditioner)
};

struct MarchArgs {
    const double* a;   // contracted fields
    const double* b;
    const double* c;
    const double* U;   // advecting velocity
    const double* V;
    const double* d0;  // pointwise diagonals (NS: Re G_x u, Re G_y u, Re G_x v, Re G_y v ; CD: Pe G_x T, Pe G_y T)
    const double* d1;
    const double* d2;
    const double* d3;
    const double* e0;  // pointwise partners (CD: du, dv ; NS: e0 = T for the buoyancy term)
    const double* e1;
    double* y0;
    double* y1;
    double* y2;
    double cconv;      // Re or Pe (or the scale of MODE_G)
    double cbuoy;      // coefficient of M o e0 in y1 (NS: -Gr/Re), 0 = off
    int zero;          // always 0 (see opaque_zero in sem_march.cuh)
    BCSpec bc;
};

// In-kernel interface exchange of a partitioned apply (sem_march3_kernel; peer-memory mailboxes, see sem_comm.cuh).
// mask == 0: none.  The strips that finish an interface line store their segment of it straight into the neighbour's
// mailbox over NVLink and release a per-strip epoch flag there; at the end of its chunk the same CTA acquires the
// neighbour's flag for the same segment and adds what arrived.  Index [0]: local line 0 (left neighbour), [1]: last line.
struct XchArgs {
    int mask;                          // bit 0: line 0 is exchanged, bit 1: the last line
    int nedge;                         // the first `nedge` values of blockIdx.y are the edge chunks below
    int e_lo[2], e_hi[2];              // element columns [e_lo, e_hi) of the edge chunks
    unsigned long long slot_len;       // doubles between the fields of a mailbox slot
    unsigned long long parity_stride;  // doubles between the two epoch parities
    unsigned long long flag_stride;    // 64-bit words between the fields of a flag array
    double* peer_slot[2];              // parity 0, field 0 of the slot this rank writes in the neighbour's mailbox
    unsigned long long* peer_flag[2];  // the neighbour's arrival flags of that slot [field][strip]
    const double* my_slot[2];          // the slot the neighbour writes in this rank's mailbox
    unsigned long long* my_flag[2];    // this rank's arrival flags [field][strip]
    unsigned long long* epoch[2];      // local epoch counters [field][strip] (device memory: graph-replayable)
};

// 2-D TMA tensor maps of the staged fields of a marching launch (opaque 128-byte CUtensorMap objects, encoded on the host:
// sem_tmap.cu).  Field = [NX][LD] doubles, box = [P lines][PITCH columns]: one cp.async.bulk.tensor.2d per field and step.
struct alignas(64) TMap { unsigned long long opaque[16]; };
struct TmaMaps { TMap m[5]; };
// encode (or fetch from a small cache) the tensor map of a field for a box of box_cols x box_lines doubles
int tmap_get(const double* base, const MeshDev& g, int box_cols, int box_lines, TMap* out);

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace semb

// error plumbing shared by the host files
namespace semb {
void set_error(const std::string& s);
}
#define SEM_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (call);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            semb::set_error(std::string(#call) + " failed: " + cudaGetErrorString(_e) + " at " +    \
                            __FILE__ + ":" + std::to_string(__LINE__));                             \
            return -1;                                                                              \
        }                                                                                           \
    } while (0)
