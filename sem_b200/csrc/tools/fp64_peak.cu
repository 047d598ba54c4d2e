// FP64 peak microbenchmark for B200 (sm_100a): what the FP64 FMA pipe and the FP64 tensor path (DMMA, mma.sync ... f64)
// sustain from registers, and what the library DGEMM reaches at the fast-diagonalisation sizes.  The numbers are the
// denominators of every "FP64 pipe" statement in DESIGN.md / profiles (measure before claiming which roof binds).
//   build: make -C sem_b200/csrc tools        run: sem_b200/csrc/build/fp64_peak [--cublas]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef WITH_CUBLAS
#include <cublas_v2.h>
#endif

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int CH>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = threadIdx.x * 1e-9 + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) acc[c] = fma(acc[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += acc[c];
    if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&d)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int SHAPE, int CH>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a0, double b0) {
    double s = 0;
    if constexpr (SHAPE == 0) {
        double d[CH][2];
#pragma unroll
        for (int c = 0; c < CH; ++c) d[c][0] = d[c][1] = c;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int c = 0; c < CH; ++c) dmma884(d[c], a0, b0);
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) s += d[c][0] + d[c][1];
    } else {
        double d[CH][4];
        double a[8], b[4];
#pragma unroll
        for (int q = 0; q < 8; ++q) a[q] = a0 + q;
#pragma unroll
        for (int q = 0; q < 4; ++q) b[q] = b0 + q;
#pragma unroll
        for (int c = 0; c < CH; ++c) d[c][0] = d[c][1] = d[c][2] = d[c][3] = c;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                if constexpr (SHAPE == 1) { double aa[2] = {a[0], a[1]}; dmma1684(d[c], aa, b[0]); }
                if constexpr (SHAPE == 2) { double aa[4] = {a[0], a[1], a[2], a[3]}; double bb[2] = {b[0], b[1]}; dmma1688(d[c], aa, bb); }
                if constexpr (SHAPE == 3) dmma16816(d[c], a, b);
            }
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
    }
    if (s == 123.456) out[0] = s;
}

template <class F>
static double time_ms(F&& launch, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    int sm = 0;
    CK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0));
    double* out; CK(cudaMalloc(&out, 8));
    const int iters = 20000;
    printf("{\"sm_count\": %d", sm);
    for (int wpb : {4, 8}) {
        const int blocks = sm * (wpb == 4 ? 4 : 2) * 2, threads = 256;
        (void)wpb;
        {
            const double ms = time_ms([&] { k_dfma<16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
            printf(",\n \"dfma_tflops_b%d\": %.2f", blocks, 2.0 * 16 * iters * (double)blocks * threads / ms * 1e-9);
        }
    }
    const int blocks = sm * 8, threads = 256, warps = threads / 32;
    auto rep = [&](const char* name, double flop_per_mma, int ch, double ms) {
        printf(",\n \"%s\": %.2f", name, flop_per_mma * ch * iters * (double)blocks * warps / ms * 1e-9);
    };
    rep("dmma_m8n8k4_ch4_tflops", 2.0 * 8 * 8 * 4, 4, time_ms([&] { k_dmma<0, 4><<<blocks, threads>>>(out, iters, 1.0, 1e-9); }));
    rep("dmma_m8n8k4_ch8_tflops", 2.0 * 8 * 8 * 4, 8, time_ms([&] { k_dmma<0, 8><<<blocks, threads>>>(out, iters, 1.0, 1e-9); }));
    rep("dmma_m8n8k4_ch16_tflops", 2.0 * 8 * 8 * 4, 16, time_ms([&] { k_dmma<0, 16><<<blocks, threads>>>(out, iters, 1.0, 1e-9); }));
    rep("dmma_m16n8k4_ch8_tflops", 2.0 * 16 * 8 * 4, 8, time_ms([&] { k_dmma<1, 8><<<blocks, threads>>>(out, iters, 1.0, 1e-9); }));
    rep("dmma_m16n8k8_ch8_tflops", 2.0 * 16 * 8 * 8, 8, time_ms([&] { k_dmma<2, 8><<<blocks, threads>>>(out, iters, 1.0, 1e-9); }));
    rep("dmma_m16n8k16_ch8_tflops", 2.0 * 16 * 8 * 16, 8, time_ms([&] { k_dmma<3, 8><<<blocks, threads>>>(out, iters / 2, 1.0, 1e-9); }) * 2);
#ifdef WITH_CUBLAS
    if (argc > 1 && !strcmp(argv[1], "--cublas")) {
        cublasHandle_t h; cublasCreate(&h);
        for (int n : {65, 513, 1025, 2049, 4097, 8193}) {
            double *A, *B, *C;
            const size_t ld = (size_t)(n + 15) / 16 * 16;
            CK(cudaMalloc(&A, ld * n * 8)); CK(cudaMalloc(&B, ld * n * 8)); CK(cudaMalloc(&C, ld * n * 8));
            CK(cudaMemset(A, 0, ld * n * 8)); CK(cudaMemset(B, 0, ld * n * 8));
            const double one = 1, zero = 0;
            const double ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, (int)ld, B, (int)ld, &zero, C, (int)ld); }, 3);
            printf(",\n \"cublas_dgemm_nt_%d_ms\": %.4f, \"cublas_dgemm_nt_%d_tflops\": %.2f", n, ms, n, 2.0 * n * n * n / ms * 1e-9);
            const double ms2 = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, (int)ld, B, (int)ld, &zero, C, (int)ld); }, 3);
            printf(",\n \"cublas_dgemm_nn_%d_ms\": %.4f, \"cublas_dgemm_nn_%d_tflops\": %.2f", n, ms2, n, 2.0 * n * n * n / ms2 * 1e-9);
            cudaFree(A); cudaFree(B); cudaFree(C);
        }
        cublasDestroy(h);
    }
#endif
    printf("\n}\n");
    return 0;
}
