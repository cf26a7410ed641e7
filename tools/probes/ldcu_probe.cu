// Microbenchmark: FP64 FMA throughput when the broadcast operand comes from the kernel-parameter
// constant bank through the uniform datapath (LDCU.64 -> DFMA with a UR operand) instead of shared
// memory.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldcu_probe ldcu_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
struct Cores { double g[3600]; };
template <int QPT, int W>
__global__ void __launch_bounds__(512) k(const __grid_constant__ Cores C, const double* __restrict__ x,
                                         double* __restrict__ out, int rows, int reps) {
    double c0[QPT], c1[QPT], tw[QPT], acc[QPT][W];
    for (int q = 0; q < QPT; ++q) {
        c0[q] = x[(threadIdx.x + blockIdx.x * blockDim.x) * QPT + q]; c1[q] = c0[q] * 0.5; tw[q] = 1.0;
        for (int l = 0; l < W; ++l) acc[q][l] = 0.0;
    }
    for (int r = 0; r < reps; ++r) {
        const double* gp = C.g;
        for (int t = 0; t < rows; ++t) {
#pragma unroll
            for (int l = 0; l < W; ++l)
#pragma unroll
                for (int q = 0; q < QPT; ++q) acc[q][l] = fma(c0[q], gp[l], acc[q][l]);
#pragma unroll
            for (int q = 0; q < QPT; ++q) { double c2 = fma(tw[q], c1[q], -c0[q]); c0[q] = c1[q]; c1[q] = c2; }
            gp += W;
        }
    }
    double s = 0;
    for (int q = 0; q < QPT; ++q) for (int l = 0; l < W; ++l) s += acc[q][l];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
}
extern int g_doubles;
template <int QPT, int W> void run(int threads, int blocks_per_sm) {
    Cores* h = new Cores; for (int i = 0; i < 3600; ++i) h->g[i] = 1e-3 * (i % 17);
    int sms = 148, blocks = sms * blocks_per_sm, rows = g_doubles / W, reps = 200 * 3600 / g_doubles;
    double *x, *o; cudaMalloc(&x, sizeof(double) * blocks * threads * QPT); cudaMalloc(&o, sizeof(double) * blocks * threads);
    cudaMemset(x, 0, sizeof(double) * blocks * threads * QPT);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int it = 0; it < 4; ++it) {
        cudaEventRecord(e0); k<QPT, W><<<blocks, threads>>>(*h, x, o, rows, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (it && ms < best) best = ms;
    }
    double fma = (double)blocks * threads * QPT * (W + 1) * rows * reps;
    printf("QPT=%d W=%d threads=%d x%d/SM: %.3f ms  %.2f TFLOP/s  (%s)\n", QPT, W, threads, blocks_per_sm, best,
           2 * fma / best / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(x); cudaFree(o); delete h;
}
int g_doubles = 3600;
int main(int argc, char** argv) {
    for (int sz : {500, 1000, 1500, 2000, 2500, 3000, 3300, 3600}) { g_doubles = sz; printf("table %d doubles (%d B): ", sz, sz * 8); run<2, 11>(512, 1); }
    return 0;
}
