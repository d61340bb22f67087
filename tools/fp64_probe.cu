// tools/fp64_probe.cu -- round-2 groundwork: how do the FP64 tensor-core path (DMMA, mma.sync
// m8n8k4) and the FP64 pipe (DFMA) compare on B200 in throughput and in power?  The stage kernel's sustained rate is set by the
// 1000 W cap, and 216 of its 564 FP64 operations per point are linear stencil sums that could be written as banded products.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe tools/fp64_probe.cu && ./fp64_probe [seconds-per-leg]
// Each leg runs for the given time (default 3 s) so that `nvidia-smi --query-gpu=power.draw,clocks.sm -lms 200` sees it.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) dfma_leg(double *out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// D(8x8) += A(8x4) B(4x8): one A and one B element per thread, two accumulators per thread
__global__ void __launch_bounds__(256) dmma_leg(double *out, int iters, double a, double b) {
    double c[4][2];
#pragma unroll
    for (int i = 0; i < 4; i++) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = -c[i][0]; }
    const double fa = a + (threadIdx.x & 3) * 1e-9, fb = b + (threadIdx.x >> 2) * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(fa), "d"(fb));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// both in every thread: do the DMMA tensor path and the FP64 pipe run side by side (rates add) or share one datapath (rates do not)?
__global__ void __launch_bounds__(256) mixed_leg(double *out, int iters, double a, double b) {
    double x[8], c[4][2];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 4; i++) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = -c[i][0]; }
    const double fa = a + (threadIdx.x & 3) * 1e-9, fb = b + (threadIdx.x >> 2) * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(fa), "d"(fb));
            x[2 * i] = fma(x[2 * i], a, b); x[2 * i + 1] = fma(x[2 * i + 1], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
#pragma unroll
    for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main(int argc, char **argv) {
    const double secs = argc > 1 ? atof(argv[1]) : 3.0;
    int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
    double *out; cudaMalloc(&out, (size_t)blocks * threads * sizeof(double));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int leg = 0; leg < 3; leg++) {
        double total_ms = 0.0; long launches = 0;
        while (total_ms < secs * 1e3) {
            cudaEventRecord(e0);
            if (leg == 0) dfma_leg<<<blocks, threads>>>(out, iters, 0.999999, 1e-6);
            else if (leg == 1) dmma_leg<<<blocks, threads>>>(out, iters, 0.999999, 1e-6);
            else mixed_leg<<<blocks, threads>>>(out, iters, 0.999999, 1e-6);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); total_ms += ms; launches++;
        }
        // flops: DFMA leg 8 FMA/thread/iter; DMMA leg 4 x (8*8*4 FMA per warp) per iter
        const double f_pipe = (double)launches * blocks * threads * iters * 8.0, f_mma = (double)launches * blocks * (threads / 32) * iters * 4.0 * 256.0;
        const double fma = leg == 0 ? f_pipe : leg == 1 ? f_mma : f_pipe + f_mma;
        printf("fp64_probe %s: %.1f TFLOP/s over %.1f s (%ld launches)\n", leg == 0 ? "DFMA pipe " : leg == 1 ? "DMMA m8n8k4" : "DFMA + DMMA interleaved (sum of both)", 2.0 * fma / (total_ms * 1e-3) / 1e12,
               total_ms * 1e-3, launches);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
    }
    return 0;
}
