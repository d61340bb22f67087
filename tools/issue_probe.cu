// tools/issue_probe.cu -- how much room does a warp scheduler of B200 have next to its FP64 instructions?  (the
// stage kernel executes 574 FP64 and 410 other warp instructions per 32 points.  The FP64 pipe has 16 lanes per scheduler: a warp
// instruction keeps it busy for two cycles.  If the scheduler can issue another instruction to another pipe in the second cycle,
// the kernel's floor is the FP64 pipe -- 574 x 2 cycles, 4.3 ms at 512^3; if an FP64 instruction also costs two ISSUE slots, the floor
// is 574 x 2 + 410 slots, 5.9 ms.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_probe tools/issue_probe.cu && ./issue_probe
// Legs: 32 independent FP64 FMAs per loop trip (8 accumulators x 4) with NI companion instructions interleaved one-to-one --
// FP32 adds (full-rate pipe), 32-bit logic ops (full-rate integer pipe) or integer multiply-adds (half-rate) on 8 independent chains
// -- at 2 and 4 warps per scheduler.  Cycles are read on the device (clock64 of one warp), so the result does not depend on the clock.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

enum { FADD = 0, LOP = 1, IMAD = 2 };

template <int NI, int KIND>
__global__ void __launch_bounds__(512) leg(double *out, long long *cyc, int iters, double a, double b, int seed, float fs) {
    double x[8];
    int n[8]; float f[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x * 1e-3 + i; n[i] = threadIdx.x * 7 + i + seed; f[i] = threadIdx.x * 0.5f + i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                x[i] = fma(x[i], a, b);
                if (((r * 8 + i) % (32 / (NI > 0 ? NI : 32))) == 0 && NI > 0) {      // NI companions, evenly spread
                    if (KIND == FADD) f[i] = f[i] + fs;
                    else if (KIND == LOP) n[i] = (n[i] ^ seed) + 1;                   // LOP3 + IADD: counted as two below
                    else n[i] = n[i] * 3 + seed;
                }
            }
        }
    }
    const long long t1 = clock64();
    double s = 0.0; int m = 0; float g = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) { s += x[i]; m ^= n[i]; g += f[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + m + g;
    if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = t1 - t0;
}

template <int NI, int KIND>
static void run(const char *what, int warps_per_sched, double *out, long long *cyc) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int threads = 128 * warps_per_sched;           // ONE block per SM: 4 schedulers x warps_per_sched warps, co-resident by construction
    const int iters = 20000;
    leg<NI, KIND><<<sms, threads>>>(out, cyc, 100, 1.0000001, 1e-9, 3, 1e-3f);
    leg<NI, KIND><<<sms, threads>>>(out, cyc, iters, 1.0000001, 1e-9, 3, 1e-3f);
    long long c = 0; cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double per = (double)c / iters;
    const int nc = KIND == LOP ? 2 * NI : NI;             // companion instructions per warp and trip
    printf("issue_probe %-22s %d warps/scheduler: %6.1f cycles per trip | FP64 pipe alone %d | if companions co-issue %d | if an FP64 instruction takes two issue slots %d\n",
           what, warps_per_sched, per, 64 * warps_per_sched, (64 > 32 + nc ? 64 : 32 + nc) * warps_per_sched, (64 + nc) * warps_per_sched);
}

int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, sizeof(double) * 512 * 1024); cudaMalloc(&cyc, sizeof(long long));
    for (int w = 2; w <= 4; w += 2) {
        run<0, FADD>("32 DFMA", w, out, cyc);
        run<16, FADD>("32 DFMA + 16 FADD", w, out, cyc);
        run<32, FADD>("32 DFMA + 32 FADD", w, out, cyc);
        run<16, LOP>("32 DFMA + 16 LOP3+IADD", w, out, cyc);
        run<32, IMAD>("32 DFMA + 32 IMAD", w, out, cyc);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
