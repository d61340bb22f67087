// tools/issue_probe.cu -- does an FP64 instruction cost a warp scheduler of B200 one issue slot or two?  (the
// stage kernel executes 574 FP64 and 410 other warp instructions per 32 points; if the others can be issued in the cycle the FP64
// pipe -- 16 lanes per scheduler, two cycles per warp instruction -- is busy anyway, the kernel's floor is the FP64 pipe (4.3 ms at
// 512^3); if not, it is 574 x 2 + 410 issue slots (5.9 ms).)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_probe tools/issue_probe.cu && ./issue_probe
// Legs: NF FP64 FMAs and NI independent integer ops (or shared-memory loads) per loop trip, interleaved, every chain independent
// (8 FP64 + 8 integer accumulators per thread), for 2 and for 4 warps per scheduler.  Reported: cycles per loop trip per scheduler.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int NI, bool LDS>
__global__ void __launch_bounds__(256) leg(double *out, int iters, double a, double b, int seed) {
    __shared__ int sh[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) sh[i] = i ^ seed;
    __syncthreads();
    double x[8];
    int n[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x * 1e-3 + i; n[i] = threadIdx.x * 7 + i + seed; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {                   // 32 FP64 FMAs per trip
#pragma unroll
            for (int i = 0; i < 8; i++) {
                x[i] = fma(x[i], a, b);
                if ((r * 8 + i) < NI) {                 // NI of the 32 slots get a companion instruction
                    if (LDS) n[i] += sh[(n[i] + threadIdx.x) & 1023];
                    else n[i] = n[i] * 3 + seed;      // IMAD
                }
            }
        }
    }
    double s = 0.0; int m = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { s += x[i]; m ^= n[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + m;
}

template <int NI, bool LDS>
static void run(const char *what, int warps_per_sched, double *out) {
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const int blocks = sms * warps_per_sched / 2;        // 256 threads = 8 warps = 2 per scheduler
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    leg<NI, LDS><<<blocks, 256>>>(out, 100, 1.0000001, 1e-9, 3);
    cudaEventRecord(e0);
    leg<NI, LDS><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    // per scheduler and loop trip: warps_per_sched warps x (32 FP64 + NI others)
    const double cyc = ms * 1e-3 * khz * 1e3 / iters;
    printf("issue_probe %-22s %d warps/scheduler: %6.1f cycles per trip (nominal clock) = %.2f per FP64 instruction; FP64-pipe floor %d, one-slot-each floor %d, two-slots-per-FP64 floor %d\n",
           what, warps_per_sched, cyc, cyc / (32.0 * warps_per_sched), 64 * warps_per_sched, (32 + NI) * warps_per_sched, (64 + NI) * warps_per_sched);
}

int main() {
    double *out; cudaMalloc(&out, sizeof(double) * 256 * 1024);
    for (int w = 2; w <= 4; w += 2) {
        run<0, false>("32 DFMA", w, out);
        run<16, false>("32 DFMA + 16 IMAD", w, out);
        run<32, false>("32 DFMA + 32 IMAD", w, out);
        run<16, true>("32 DFMA + 16 LDS+IADD", w, out);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
