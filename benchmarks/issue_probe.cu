// issue_probe.cu -- does an FP64 instruction cost the warp scheduler one issue slot or two?
//
// DESIGN.md section 4 models ce_sample_kernel as "cycles = all instructions + FP64 instructions, per scheduler".  This
// probe checks the model away from that kernel: loops of independent DFMA chains, of independent integer (add / xor) chains,
// and of both interleaved, eight chains of each kind per thread and sixteen warps per scheduler so that latency is hidden.
// If integer instructions could issue in the shadow of an FP64 instruction, the mixed loop would take as long as its
// DFMAs alone; if an FP64 instruction holds the scheduler for two cycles, it takes the sum.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/issue_probe benchmarks/issue_probe.cu && build/issue_probe
//
// Prints cycles per loop trip per scheduler for each mix (SM clock read with clock64 inside the kernel).
#include <cstdio>
#include <cuda_runtime.h>

template <int NF, int NI> __global__ void __launch_bounds__(256) probe(double *out, long long *cycles, int trips, double a, double b, unsigned m)
{
    double f[8];
    unsigned u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        f[i] = threadIdx.x + i;
        u[i] = threadIdx.x * 8 + i;
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int t = 0; t < trips; ++t) {
        // one trip = 16 steps; a step issues NF DFMAs and NI integer adds, on rotating chains
#pragma unroll
        for (int s = 0; s < 16; ++s) {
#pragma unroll
            for (int k = 0; k < NF; ++k) {
                double &x = f[(s * NF + k) & 7];
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(a), "d"(b));
            }
#pragma unroll
            for (int k = 0; k < NI; ++k) {
                // add and xor alternate on a chain, so that ptxas cannot fold consecutive operations into one
                unsigned &y = u[(s * NI + k) & 7];
                if (((s * NI + k) >> 3) & 1) asm volatile("xor.b32 %0, %0, %1;" : "+r"(y) : "r"(m));
                else asm volatile("add.u32 %0, %0, %1;" : "+r"(y) : "r"(m));
            }
        }
    }
    const long long t1 = clock64();
    double sf = 0.0;
    unsigned su = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sf += f[i];
        su += u[i];
    }
    if (sf == a) out[0] = sf; // keep the chains alive: neither condition can be decided at compile time
    if (su == m) out[1] = (double)su;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int NF, int NI> static void run(const char *what, int sms)
{
    const int ctas_per_sm = 8, trips = 4096; // 8 x 256 threads fill an SM: 64 warps, 16 per scheduler (a partial fill is placed unevenly)
    const int grid = sms * ctas_per_sm;
    double *out;
    long long *cyc, *h = new long long[grid];
    cudaMalloc(&out, 16);
    cudaMalloc(&cyc, sizeof(long long) * grid);
    int resident = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, probe<NF, NI>, 256, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<NF, NI><<<grid, 256>>>(out, cyc, 64, 1.0000001, 1e-9, 3u); // warm-up
    cudaEventRecord(e0);
    probe<NF, NI><<<grid, 256>>>(out, cyc, trips, 1.0000001, 1e-9, 3u);
    cudaEventRecord(e1);
    cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    double mean = 0.0;
    for (int i = 0; i < grid; ++i) mean += (double)h[i];
    mean /= grid;
    // every SM runs its 8 CTAs (64 warps, 16 per scheduler) side by side for `mean` cycles: cycles per scheduler divided
    // by the trips of its 16 warps.  The same from the kernel's duration and the SM clock it implies is printed beside it.
    const double per_warp_trip = mean / trips / 16.0;
    const double ghz = mean / (ms * 1e6);
    printf("%-14s %d DFMA + %d integer per step: %6.2f scheduler cycles per warp-trip = %5.2f per step  (kernel %.3f ms, "
           "CTA clock64 span %.0f cycles -> %.2f GHz if the CTAs ran side by side; %d CTAs can be resident per SM)  %s\n",
           what, NF, NI, per_warp_trip, per_warp_trip / 16.0, ms, mean, ghz, resident, cudaGetErrorString(cudaGetLastError()));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    cudaFree(cyc);
    delete[] h;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs; per scheduler: 16 warps, 8 independent chains of each kind per thread\n", p.name, p.multiProcessorCount);
    run<1, 0>("FP64 only", p.multiProcessorCount);
    run<0, 1>("integer only", p.multiProcessorCount);
    run<0, 2>("integer only", p.multiProcessorCount);
    run<1, 1>("mixed 1:1", p.multiProcessorCount);
    run<1, 2>("mixed 1:2", p.multiProcessorCount);
    run<2, 1>("mixed 2:1", p.multiProcessorCount);
    run<1, 4>("mixed 1:4", p.multiProcessorCount);
    return 0;
}
