// FP64 issue rate of the CUDA cores of this GPU (not the tensor cores): independent DMUL / DADD / DFMA chains, no memory.
// Decides whether an fp64 kernel with little traffic per flop (Gram block, matrix-powers basis) is HBM- or pipe-bound.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false tools/fp64_rate.cu -o /tmp/fp64_rate && /tmp/fp64_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>   // 0: a = a * b + c as DMUL + DADD (fmad=false, what libpkrylov does); 1: DFMA
__global__ void k(double* out, int iters, double b, double c) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = a[i] * b + c;
            else a[i] = fma(a[i], b, c);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 4096;
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
            else k<1><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double ops = (double)blocks * threads * iters * 8;     // a*b+c per op
        const double instr = ops * (mode == 0 ? 2 : 1);
        printf("%s: %.3f ms, %.2f T thread-instr/s = %.2f warp-instr/clk/SM at %d MHz, %.2f TFLOP/s\n",
               mode == 0 ? "DMUL+DADD (fmad=false)" : "DFMA", ms, instr / ms / 1e9,
               instr / 32 / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3), p.clockRate / 1000, 2 * ops / ms / 1e9);
    }
    return 0;
}
