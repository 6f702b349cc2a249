// Developer micro-benchmark: fp64 tensor-core (DMMA m8n8k4) throughput of one CTA per SM as a function of the warps per
// scheduler and of the independent accumulator chains per warp -- how many warps does an MMA-bound kernel need?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_occupancy tools/dmma_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int ILP>
__global__ void k(double* out, int iters, double a, double b) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    if (s == 1.2345) out[0] = s;
}
template <int ILP>
void run(int warps, int sms) {
    double* out; cudaMalloc(&out, 8);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP><<<sms, 32 * warps>>>(out, 100, 1.0, 1e-9);
    cudaEventRecord(e0);
    k<ILP><<<sms, 32 * warps>>>(out, iters, 1.0, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = (double)sms * warps * iters * ILP * 512.0;
    printf("warps/CTA %2d (%d per scheduler)  chains/warp %2d : %6.2f TFLOP/s\n", warps, warps / 4, ILP, flops / (ms * 1e-3) / 1e12);
    cudaFree(out);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int w : {4, 8, 16, 32}) { run<2>(w, sms); run<6>(w, sms); run<12>(w, sms); run<18>(w, sms); }
    return 0;
}
