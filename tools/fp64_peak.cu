// FP64 micro-benchmarks for the roofline denominators that MEASURED_PEAKS.json lacks.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
// Prints one JSON object: DFMA (vector pipe) peak, DMMA (mma.sync f64) peak, both together,
// fp64 reciprocal and 64-bit shuffle throughput.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

// mma.sync m8n8k4 f64: 256 MAC per warp instruction
template <int ILP>
__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double a, double b) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
    if (s == 12345.678) out[0] = s;
}

// mma.sync m16n8k8 f64 (sm_90+ shape): 1024 MAC per warp instruction
template <int ILP>
__global__ void __launch_bounds__(256) k_dmma1688(double* out, int iters, double a, double b) {
    double c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; i++) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

// half the warps DFMA, half DMMA: do the two share a pipe?
__global__ void __launch_bounds__(256) k_mixed(double* out, int iters, double a, double b) {
    const int ILP = 8;
    int warp = threadIdx.x >> 5;
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
    if (warp & 1) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
            }
        }
    } else {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) { c0[i] = fma(c0[i], a, b); c1[i] = fma(c1[i], a, b); }
#pragma unroll
            for (int i = 0; i < ILP; i++) { c0[i] = fma(c0[i], a, b); c1[i] = fma(c1[i], a, b); }
#pragma unroll
            for (int i = 0; i < ILP; i++) { c0[i] = fma(c0[i], a, b); c1[i] = fma(c1[i], a, b); }
#pragma unroll
            for (int i = 0; i < ILP; i++) { c0[i] = fma(c0[i], a, b); c1[i] = fma(c1[i], a, b); }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
    if (s == 12345.678) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) k_drcp(double* out, int iters, double a) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = 1.5 + threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = 1.0 / (acc[i] + a);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) k_shfl64(double* out, int iters) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = 1.5 + threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = __shfl_xor_sync(0xffffffffu, acc[i], 1);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

// smem-fed DFMA: one broadcast LDS.64 per DFMA (the "weights in shared memory" pattern)
__global__ void __launch_bounds__(256) k_dfma_lds(double* out, int iters, double b) {
    __shared__ double w[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) w[i] = 1.0 + i * 1e-9;
    __syncthreads();
    const int ILP = 16;
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++) {
        const double* wp = w + (it & 31) * 16;
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], wp[i], b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 64));
    int iters = 20000;
    int grid = sms * 8, block = 256;
    double nthreads = (double)grid * block;

    double t_dfma = time_ms([&] { k_dfma<16><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
    double dfma_tf = nthreads * 16.0 * iters * 2 / (t_dfma * 1e-3) / 1e12;

    double t_884 = time_ms([&] { k_dmma884<8><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
    double dmma884_tf = (nthreads / 32) * 8.0 * iters * 256 * 2 / (t_884 * 1e-3) / 1e12;

    double t_1688 = time_ms([&] { k_dmma1688<4><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
    double dmma1688_tf = (nthreads / 32) * 4.0 * iters * 1024 * 2 / (t_1688 * 1e-3) / 1e12;

    double t_mixed = time_ms([&] { k_mixed<<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
    // per iteration: odd warps 8 mma (8*256 MAC), even warps 64 dfma (64*32 MAC) -> equal MACs
    double mixed_tf = (nthreads / 32) * iters * 2048.0 * 2 / (t_mixed * 1e-3) / 1e12;

    double t_rcp = time_ms([&] { k_drcp<8><<<grid, block>>>(out, iters / 10, 0.25); });
    double rcp_g = nthreads * 8.0 * (iters / 10) / (t_rcp * 1e-3) / 1e9;

    double t_sh = time_ms([&] { k_shfl64<8><<<grid, block>>>(out, iters); });
    double sh_g = nthreads * 8.0 * iters / (t_sh * 1e-3) / 1e9;

    double t_lds = time_ms([&] { k_dfma_lds<<<grid, block>>>(out, iters, 1e-9); });
    double lds_tf = nthreads * 16.0 * iters * 2 / (t_lds * 1e-3) / 1e12;

    int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_attr\": %d, \"dfma_tflops\": %.3f, \"dmma_m8n8k4_tflops\": %.3f, "
           "\"dmma_m16n8k8_tflops\": %.3f, \"dfma_plus_dmma_tflops\": %.3f, \"drcp_gops\": %.2f, \"shfl64_gops\": %.2f, "
           "\"dfma_lds_broadcast_tflops\": %.3f}\n",
           prop.name, sms, clk, dfma_tf, dmma884_tf, dmma1688_tf, mixed_tf, rcp_g, sh_g, lds_tf);
    return 0;
}
