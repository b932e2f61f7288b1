// fp64_peak.cu — measured FP64 peaks of this GPU for the roofline of the sparse-GP sweep (aug_sparse.cu):
// DMMA (mma.sync.m8n8k4.f64) and plain DFMA issue rates, all SMs, CUDA events.  Standalone tool:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_peak tools/fp64_peak.cu && tools/fp64_peak
#include <cuda_runtime.h>
#include <stdio.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void __launch_bounds__(512, 1) k_dmma(double* out, int iters, double a0, double b0) {
    double c[CHAINS][2];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) c[i][0] = c[i][1] = 0.0;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}


// the same stream with DISTINCT A / B operand registers per chain (the real kernels never reuse an operand register pair)
template <int CHAINS>
__global__ void __launch_bounds__(512, 1) k_dmma_distinct(double* out, int iters, double a0, double b0) {
    double c[CHAINS][2], a[CHAINS], b[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { c[i][0] = c[i][1] = 0.0; a[i] = a0 + (threadIdx.x + 7 * i) * 1e-9; b[i] = b0 + i * 1e-7; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) dmma(c[i][0], c[i][1], a[i], b[i]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

template <int CHAINS>
__global__ void __launch_bounds__(512, 1) k_dfma(double* out, int iters, double a0, double b0) {
    double c[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) c[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c[i];
    if (s == 12345.678) out[0] = s;
}


// How long does a scalar FP64 instruction of one warp wait while the other warps of the SM stream DMMAs?
// warp 0: `iters` dependent DFMAs (clock64 around them); warps 1..: DMMA streams until warp 0 is done (flag in smem).
__global__ void __launch_bounds__(512, 1) k_mix(double* out, long long* cyc, int iters, int dmma_warps, double a0) {
    __shared__ volatile int done;
    if (threadIdx.x == 0) done = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        double c = a0;
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) c = fma(c, 1.0000001, 1e-9);
        long long t1 = clock64();
        if (c == 12345.678) out[0] = c;
        if (threadIdx.x == 0) { cyc[blockIdx.x] = t1 - t0; done = 1; }
    } else if (warp <= dmma_warps) {
        double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        double a = a0 + threadIdx.x * 1e-9, b = 1.0;
        while (!done) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) dmma(c[i][0], c[i][1], a, b);
        }
        if (c[0][0] + c[1][0] + c[2][1] + c[3][1] == 12345.678) out[1] = c[0][0];
    }
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* out; cudaMalloc(&out, 64);
    const int iters = 20000;
    for (int warps = 4; warps <= 16; warps *= 2) {
        float ms = time_ms([&] { k_dmma<8><<<sms, warps * 32, 0>>>(out, iters, 1.0, 1.0); });
        double flops = 2.0 * 256 * 8 * (double)iters * warps * sms;
        printf("{\"kernel\": \"dmma_m8n8k4\", \"warps_per_sm\": %d, \"chains\": 8, \"ms\": %.4f, \"tflops\": %.3f}\n", warps, ms, flops / ms * 1e-9);
    }
    {
        float ms = time_ms([&] { k_dmma<2><<<sms, 512, 0>>>(out, iters, 1.0, 1.0); });
        double flops = 2.0 * 256 * 2 * (double)iters * 16 * sms;
        printf("{\"kernel\": \"dmma_m8n8k4\", \"warps_per_sm\": 16, \"chains\": 2, \"ms\": %.4f, \"tflops\": %.3f}\n", ms, flops / ms * 1e-9);
    }
    for (int warps = 4; warps <= 16; warps *= 2) {
        float ms = time_ms([&] { k_dmma_distinct<8><<<sms, warps * 32, 0>>>(out, iters, 1.0, 1.0); });
        double flops = 2.0 * 256 * 8 * (double)iters * warps * sms;
        printf("{\"kernel\": \"dmma_m8n8k4_distinct_operands\", \"warps_per_sm\": %d, \"chains\": 8, \"ms\": %.4f, \"tflops\": %.3f}\n", warps, ms, flops / ms * 1e-9);
    }
    for (int warps = 8; warps <= 16; warps *= 2) {
        float ms = time_ms([&] { k_dfma<8><<<sms, warps * 32, 0>>>(out, iters, 1.0000001, 1e-9); });
        double flops = 2.0 * 32 * 8 * (double)iters * warps * sms;
        printf("{\"kernel\": \"dfma\", \"warps_per_sm\": %d, \"chains\": 8, \"ms\": %.4f, \"tflops\": %.3f}\n", warps, ms, flops / ms * 1e-9);
    }
    {
        long long* cyc; cudaMalloc(&cyc, sizeof(long long) * sms);
        for (int dw = 0; dw <= 15; dw = dw == 0 ? 3 : (dw == 3 ? 7 : (dw == 7 ? 15 : 16))) {
            k_mix<<<sms, 512>>>(out, cyc, 2000, dw, 1.0);
            cudaDeviceSynchronize();
            long long h[256]; cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < sms; ++i) avg += (double)h[i];
            printf("{\"kernel\": \"dependent_dfma_next_to_dmma\", \"dmma_warps_per_sm\": %d, \"clk_per_dfma\": %.1f}\n", dw, avg / sms / 2000.0);
        }
    }
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
    return 0;
}
