// Micro-probe of the FP64 tensor pipe on sm_100a: DMMA.8x8x4 throughput versus warps/SM, independent
// accumulators per warp and operand register pattern.  Build: nvcc -arch=sm_100a -O3 -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void same_ab(long iters, double* sink) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (long it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[0] = s;
}

// 4 x 8 register-blocked pattern of the real kernel (32 accumulators, 4 a's, 8 b's)
__global__ void blocked_4x8(long iters, double* sink) {
    double c[4][8][2];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int u = 0; u < 8; ++u) c[t][u][0] = c[t][u][1] = 0.0;
    double a[4], b[8];
#pragma unroll
    for (int t = 0; t < 4; ++t) a[t] = 1.0 + (threadIdx.x + t) * 1e-9;
#pragma unroll
    for (int u = 0; u < 8; ++u) b[u] = 1.0 - (threadIdx.x + u) * 1e-9;
    for (long it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int u = 0; u < 8; ++u) dmma(c[t][u][0], c[t][u][1], a[t], b[u]);
    }
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int u = 0; u < 8; ++u) s += c[t][u][0] + c[t][u][1];
    if (s == 123.456) sink[0] = s;
}

// 2 x 4 blocking (8 accumulators) and 4 x 4 (16)
template <int TM, int TN>
__global__ void blocked(long iters, double* sink) {
    double c[TM][TN][2];
#pragma unroll
    for (int t = 0; t < TM; ++t)
#pragma unroll
        for (int u = 0; u < TN; ++u) c[t][u][0] = c[t][u][1] = 0.0;
    double a[TM], b[TN];
#pragma unroll
    for (int t = 0; t < TM; ++t) a[t] = 1.0 + (threadIdx.x + t) * 1e-9;
#pragma unroll
    for (int u = 0; u < TN; ++u) b[u] = 1.0 - (threadIdx.x + u) * 1e-9;
    for (long it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < TM; ++t)
#pragma unroll
            for (int u = 0; u < TN; ++u) dmma(c[t][u][0], c[t][u][1], a[t], b[u]);
    }
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < TM; ++t)
#pragma unroll
        for (int u = 0; u < TN; ++u) s += c[t][u][0] + c[t][u][1];
    if (s == 123.456) sink[0] = s;
}

__global__ void dfma16(long iters, double* sink) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (long it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456) sink[0] = s;
}

template <typename F>
double time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double* sink;
    cudaMalloc(&sink, 64);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"rows\": [\n", p.name, sms);
    const long iters = 4000;
    auto report = [&](const char* name, int warps_per_sm, int nacc, double ms, bool last = false) {
        const double flops = (double)sms * warps_per_sm * iters * nacc * 512.0;
        printf(" {\"kernel\": \"%s\", \"warps_per_sm\": %d, \"acc_per_warp\": %d, \"tflops\": %.2f}%s\n", name, warps_per_sm,
               nacc, flops / (ms * 1e-3) / 1e12, last ? "" : ",");
    };
    for (int wps : {4, 8, 16, 32}) {
        const int threads = 128, ctas = sms * (wps / 4);
        report("same_ab", wps, 4, time_ms([&] { same_ab<4><<<ctas, threads>>>(iters, sink); }));
        report("same_ab", wps, 8, time_ms([&] { same_ab<8><<<ctas, threads>>>(iters, sink); }));
        report("same_ab", wps, 16, time_ms([&] { same_ab<16><<<ctas, threads>>>(iters, sink); }));
        report("same_ab", wps, 32, time_ms([&] { same_ab<32><<<ctas, threads>>>(iters, sink); }));
        report("blocked_2x4", wps, 8, time_ms([&] { blocked<2, 4><<<ctas, threads>>>(iters, sink); }));
        report("blocked_4x4", wps, 16, time_ms([&] { blocked<4, 4><<<ctas, threads>>>(iters, sink); }));
        report("blocked_2x8", wps, 16, time_ms([&] { blocked<2, 8><<<ctas, threads>>>(iters, sink); }));
        report("blocked_4x8", wps, 32, time_ms([&] { blocked_4x8<<<ctas, threads>>>(iters, sink); }));
    }
    {
        const double ms = time_ms([&] { dfma16<<<sms * 8, 256>>>(iters * 8, sink); });
        const double flops = (double)sms * 8 * 256 * iters * 8 * 16 * 2.0;
        printf(" {\"kernel\": \"dfma16\", \"warps_per_sm\": 64, \"acc_per_warp\": 16, \"tflops\": %.2f}\n", flops / (ms * 1e-3) / 1e12);
    }
    printf("]}\n");
    return 0;
}
