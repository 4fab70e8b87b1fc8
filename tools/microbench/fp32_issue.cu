// Microbenchmark: issue throughput of scalar FADD/FMUL vs packed FADD2/FMUL2 on sm_100a with FMA
// contraction disabled (the parity constraint of this project).  Prints G lane-ops/s per variant.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 1e-3f + i; b[i] = 1.0f + 1e-7f * i; }
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {        // 8 independent FADD + 8 FMUL per iter (scalar)
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = __fadd_rn(a[i], b[i]); a[i] = __fmul_rn(a[i], b[i]); }
        } else if (MODE == 1) { // packed: 4 FADD2 + 4 FMUL2 (same lane-ops)
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                float2 x = make_float2(a[i], a[i + 1]), y = make_float2(b[i], b[i + 1]);
                x = __fadd2_rn(x, y); x = __fmul2_rn(x, y);
                a[i] = x.x; a[i + 1] = x.y;
            }
        } else if (MODE == 2) { // FFMA reference (fused; not usable for parity)
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = __fmaf_rn(a[i], b[i], b[i]); a[i] = __fmaf_rn(a[i], b[i], b[i]); }
        } else {                // FFMA2
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                float2 x = make_float2(a[i], a[i + 1]), y = make_float2(b[i], b[i + 1]);
                x = __ffma2_rn(x, y, y); x = __ffma2_rn(x, y, y);
                a[i] = x.x; a[i + 1] = x.y;
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name) {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(float));
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(d, 100, 1.0f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d, iters, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double laneops = 148.0 * 8 * 256 * (double)iters * 16;
    printf("%-8s %8.3f ms  %8.1f G lane-ops/s  (%.1f lane-ops/clk/SM at 1.9 GHz)\n", name, ms,
           laneops / ms / 1e6, laneops / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(d);
}

int main() {
    run<0>("scalar");
    run<1>("packed");
    run<2>("ffma");
    run<3>("ffma2");
    return 0;
}
