// Microbenchmark: issue rate of scalar FFMA/FADD/FMUL vs packed FFMA2/FADD2/FMUL2 and MUFU on sm_100a.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o f32x2 f32x2.cu && ./f32x2
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int CH = 8;   // independent chains per thread

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b) {
    float x[CH], y[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { x[i] = threadIdx.x * 1e-3f + i; y[i] = x[i] + 0.5f; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (MODE == 0) { x[i] = fmaf(x[i], a, b); y[i] = fmaf(y[i], a, b); }
            if (MODE == 1) { x[i] = __fadd_rn(x[i], a); y[i] = __fadd_rn(y[i], a); }
            if (MODE == 2) { x[i] = __fmul_rn(x[i], a); y[i] = __fmul_rn(y[i], a); }
            if (MODE >= 3 && MODE <= 5) {
                unsigned long long v, aa, bb;
                asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x[i]), "f"(y[i]));
                asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
                asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
                if (MODE == 3) asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(aa), "l"(bb));
                if (MODE == 4) asm("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(aa));
                if (MODE == 5) asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(aa));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(x[i]), "=f"(y[i]) : "l"(v));
            }
            if (MODE == 6) {
                asm("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
                asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(y[i]));
            }
            if (MODE == 7) {   // mixed: packed fma + alu-pipe op
                unsigned long long v, aa, bb;
                asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x[i]), "f"(y[i]));
                asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
                asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
                asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(aa), "l"(bb));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(x[i]), "=f"(y[i]) : "l"(v));
                x[i] = fminf(x[i], 3.0e38f);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float* out, int flops_per_op) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    dim3 grid(sms * 8), block(256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, block>>>(out, 1.0000001f, 1e-7f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<grid, block>>>(out, 1.0000001f, 1e-7f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double elems = (double)grid.x * 256 * ITERS * CH * 2;     // scalar results produced
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double per_clk_sm = elems / (ms * 1e-3) / (clk * 1e3) / sms;
    printf("%-28s %8.3f ms  %7.1f results/clk/SM (at %d MHz nominal)  %6.2f Tresults/s\n", name, ms, per_clk_sm, clk / 1000,
           elems / (ms * 1e-3) / 1e12);
    (void)flops_per_op;
}

int main() {
    float* out; cudaMalloc(&out, 1 << 26);
    run<0>("FFMA scalar", out, 2);
    run<1>("FADD scalar", out, 1);
    run<2>("FMUL scalar", out, 1);
    run<3>("FFMA2 packed", out, 2);
    run<4>("FADD2 packed", out, 1);
    run<5>("FMUL2 packed", out, 1);
    run<6>("MUFU rsq+rcp", out, 1);
    run<7>("FFMA2 + FMNMX", out, 1);
    return 0;
}
